// Gathered-row GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
//   out[r(i), :] = epi( pro(a[r(i), :]) . B )        m rows (gathered through `rows`), K, N <= 128
//
// These contractions (X.W^T, the DeletionLayer's x[mask].W_del, their input gradients) have an
// arithmetic intensity of 21-32 flop/B: HBM-bound on tensor cores, FMA-bound on CUDA cores.
// fp32 parity (1e-5) is kept with the 3xTF32 split  a = a_hi + a_lo, b = b_hi + b_lo,
//   a.b ~= a_hi.b_hi + a_lo.b_hi + a_hi.b_lo      (a_hi = a rounded to tf32)
// which costs 3 MMAs per k-step and still leaves the kernel memory-bound.
//
// Structure (one persistent CTA per SM, 416 threads):
//   warps 0-7  producers : gather 128 rows x 32 k (one 128-byte swizzle row each) per stage with
//                          128-bit loads kept 3 stages ahead in a register ring, optional ReLU,
//                          hi/lo split, st.shared in the canonical K-major SWIZZLE_128B layout,
//                          fence.proxy.async, arrive on full[stage];
//   warp  12   MMA issue : one elected lane waits full[stage], issues 12 tcgen05.mma
//                          (M=128, N, K=8, kind::tf32) per stage, tcgen05.commit -> empty[stage],
//                          and -> tmem_full[acc] after the last stage of a tile;
//   warps 8-11 epilogue  : tcgen05.ld the 128 x N fp32 accumulators (thread = row), bias / row
//                          scale / ReLU / gate, 128-bit stores to the (scattered) output rows,
//                          arrive on tmem_empty[acc].
// B (<= 128 x 128, hi and lo) is staged once per CTA and stays resident in shared memory; the
// accumulator is double buffered in TMEM so the epilogue of tile i overlaps the loads and MMAs
// of tile i+1.
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace gd {
namespace tc {

constexpr int MAX_STAGES = 4;      // A ring depth (runtime: as many as shared memory allows)
constexpr int NUM_PRODUCER_WARPS = 8;
constexpr int MMA_WARP = 12;                     // warps 8-11: epilogue (warp % 4 = TMEM lane quarter)
constexpr int NUM_THREADS = 13 * 32;
constexpr int EPI_LD = 36;                       // padded row (floats) of the per-warp epilogue transpose tile
constexpr int EPI_BYTES = 4 * 32 * EPI_LD * 4;
constexpr int PREFETCH = 4;                      // producer register ring depth (stages of loads in flight)


// Warp roles of gemm_rows_tc_kernel: 16 producers, 4 epilogue warps (one per TMEM lane quarter), 1 MMA issuer.  Measured (tools/gemm_sweep.py with GD_TC_DEBUG):
// with 4 epilogue warps running a flag-generic epilogue the kernel was epilogue bound at 5.5 us per
// 128 x 128 tile whatever the producers did; the epilogue is therefore specialised at compile time
// (EPI_* bits) and spread over twice the warps.
constexpr int ROWS_PRODUCER_WARPS = 16;                       // 512 threads: 8 per row, 64 rows per pass, 2 passes per stage
constexpr int ROWS_PASSES = 128 / (ROWS_PRODUCER_WARPS * 4);
constexpr int ROWS_EPI_WARPS = 4;
constexpr int ROWS_MMA_WARP = ROWS_PRODUCER_WARPS + ROWS_EPI_WARPS;
constexpr int ROWS_THREADS = (ROWS_MMA_WARP + 1) * 32;
constexpr int RE_COLS = 16;                                   // columns per epilogue step
constexpr int RE_LD = RE_COLS;                                // row (floats) of the per-warp transpose tile; 16-byte chunk c of row r sits at c ^ ((r >> 1) & 3):
                                                              // conflict-free for the row-per-lane writes and the 4-lanes-per-row reads
constexpr int ROWS_EPI_BYTES = ROWS_EPI_WARPS * 32 * RE_LD * 4;
enum : int { EPI_BIAS = 1, EPI_SCALE = 2, EPI_RELU = 4, EPI_BITS = 8, EPI_GATE = 16, EPI_ALL = 31 };

template <int EPI>
__global__ void __launch_bounds__(ROWS_THREADS, 1) gemm_rows_tc_kernel(const Args g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve (all operand tiles 1024-byte aligned)
    // align by OFFSET (not through an integer cast) so the compiler keeps the shared address space: the cast version
    // compiled every tile store to a generic ST.E and the proxy fence to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int kchunks = (g.k + KC - 1) / KC;
    const int b_tile = g.n * 128;                                  // bytes of one [n x 128 B] B tile
    uint8_t* b_hi = smem;                                          // [kchunks][n x 128 B]
    uint8_t* b_lo = b_hi + kchunks * b_tile;
    uint8_t* a_ring = b_lo + kchunks * b_tile;                     // [stages][hi 16 KB | lo 16 KB]
    a_ring += (1024u - (smem_u32(a_ring) & 1023u)) & 1023u;
    float* epi_buf = reinterpret_cast<float*>(a_ring + g.stages * 2 * TILE_BYTES);
    __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], tfull_bar[2], tempty_bar[2];
    const int STAGES = g.stages;
    __shared__ uint32_t tmem_base_smem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // per tile: main accumulator (a_hi.b_hi) + correction accumulator (a_lo.b_hi + a_hi.b_lo), double buffered.
    // The tensor core truncates when it adds into the fp32 accumulator; keeping the 2^-11-scaled correction
    // terms out of the main chain cuts that (biased) rounding from 3 to 1 accumulation per k-step.
    const uint32_t tmem_cols = g.n <= 32 ? 128 : (g.n <= 64 ? 256 : 512);

    if (tid == 0) {
        // one arrival per warp (after a __syncwarp): 256 per-thread arrivals on one mbarrier serialise
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], ROWS_PRODUCER_WARPS); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], ROWS_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == ROWS_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    pdl_wait();                      // barrier init / TMEM allocation above overlap the predecessor's tail
    pdl_trigger();
    // ---- stage B once (hi / lo, K-major SWIZZLE_128B), all threads; 4 chunks per thread in flight
    {
        const int total = kchunks * g.n * 8;
        const bool vec = g.b_is_nk && (g.k & 3) == 0 && ((uintptr_t)g.b & 15) == 0;
        for (int base = tid; base < total; base += 4 * ROWS_THREADS) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * ROWS_THREADS;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx < total) {
                    const int c = idx / (g.n * 8), rem = idx - c * g.n * 8;
                    const int nn = rem >> 3, j = rem & 7;
                    const int kk0 = c * KC + j * 4;
                    if (vec && kk0 + 4 <= g.k) v[u] = __ldg(reinterpret_cast<const float4*>(g.b + (int64_t)nn * g.k + kk0));
                    else {
                        float* vp = reinterpret_cast<float*>(&v[u]);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int kk = kk0 + e;
                            if (kk < g.k) vp[e] = g.b_is_nk ? __ldg(g.b + (int64_t)nn * g.k + kk) : __ldg(g.b + (int64_t)kk * g.n + nn);
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * ROWS_THREADS;
                if (idx < total) {
                    const int c = idx / (g.n * 8), rem = idx - c * g.n * 8;
                    const int nn = rem >> 3, j = rem & 7;
                    float4 hi, lo;
                    split4(v[u], hi, lo);
                    *reinterpret_cast<float4*>(b_hi + c * b_tile + swz(nn, j)) = hi;
                    *reinterpret_cast<float4*>(b_lo + c * b_tile + swz(nn, j)) = lo;
                }
            }
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp < ROWS_PRODUCER_WARPS) {
        // ================================ producers ================================
        // 512 threads: 8 threads per row (one 16-byte chunk each), 64 rows per pass, 2 passes per stage (the per-stage
        // latency of a producer warp, ~1200 cycles with 4 passes, was the floor of the kernel: twice the warps, half the work each).
        // Loads run PREFETCH stages ahead of the shared-memory stores through a register ring, so
        // ~48 KB of gathers are in flight per SM while earlier stages are split / stored / consumed.
        const int j = tid & 7, r0 = tid >> 3;
        int p_tile = blockIdx.x, p_c = 0;                         // prefetch cursor (tile, k-chunk)
        const float* psrc[ROWS_PASSES];
        int32_t rid_next[ROWS_PASSES];
        auto load_rids = [&](int tile, int32_t (&rid)[ROWS_PASSES]) {
#pragma unroll
            for (int p = 0; p < ROWS_PASSES; ++p) {
                const int64_t i = (int64_t)tile * BM + r0 + (128 / ROWS_PASSES) * p;
                rid[p] = (tile < g.num_tiles && i < g.m) ? (g.rows ? __ldg(g.rows + i) : (int32_t)i) : -1;
            }
        };
        auto set_src = [&](const int32_t (&rid)[ROWS_PASSES]) {
#pragma unroll
            for (int p = 0; p < ROWS_PASSES; ++p) psrc[p] = rid[p] >= 0 ? g.a + (int64_t)rid[p] * g.lda + j * 4 : nullptr;
        };
        // Whole rows of the tile after next are pulled into L2 with one bulk prefetch per row: the stage loads
        // below touch a row in four 128-byte pieces spread over time, which DRAM serves poorly on its own.
        const bool pf_ok = (g.k & 3) == 0 && !(g.debug & 32);
        auto l2_prefetch_rows = [&](const int32_t (&rid)[ROWS_PASSES]) {
            if (j != 0 || !pf_ok) return;
#pragma unroll
            for (int p = 0; p < ROWS_PASSES; ++p)
                if (rid[p] >= 0)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(g.a + (int64_t)rid[p] * g.lda), "r"(g.k * 4) : "memory");
        };
        auto issue = [&](float4 (&buf)[ROWS_PASSES]) {                       // loads of stage (p_tile, p_c); advance cursor
            if (p_tile >= g.num_tiles) return;
            const int kbase = p_c * KC + j * 4;
#pragma unroll
            for (int p = 0; p < ROWS_PASSES; ++p) {
                buf[p] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (psrc[p] != nullptr && !(g.debug & 2)) {
                    if (kbase + 4 <= g.k) buf[p] = __ldg(reinterpret_cast<const float4*>(psrc[p] + p_c * KC));
                    else {
                        float* vp = reinterpret_cast<float*>(&buf[p]);
                        for (int e = 0; e < 4; ++e) if (kbase + e < g.k) vp[e] = __ldg(psrc[p] + p_c * KC + e);
                    }
                }
            }
            if (++p_c == kchunks) {
                p_c = 0;
                p_tile += gridDim.x;
                set_src(rid_next);
                load_rids(p_tile + gridDim.x, rid_next);           // row ids one tile ahead of the prefetch cursor
                l2_prefetch_rows(rid_next);
            }
        };
        uint32_t stage = 0, phase = 0;
        auto consume = [&](const float4 (&buf)[ROWS_PASSES]) {               // split + store one stage
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* ahi = a_ring + stage * 2 * TILE_BYTES;
            uint8_t* alo = ahi + TILE_BYTES;
#pragma unroll
            for (int p = 0; p < ROWS_PASSES; ++p) {
                float4 x = buf[p];
                if (g.relu_in) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
                float4 hi, lo;
                split4(x, hi, lo);
                const uint32_t o = swz(r0 + (128 / ROWS_PASSES) * p, j);
                *reinterpret_cast<float4*>(ahi + o) = hi;
                *reinterpret_cast<float4*>(alo + o) = lo;
            }
            // No proxy fence here: ptxas lowers fence.proxy.async to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, and the MEMBAR
            // waits for this thread's PREFETCHED global loads (the next stages) - it serialised the register ring.
            // The stores are published by the (release) arrive; the MMA thread, which has no loads in flight,
            // executes the generic->async proxy fence after it has acquired the barrier.
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        };
        {
            int32_t rid0[ROWS_PASSES];
            load_rids(p_tile, rid0);
            set_src(rid0);
            load_rids(p_tile + gridDim.x, rid_next);
            l2_prefetch_rows(rid_next);
        }
        const int my_tiles = blockIdx.x < g.num_tiles ? (g.num_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
        const int total = my_tiles * kchunks;
        float4 buf[PREFETCH][ROWS_PASSES];
#pragma unroll
        for (int d = 0; d < PREFETCH; ++d) issue(buf[d]);
        for (int s = 0; s < total; s += PREFETCH) {
#pragma unroll
            for (int d = 0; d < PREFETCH; ++d) {
                if (s + d < total) {
                    consume(buf[d]);
                    issue(buf[d]);
                }
            }
        }
    } else if (warp == ROWS_MMA_WARP) {
        // ================================ MMA issuer ================================
        if (elect_one()) {
            const uint32_t idesc = make_idesc(g.n);
            const uint64_t da0 = make_desc(smem_u32(a_ring)), db_hi0 = make_desc(smem_u32(b_hi)), db_lo0 = make_desc(smem_u32(b_lo));
            uint32_t stage = 0, phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
                const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + acc * 2 * g.n, dc = d + g.n;
                for (int c = 0; c < kchunks; ++c) {
                    mbar_wait(&full_bar[stage], phase);
                    fence_proxy_async();                           // the producers' generic stores -> async proxy (see consume())
                    tc_fence_after();
                    const uint64_t da_hi = desc_advance(da0, stage * 2 * TILE_BYTES), da_lo = desc_advance(da_hi, TILE_BYTES);
                    const uint64_t db_hi = desc_advance(db_hi0, c * b_tile), db_lo = desc_advance(db_lo0, c * b_tile);
                    if (!(g.debug & 4)) {
#pragma unroll
                        for (int ks = 0; ks < KC / 8; ++ks) {       // one k-step = 8 tf32 = 32 bytes along the swizzled row
                            const uint32_t accum = (ks != 0) || (c != 0);
                            umma_tf32(dc, da_lo + 2 * ks, db_hi + 2 * ks, idesc, accum);
                            umma_tf32(dc, da_hi + 2 * ks, db_lo + 2 * ks, idesc, 1);
                            umma_tf32(d, da_hi + 2 * ks, db_hi + 2 * ks, idesc, accum);
                        }
                    }
                    umma_commit(&empty_bar[stage]);                // frees the smem stage when the MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[acc]);                      // accumulator of this tile complete
            }
        }
        __syncwarp();
    } else {
        // ================================ epilogue ================================
        // thread = accumulator row (TMEM lane).  A warp owns one TMEM lane quarter (all column groups);
        // each 16-column step is transposed through a per-warp shared-memory tile so
        // that the (scattered) output rows are written as 64-byte runs (4 lanes x 16 B per row, 8 rows
        // per instruction).
        const int ew = warp - ROWS_PRODUCER_WARPS;
        const int q = warp & 3;                                    // TMEM lane quarter (hardware: warp id % 4)
        const int lr = q * 32 + lane;                              // row inside the tile == TMEM lane
        float* tbuf = epi_buf + ew * (32 * RE_LD);
        const int cl = lane & 3, rl = lane >> 2;                   // coalesced phase: 16-byte chunk / row-in-group
        const bool has_bias = (EPI & EPI_BIAS) && g.bias != nullptr;
        const bool has_scale = (EPI & EPI_SCALE) && g.out_scale != nullptr;
        const bool relu_out = (EPI & EPI_RELU) && g.relu_out;
        const bool has_gbits = (EPI & EPI_BITS) && g.gate_bits != nullptr;
        const bool has_mask = (EPI & EPI_BITS) && g.relu_mask_out != nullptr;
        const bool has_gate = (EPI & EPI_GATE) && g.gate != nullptr;
        const int nw = g.n >> 5;
        int it = 0;
        for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
            const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
            const int64_t gi = (int64_t)tile * BM + lr;
            const int32_t r = gi < g.m ? (g.rows ? __ldg(g.rows + gi) : (int32_t)gi) : -1;
            const float sc = (has_scale && r >= 0) ? __ldg(g.out_scale + r) : 1.0f;
            int32_t rr[4];                                         // rows of the coalesced phase (constant per tile)
#pragma unroll
            for (int i = 0; i < 4; ++i) rr[i] = __shfl_sync(0xffffffffu, r, i * 8 + rl);
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            if (!(g.debug & 1)) {
                for (int c32 = 0; c32 < g.n; c32 += 32) {
                    uint32_t gword = 0xffffffffu, pos_bits = 0;
                    if (has_gbits && r >= 0) gword = __ldg(g.gate_bits + (int64_t)r * nw + (c32 >> 5));
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int c0 = c32 + h * RE_COLS;
                        uint32_t v[16], vc[16];
                        const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 2 * g.n + c0;
                        if (!(g.debug & 16)) {
                            tmem_ld16_nowait(t0, v);
                            tmem_ld16_nowait(t0 + g.n, vc);
                            tmem_wait_ld();
                        } else {
#pragma unroll
                            for (int e = 0; e < 16; ++e) { v[e] = 0; vc[e] = 0; }
                        }
#pragma unroll
                        for (int e = 0; e < RE_COLS; e += 4) {
                            float o[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                float x = __uint_as_float(v[e + u]) + __uint_as_float(vc[e + u]);
                                if (has_bias) x += __ldg(g.bias + c0 + e + u);
                                if (has_scale) x *= sc;
                                if (relu_out) x = fmaxf(x, 0.f);
                                if (EPI & EPI_BITS) {
                                    const int bit = h * RE_COLS + e + u;
                                    if (!((gword >> bit) & 1u)) x = 0.f;
                                    if (x > 0.f) pos_bits |= 1u << bit;
                                }
                                o[u] = x;
                            }
                            *reinterpret_cast<float4*>(tbuf + lane * RE_LD + (((e >> 2) ^ ((lane >> 1) & 3)) << 2)) = make_float4(o[0], o[1], o[2], o[3]);
                        }
                        __syncwarp();
                        float4 gt[4];
                        if (has_gate) {                            // fp32 gate rows first (all loads in flight)
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                gt[i] = make_float4(1.f, 1.f, 1.f, 1.f);
                                if (rr[i] >= 0)
                                    gt[i] = __ldg(reinterpret_cast<const float4*>(g.gate + (int64_t)rr[i] * g.ldgate + c0 + cl * 4));
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float4 o = lds128(tbuf + (i * 8 + rl) * RE_LD + ((cl ^ (((i * 8 + rl) >> 1) & 3)) << 2));
                            if (has_gate) {
                                if (!(gt[i].x > 0.f)) o.x = 0.f;
                                if (!(gt[i].y > 0.f)) o.y = 0.f;
                                if (!(gt[i].z > 0.f)) o.z = 0.f;
                                if (!(gt[i].w > 0.f)) o.w = 0.f;
                            }
                            if (rr[i] >= 0 && !((g.debug & 8) && o.x != 123.456f))
                                *reinterpret_cast<float4*>(g.out + (int64_t)rr[i] * g.ldo + c0 + cl * 4) = o;
                        }
                        __syncwarp();
                    }
                    if (has_mask && r >= 0) g.relu_mask_out[(int64_t)r * nw + (c32 >> 5)] = pos_bits;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
    }
    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == ROWS_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
    }
}

constexpr size_t SMEM_LIMIT = 227 * 1024 - 4096;   // leave room for the static barriers / row ids
static size_t fixed_bytes(int k, int n) {
    const int kchunks = (k + KC - 1) / KC;
    return 1024 + (size_t)2 * kchunks * n * 128 + 1024 + ROWS_EPI_BYTES;
}
static int num_stages(int k, int n) {
    const size_t fixed = fixed_bytes(k, n);
    if (fixed + 2 * TILE_BYTES > SMEM_LIMIT) return 0;
    return (int)std::min<size_t>(MAX_STAGES, (SMEM_LIMIT - fixed) / (2 * TILE_BYTES));
}
static size_t smem_bytes(int k, int n) { return fixed_bytes(k, n) + (size_t)num_stages(k, n) * 2 * TILE_BYTES; }


// =====================================================================================
// Weight-gradient contraction  c[k1, n2] = sum_i a_scale[r(i)] * pro(a[r(i), :k1])^T (x) g[r(i), :n2]
// (dW_del = x[S]^T . dOut[S]) on the tensor cores.  The contraction index is the ROW, so both
// operands are transposed while they are staged: a stage holds 32 rows as A^T [128 x 32] and
// G^T [n2 x 32] (K-major, SWIZZLE_128B, hi / lo).  Every CTA owns a contiguous range of rows and
// accumulates in TMEM; every FLUSH stages the accumulator pair is drained by the epilogue warps
// into the CTA's partial buffer (plain fp32 adds, fixed order) while the MMAs continue on the
// other pair — this bounds the tensor core's truncating accumulation chain to 32 steps.  A tiny
// second kernel adds the per-CTA partials in CTA order: deterministic, no atomics.
constexpr int TN_FLUSH = 8;         // stages (x32 rows) per accumulator flush
constexpr int TN_PREFETCH = 3;      // producer register ring depth (stages of row loads in flight)
constexpr int TN_ATOM_COL = 4096;   // bytes of one 32-feature atom column of a stage: 8 k-atoms (4 rows x 128 B) = 32 rows

__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tn_tc_kernel(const TnArgs t) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // align by OFFSET (not through an integer cast) so the compiler keeps the shared address space: the cast version
    // compiled every tile store to a generic ST.E and the proxy fence to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int g_tile = t.n2 * 128;                                  // bytes of one [n2 x 128 B] tile
    const int stage_bytes = 2 * TILE_BYTES + 2 * g_tile;            // A^T hi | A^T lo | G^T hi | G^T lo
    const int STAGES = t.stages;
    float* epi_buf = reinterpret_cast<float*>(smem + STAGES * stage_bytes);
    __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_smem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t tmem_cols = t.n2 <= 32 ? 128 : (t.n2 <= 64 ? 256 : 512);
    if (tid == 0) {
        for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(&full_bar[s], NUM_PRODUCER_WARPS); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    pdl_wait();
    pdl_trigger();
    // rows k1..127 of the A^T tiles are never written by the producers: zero every A^T tile once
    for (int i = tid; i < STAGES * 2 * TILE_BYTES / 16; i += NUM_THREADS) {
        const int s = i / (2 * TILE_BYTES / 16), o = i - s * (2 * TILE_BYTES / 16);
        *reinterpret_cast<float4*>(smem + s * stage_bytes + o * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    const int64_t r_beg = (int64_t)blockIdx.x * t.rows_per_cta;
    const int64_t r_end = min(t.m, r_beg + t.rows_per_cta);
    const int nstages = r_end > r_beg ? (int)((r_end - r_beg + KC - 1) / KC) : 0;   // 32 rows per stage
    const int ngroups = (nstages + TN_FLUSH - 1) / TN_FLUSH;

    if (warp < NUM_PRODUCER_WARPS) {
        // ------------------------------ producers: MN-major stage fill ------------------------------
        // The contraction index is the ROW, so row-major A[rows, k1] / G[rows, n2] ARE the MN-major operand
        // layouts of tcgen05 (features contiguous, one k-row per matrix row): no transposition, every lane
        // copies 16-byte chunks (SWIZZLE_128B_BASE32B placement).  8 lanes cover the 128 B of one k-row of one 32-feature atom (a conflict-free
        // quarter-warp store), a warp covers 4 rows, the 8 warps the 32 rows of a stage.
        const int rr = warp * 4 + (lane >> 3), cj = lane & 7;      // row of the stage, 16-byte chunk inside an atom row
        const int a4 = t.k1 >> 2, g4 = t.n2 >> 2;                   // float4 per row
        const uint32_t row_off = (uint32_t)((rr >> 2) * 512 + (rr & 3) * 128 + ((((cj >> 1) ^ (rr & 3)) << 5) | ((cj & 1) << 4)));
        uint32_t stage = 0, phase = 0;
        auto fetch_rid = [&](int s) -> int32_t {
            const int64_t i = r_beg + (int64_t)s * KC + rr;
            return (s < nstages && i < r_end) ? (t.rows ? __ldg(t.rows + i) : (int32_t)i) : -1;
        };
        // Register ring: the loads of TN_PREFETCH stages are in flight while earlier stages are split and stored
        // (the row id of the stage after those is already loaded, so an issue is never two dependent latencies).
        float4 va[TN_PREFETCH][4], vg[TN_PREFETCH][4];
        float sc[TN_PREFETCH];
        int s_issue = 0;
        int32_t rid_nxt = fetch_rid(0);
        auto issue = [&](float4 (&xa)[4], float4 (&xg)[4], float& xs) {
            const int32_t r = rid_nxt;
            ++s_issue;
            rid_nxt = fetch_rid(s_issue);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                xa[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                xg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r >= 0) {
                    if (cj + 8 * i < a4) xa[i] = __ldg(reinterpret_cast<const float4*>(t.a + (int64_t)r * t.lda) + cj + 8 * i);
                    if (cj + 8 * i < g4) xg[i] = __ldg(reinterpret_cast<const float4*>(t.g + (int64_t)r * t.ldg) + cj + 8 * i);
                }
            }
            xs = (r >= 0 && t.a_scale) ? __ldg(t.a_scale + r) : 1.f;
        };
        auto consume = [&](const float4 (&xa)[4], const float4 (&xg)[4], float xs) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* at_hi = smem + stage * stage_bytes + row_off;
            uint8_t* at_lo = at_hi + TILE_BYTES;
            uint8_t* gt_hi = at_lo + TILE_BYTES;
            uint8_t* gt_lo = gt_hi + g_tile;
#pragma unroll
            for (int i = 0; i < 4; ++i) {                          // atom i = features 32 i .. 32 i + 31, 4 KB apart
                if (cj + 8 * i < a4) {
                    float4 x = xa[i];
                    if (t.relu_a) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
                    x.x *= xs; x.y *= xs; x.z *= xs; x.w *= xs;
                    float4 hi, lo;
                    split4(x, hi, lo);
                    *reinterpret_cast<float4*>(at_hi + i * TN_ATOM_COL) = hi;
                    *reinterpret_cast<float4*>(at_lo + i * TN_ATOM_COL) = lo;
                }
                if (cj + 8 * i < g4) {
                    float4 hi, lo;
                    split4(xg[i], hi, lo);
                    *reinterpret_cast<float4*>(gt_hi + i * TN_ATOM_COL) = hi;
                    *reinterpret_cast<float4*>(gt_lo + i * TN_ATOM_COL) = lo;
                }
            }
            // (no proxy fence on this side: see gemm_rows_tc_kernel's consume())
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[stage]);               // one arrival per warp
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        };
#pragma unroll
        for (int d = 0; d < TN_PREFETCH; ++d) issue(va[d], vg[d], sc[d]);
        for (int s = 0; s < nstages; s += TN_PREFETCH) {
#pragma unroll
            for (int d = 0; d < TN_PREFETCH; ++d) {
                if (s + d < nstages) {
                    consume(va[d], vg[d], sc[d]);
                    issue(va[d], vg[d], sc[d]);
                }
            }
        }
    } else if (warp == MMA_WARP) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc_mn(t.n2), idesc2 = make_idesc_mn(2 * t.n2);
            // stage layout: A^T hi | A^T lo | G^T hi | G^T lo; rows 8 ks .. 8 ks + 7 of a stage = two 4-row k-atoms (512 B each)
            // inside every 4 KB feature-atom column, i.e. + 1024 bytes per k-step
            const uint64_t da0 = make_desc_mn(smem_u32(smem), TN_ATOM_COL, 512);
            uint32_t stage = 0, phase = 0;
            for (int grp = 0; grp < ngroups; ++grp) {
                const uint32_t acc = grp & 1, acc_phase = (grp >> 1) & 1;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + acc * 2 * t.n2, dc = d + t.n2;
                const int s_end = min(nstages, (grp + 1) * TN_FLUSH);
                for (int s = grp * TN_FLUSH; s < s_end; ++s) {
                    mbar_wait(&full_bar[stage], phase);
                    fence_proxy_async();                           // the producers' generic stores -> async proxy
                    tc_fence_after();
                    const uint64_t da_hi = desc_advance(da0, stage * stage_bytes), da_lo = desc_advance(da_hi, TILE_BYTES);
                    const uint64_t db_hi = desc_advance(da_lo, TILE_BYTES);        // G^T hi, immediately followed by G^T lo
                    const bool first = (s == grp * TN_FLUSH);
                    // G^T lo follows G^T hi in atom-column order (same lbo), and the correction accumulator follows the main one
                    // in TMEM: a_hi x [g_hi | g_lo] is ONE instruction of N = 2 n2 - two MMAs per k-step instead of three
#pragma unroll
                    for (int ks = 0; ks < KC / 8; ++ks) {
                        const uint32_t accum = !(first && ks == 0);
                        umma_tf32(d, da_hi + 64 * ks, db_hi + 64 * ks, idesc2, accum);      // main | corr (+)= a_hi x [g_hi | g_lo]
                        umma_tf32(dc, da_lo + 64 * ks, db_hi + 64 * ks, idesc, 1);          // corr += a_lo x g_hi
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[acc]);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------ epilogue: drain a flush group into the CTA's partial ------------------------------
        const int q = warp & 3;
        const int f = q * 32 + lane;                               // TMEM lane == row of c
        float* tbuf = epi_buf + q * (32 * EPI_LD);
        const int cl = lane & 7, rl = lane >> 3;
        float* part = t.partial + (int64_t)blockIdx.x * t.k1 * t.n2;
        for (int grp = 0; grp < ngroups; ++grp) {
            const uint32_t acc = grp & 1, acc_phase = (grp >> 1) & 1;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            for (int c0 = 0; c0 < t.n2; c0 += 32) {
                uint32_t v[32], vc[32];
                const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 2 * t.n2 + c0;
                tmem_ld32(t0, v);
                tmem_ld32(t0 + t.n2, vc);
#pragma unroll
                for (int e = 0; e < 32; e += 4)
                    *reinterpret_cast<float4*>(tbuf + lane * EPI_LD + e) =
                        make_float4(__uint_as_float(v[e]) + __uint_as_float(vc[e]), __uint_as_float(v[e + 1]) + __uint_as_float(vc[e + 1]),
                                    __uint_as_float(v[e + 2]) + __uint_as_float(vc[e + 2]), __uint_as_float(v[e + 3]) + __uint_as_float(vc[e + 3]));
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = q * 32 + i * 4 + rl;
                    if (row < t.k1) {
                        float4 o = *reinterpret_cast<const float4*>(tbuf + (i * 4 + rl) * EPI_LD + cl * 4);
                        float4* dst = reinterpret_cast<float4*>(part + (int64_t)row * t.n2 + c0 + cl * 4);
                        if (grp > 0) { const float4 old = *dst; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
                        *dst = o;
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
        if (ngroups == 0) {                                         // CTA without rows: its partial is zero
            if (f < t.k1) for (int c = 0; c < t.n2; ++c) part[(int64_t)f * t.n2 + c] = 0.f;
        }
        (void)f;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
    }
}

static int tn_stages(int n2) {
    const size_t stage = 2 * TILE_BYTES + (size_t)2 * n2 * 128;
    return (int)std::min<size_t>(MAX_STAGES, (SMEM_LIMIT - 1024 - EPI_BYTES) / stage);
}

}  // namespace tc
}  // namespace gd

using namespace gd;

// 1 if gd_gemm_rows_tc supports the shape (else the caller uses gd_gemm_rows)
extern "C" int gd_gemm_rows_tc_supported(int32_t k, int32_t n, int64_t lda, int64_t ldo) {
    if (k <= 0 || n <= 0 || n > 128 || n % 32 != 0) return 0;
    if (lda % 4 != 0 || ldo % 4 != 0) return 0;
    return tc::num_stages(k, n) >= 2 ? 1 : 0;
}

extern "C" int gd_gemm_rows_tc(const float* a, int64_t lda, const int32_t* rows, int64_t m, int32_t k,
                               const float* b, int32_t b_is_nk, int32_t n, const float* bias,
                               const float* out_scale, const float* gate, int64_t ldgate, int32_t relu_in,
                               int32_t relu_out, float* out, int64_t ldo, uint32_t* relu_mask_out,
                               const uint32_t* gate_bits, gd_stream_t stream_) {
    cudaStream_t stream = as_stream(stream_);
    GD_CHECK_ARG(m >= 0 && k > 0 && n > 0, "bad shape");
    if (m == 0) return GD_OK;
    GD_CHECK_ARG(a && b && out, "null pointer");
    GD_CHECK_ARG(gd_gemm_rows_tc_supported(k, n, lda, ldo), "shape not supported by the tcgen05 path");
    GD_CHECK_ARG(((uintptr_t)a | (uintptr_t)out | (uintptr_t)gate) % 16 == 0 && (!gate || ldgate % 4 == 0), "operands must be 16-byte aligned");
    tc::Args g{a, lda, rows, m, k, b, b_is_nk, n, bias, out_scale, gate, ldgate, relu_in, relu_out, out, ldo,
               relu_mask_out, gate_bits, (int)ceil_div<int64_t>(m, tc::BM), tc::num_stages(k, n), 0};
    { const char* e = getenv("GD_TC_DEBUG"); g.debug = e ? atoi(e) : 0; }     // measurement only; read per call so one process can sweep it
    // weights-in-TMEM kernel (gemm_tc_wt.cu) for the shapes / epilogues of the Del-training epoch; GD_GEMM_ROWS=ring keeps this one
    {
        const char* e = getenv("GD_GEMM_ROWS");
        const bool want_wt = !(e && strcmp(e, "ring") == 0);
        if (want_wt && tc::rows_wt_supported(g)) return tc::launch_rows_wt(g, stream);
    }
    const size_t smem = tc::smem_bytes(k, n);
    const int grid = std::min(g.num_tiles, kNumSMs);
    const int need = (bias ? tc::EPI_BIAS : 0) | (out_scale ? tc::EPI_SCALE : 0) | (relu_out ? tc::EPI_RELU : 0) |
                     ((relu_mask_out || gate_bits) ? tc::EPI_BITS : 0) | (gate ? tc::EPI_GATE : 0);
    auto launch = [&](auto kern) -> int {
        GD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GD_CUDA(launch_pdl(kern, grid, tc::ROWS_THREADS, smem, stream, g));
        GD_LAUNCH_CHECK();
        return GD_OK;
    };
    switch (need) {                        // the epilogues of the Del-training epoch are compiled without the unused branches
        case 0: return launch(tc::gemm_rows_tc_kernel<0>);
        case tc::EPI_SCALE: return launch(tc::gemm_rows_tc_kernel<tc::EPI_SCALE>);
        case tc::EPI_BITS: return launch(tc::gemm_rows_tc_kernel<tc::EPI_BITS>);
        case tc::EPI_SCALE | tc::EPI_BITS: return launch(tc::gemm_rows_tc_kernel<tc::EPI_SCALE | tc::EPI_BITS>);
        default: return launch(tc::gemm_rows_tc_kernel<tc::EPI_ALL>);
    }
}

extern "C" int gd_gemm_rows_tc_batch(const float* a, int64_t lda, int64_t m, int32_t k, const float* b, int64_t b_stride,
                                     int32_t b_is_nk, int32_t n, float* out, int64_t out_stride, int64_t ldo, int32_t batch,
                                     gd_stream_t stream) {
    GD_CHECK_ARG(batch >= 0, "negative batch");
    for (int32_t i = 0; i < batch; ++i) {
        const int rc = gd_gemm_rows_tc(a, lda, nullptr, m, k, b + (int64_t)i * b_stride, b_is_nk, n, nullptr, nullptr, nullptr, 0,
                                       0, 0, out + (int64_t)i * out_stride, ldo, nullptr, nullptr, stream);
        if (rc != GD_OK) return rc;
    }
    return GD_OK;
}

// out[i] = sum_p partial[p][i] in a fixed order: warp w of a block adds the partials p = w, w + 8, ... for 32
// consecutive elements (coalesced, all loads independent), then the 8 warp sums are added in warp order.
// tr_k1 > 0: the partials are [tr_n][tr_k1] (gemm_dxdw_wt.cu) and `out` is [tr_k1][tr_n].
__global__ void __launch_bounds__(256) tn_reduce_kernel(const float* __restrict__ partial, int nparts, int64_t count,
                                                        float* __restrict__ out, int tr_k1, int tr_n) {
    __shared__ float red[8][32];
    gd::pdl_wait();
    gd::pdl_trigger();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * 32 + lane;
    float s = 0.f;
    if (i < count) {
#pragma unroll 4
        for (int p = w; p < nparts; p += 8) s += __ldg(partial + (int64_t)p * count + i);
    }
    red[w][lane] = s;
    __syncthreads();
    if (w == 0 && i < count) {
        float r = red[0][lane];
#pragma unroll
        for (int k = 1; k < 8; ++k) r += red[k][lane];
        out[tr_k1 > 0 ? (i % tr_k1) * tr_n + i / tr_k1 : i] = r;
    }
}

namespace gd { namespace tc {
int launch_tn_reduce(const float* partial, int nparts, int64_t count, float* out, cudaStream_t stream, int tr_k1, int tr_n) {
    GD_CUDA(launch_pdl(tn_reduce_kernel, (unsigned)ceil_div<int64_t>(count, 32), 256, 0, stream, partial, nparts, count, out, tr_k1, tr_n));
    GD_LAUNCH_CHECK();
    return GD_OK;
}
}}  // namespace gd::tc

extern "C" int gd_gemm_tn_rows_tc_supported(int32_t k1, int32_t n2, int64_t lda, int64_t ldg) {
    if (k1 <= 0 || k1 > 128 || k1 % 4 != 0 || n2 <= 0 || n2 > 128 || n2 % 32 != 0) return 0;
    if (lda % 4 != 0 || ldg % 4 != 0) return 0;
    return tc::tn_stages(n2) >= 2 ? 1 : 0;
}

extern "C" size_t gd_gemm_tn_tc_workspace_bytes(int32_t k1, int32_t n2) { return (size_t)kNumSMs * k1 * n2 * sizeof(float); }

extern "C" int gd_gemm_tn_rows_tc(const float* a, int64_t lda, const float* g, int64_t ldg, const int32_t* rows,
                                  int64_t m, int32_t k1, int32_t n2, int32_t relu_a, const float* a_scale, float* c,
                                  void* workspace, size_t workspace_bytes, gd_stream_t stream_) {
    cudaStream_t stream = as_stream(stream_);
    GD_CHECK_ARG(m >= 0 && k1 > 0 && n2 > 0, "bad shape");
    GD_CHECK_ARG(c != nullptr, "null output");
    if (m == 0) { GD_CUDA(cudaMemsetAsync(c, 0, (size_t)k1 * n2 * sizeof(float), stream)); return GD_OK; }
    GD_CHECK_ARG(a && g, "null pointer");
    GD_CHECK_ARG(gd_gemm_tn_rows_tc_supported(k1, n2, lda, ldg), "shape not supported by the tcgen05 path");
    GD_CHECK_ARG(((uintptr_t)a | (uintptr_t)g) % 16 == 0, "operands must be 16-byte aligned");
    if (!workspace || workspace_bytes < gd_gemm_tn_tc_workspace_bytes(k1, n2))
        return fail(GD_ERR_WORKSPACE, "gd_gemm_tn_rows_tc: workspace too small");
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(kNumSMs, ceil_div<int64_t>(m, 32 * tc::TN_FLUSH)));
    const int64_t rows_per_cta = ceil_div<int64_t>(ceil_div<int64_t>(m, grid), 32) * 32;
    tc::TnArgs t{a, lda, g, ldg, rows, m, k1, n2, relu_a, a_scale, static_cast<float*>(workspace), rows_per_cta,
                 tc::tn_stages(n2)};
    {   // A^T in tensor memory (gemm_tn_wt.cu) is the default (89 vs 104 us at 128 x 128, 65 vs 79 us at 64 x 64);
        // GD_GEMM_TN=ring keeps the shared-memory operand kernel below
        const char* e = getenv("GD_GEMM_TN");
        const bool want_wt = !(e && strcmp(e, "ring") == 0);
        if (want_wt && tc::tn_wt_supported(t)) {
            int nparts = 0;
            const int rc = tc::launch_tn_wt(t, stream, &nparts);
            if (rc != GD_OK) return rc;
            const int64_t cnt = (int64_t)k1 * n2;
            GD_CUDA(launch_pdl(tn_reduce_kernel, (unsigned)ceil_div<int64_t>(cnt, 32), 256, 0, stream, (const float*)t.partial, nparts, cnt, c, 0, 0));
            GD_LAUNCH_CHECK();
            return GD_OK;
        }
    }
    const size_t smem = 1024 + (size_t)t.stages * (2 * tc::TILE_BYTES + 2 * n2 * 128) + tc::EPI_BYTES;
    GD_CUDA(cudaFuncSetAttribute(tc::gemm_tn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GD_CUDA(launch_pdl(tc::gemm_tn_tc_kernel, grid, tc::NUM_THREADS, smem, stream, t));
    GD_LAUNCH_CHECK();
    const int64_t count = (int64_t)k1 * n2;
    GD_CUDA(launch_pdl(tn_reduce_kernel, (unsigned)ceil_div<int64_t>(count, 32), 256, 0, stream, (const float*)t.partial, grid, count, c, 0, 0));
    GD_LAUNCH_CHECK();
    return GD_OK;
}
