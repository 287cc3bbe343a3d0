// Gathered-row GEMM with the WEIGHTS resident in tensor memory (the second tcgen05 design of the row GEMM).
//
//   out[r(i), :] = epi( pro(a[r(i), :]) . B )        m rows (gathered through `rows`), K, N <= 128, multiples of 32
//
// gemm_tc.cu keeps B (hi and lo, 128 KB at K = N = 128) resident in shared memory, which leaves room for a 2-deep ring of
// 128 x 32 A stages: its phases (load, split + store, MMA, epilogue) add up instead of overlapping, and with 3xTF32 the
// tensor core re-reads 24 KB of shared-memory operands per k-step (the L1 / shared-memory data pipe is that kernel's
// busiest unit, profiles/r2_gemm_ncu_full.md).  Here the product is computed TRANSPOSED,
//
//   D[n, rows] = W[n, k] . X^T[k, rows]              M = 128 output features (lanes), N = 64 rows of a tile (columns),
//
// with W_hi and W_lo as the A operand in TENSOR MEMORY (written once per CTA with tcgen05.st, 2 k columns), so
//   * shared memory holds only X tiles: a stage is a whole 64-row tile [x_hi | x_lo] x full K (64 KB at K = 128), the
//     ring is 3 tiles deep (6 at K = 64) and there is one producer -> MMA hand-shake per 64 rows instead of four per 128;
//   * the tensor core reads only the X operand from shared memory: 3xTF32 as  W_hi x [x_hi | x_lo]  (ONE N = 128
//     instruction into the adjacent main | correction accumulators) +  W_lo x x_hi  (N = 64): 6 KB per k-step, not 24;
//   * the accumulators (main + correction, 128 columns per tile) are double buffered in the other half of TMEM.
// The accumulator comes out transposed (thread = output feature, columns = rows): an epilogue warp turns 32 rows x 32
// features through a shared-memory tile and writes 128-byte runs of the (scattered) output rows with 128-bit stores
// (storing straight from the feature-per-lane layout - one 32-bit store per row - measured 10-35 % slower); the bit-packed
// ReLU mask of a row is a __ballot_sync over the 32 features a warp holds; the row scale is applied to the INPUT row by
// the producers (s (x . W) = (s x) . W).
//
// Covers row scale, ReLU prologue / epilogue, bias, ReLU bit mask out, gate bits in; fp32 gate rows and a bias combined
// with a row scale stay on gemm_tc.cu (rows_wt_supported()).
#include <type_traits>

#include "tc_common.cuh"

namespace gd {
namespace tc {

constexpr int WT_ROWS = 64;                    // rows per tile
constexpr int WT_PRODUCER_WARPS = 16;          // 512 threads: 8 per row (one 128-byte swizzle row per pass), 64 rows
constexpr int WT_EPI_WARPS = 8;                // warp % 4 = TMEM lane quarter = 32 output features; warp / 4 = 32-row half of the tile
constexpr int WT_MMA_WARP = WT_PRODUCER_WARPS + WT_EPI_WARPS;
constexpr int WT_THREADS = (WT_MMA_WARP + 1) * 32;
constexpr int WT_MAX_STAGES = 6;
constexpr int WT_ATOM = 16384;                 // one k-atom of a stage: [64 rows hi | 64 rows lo] x 128 B
constexpr int WT_EPI_BYTES = WT_EPI_WARPS * 32 * 32 * 4;      // per warp: 32 rows x 32 features transpose tile
// tiles of row loads in flight per producer thread: 8 float4 registers = 2 tiles at K = 128, 4 tiles at K <= 64
constexpr uint32_t WT_COL_D = 256;             // TMEM: W_hi | W_lo in columns [0, 2k), accumulators 2 x (main 64 | corr 64) from 256
constexpr int WT_TMEM_COLS = 512;

template <bool SCALE, bool MASK_OUT, bool GATE_BITS, int KCH>
__global__ void __launch_bounds__(WT_THREADS, 1) gemm_rows_wt_kernel(const Args g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // by offset: keeps the shared address space
    constexpr int kchunks = KCH;                                    // k-atoms (32 tf32 = one 128-byte swizzle row), g.k == 32 KCH
    constexpr int WT_PREFETCH = KCH <= 2 ? 4 : 2;
    const int stage_bytes = kchunks * WT_ATOM;
    const int STAGES = g.stages;
    float* epi_buf = reinterpret_cast<float*>(smem + STAGES * stage_bytes);
    __shared__ uint64_t full_bar[WT_MAX_STAGES], empty_bar[WT_MAX_STAGES], tfull_bar[2], tempty_bar[2], w_bar;
    __shared__ uint32_t tmem_base_smem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], WT_PRODUCER_WARPS); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], WT_EPI_WARPS); }
        mbar_init(&w_bar, WT_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == WT_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"(WT_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    // ---- W -> tensor memory, once: thread = output feature f (TMEM lane), 16 k values per tcgen05.st; rows >= n are zero.
    //      All 8 epilogue warps take part (the two warps of a lane quarter split the k range); only the MMA thread waits for
    //      it (w_bar), the producers start loading straight away.
    if (warp >= WT_PRODUCER_WARPS && warp < WT_MMA_WARP) {
        const int q = warp & 3, half = (warp - WT_PRODUCER_WARPS) >> 2;
        const int f = q * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        const bool vec = g.b_is_nk && (((uintptr_t)g.b) & 15) == 0;
        for (int c16 = half; c16 < g.k / 16; c16 += 2) {
            float w[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) w[e] = 0.f;
            if (f < g.n) {
                if (vec) {
#pragma unroll
                    for (int e4 = 0; e4 < 4; ++e4) {
                        const float4 v = __ldg(reinterpret_cast<const float4*>(g.b + (int64_t)f * g.k + c16 * 16) + e4);
                        w[4 * e4] = v.x; w[4 * e4 + 1] = v.y; w[4 * e4 + 2] = v.z; w[4 * e4 + 3] = v.w;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int kk = c16 * 16 + e;
                        w[e] = g.b_is_nk ? __ldg(g.b + (int64_t)f * g.k + kk) : __ldg(g.b + (int64_t)kk * g.n + f);
                    }
                }
            }
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                float h, l;
                split_tf32(w[e], h, l);
                hi[e] = __float_as_uint(h); lo[e] = __float_as_uint(l);
            }
            tmem_st16(t_lane + 16 * c16, hi);
            tmem_st16(t_lane + g.k + 16 * c16, lo);
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&w_bar);
    }

    if (warp < WT_PRODUCER_WARPS) {
        // ================================ producers ================================
        // thread = (row r0 = tid / 8 of the tile, 16-byte chunk j = tid % 8 of every k-atom): a warp loads 4 rows x 128 B per
        // pass and stores one conflict-free swizzled 128-byte row per quarter-warp.  Loads run WT_PREFETCH tiles ahead of the
        // shared-memory stores in registers; whole rows of the tile after those are pulled into L2 by one bulk prefetch each.
        const int j = tid & 7, r0 = tid >> 3;
        const int my_tiles = blockIdx.x < g.num_tiles ? (g.num_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
        auto row_of = [&](int t) -> int32_t {                        // t = this CTA's t-th tile
            if (t >= my_tiles) return -1;
            const int64_t i = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * WT_ROWS + r0;
            return i < g.m ? (g.rows ? __ldg(g.rows + i) : (int32_t)i) : -1;
        };
        float4 buf[WT_PREFETCH][KCH];
        float rsc[WT_PREFETCH];                                       // row scale: applied to the INPUT row (s (x . W) = (s x) . W), so the
                                                                      // transposed accumulator needs no per-row factor in the epilogue
        int32_t rid_pf = row_of(WT_PREFETCH);                         // row id of the tile whose loads are issued next
        auto issue = [&](float4 (&b)[KCH], float& sc, int32_t rid) {
            sc = 1.0f;
            if (SCALE && rid >= 0) sc = __ldg(g.out_scale + rid);
#pragma unroll
            for (int p = 0; p < KCH; ++p) {
                b[p] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (rid >= 0) b[p] = __ldg(reinterpret_cast<const float4*>(g.a + (int64_t)rid * g.lda + p * KC) + j);
            }
        };
        auto l2_prefetch = [&](int32_t rid) {
            if (j == 0 && rid >= 0)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(g.a + (int64_t)rid * g.lda), "r"(g.k * 4) : "memory");
        };
#pragma unroll
        for (int d = 0; d < WT_PREFETCH; ++d) issue(buf[d], rsc[d], row_of(d));
        l2_prefetch(rid_pf);
        uint32_t stage = 0, phase = 0;
        for (int t = 0; t < my_tiles; t += WT_PREFETCH) {
#pragma unroll
            for (int d = 0; d < WT_PREFETCH; ++d) {
                if (t + d < my_tiles) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* st = smem + stage * stage_bytes;
#pragma unroll
                    for (int p = 0; p < KCH; ++p) {
                        {
                            float4 x = buf[d][p];
                            if (g.relu_in) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
                            if (SCALE) { const float sc = rsc[d]; x.x *= sc; x.y *= sc; x.z *= sc; x.w *= sc; }
                            float4 hi, lo;
                            split4(x, hi, lo);
                            const uint32_t o = p * WT_ATOM + swz(r0, j);
                            *reinterpret_cast<float4*>(st + o) = hi;
                            *reinterpret_cast<float4*>(st + o + WT_ATOM / 2) = lo;
                        }
                    }
                    // no proxy fence here (it would wait for this thread's prefetched loads): the MMA thread fences after acquiring
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full_bar[stage]);
                    if (++stage == (uint32_t)STAGES) { stage = 0; phase ^= 1; }
                    const int32_t rid = rid_pf;                       // loads of tile t + d + WT_PREFETCH
                    rid_pf = row_of(t + d + WT_PREFETCH + 1);
                    issue(buf[d], rsc[d], rid);
                    l2_prefetch(rid_pf);
                }
            }
        }
    } else if (warp == WT_MMA_WARP) {
        // ================================ MMA issuer ================================
        if (elect_one()) {
            const uint32_t idesc128 = make_idesc(2 * WT_ROWS), idesc64 = make_idesc(WT_ROWS);
            const uint64_t x0 = make_desc(smem_u32(smem));
            const uint32_t w_hi = tmem_base, w_lo = tmem_base + g.k;
            uint32_t stage = 0, phase = 0;
            int it = 0;
            mbar_wait(&w_bar, 0);                                     // W_hi / W_lo are in tensor memory
            tc_fence_after();
            for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
                const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                mbar_wait(&full_bar[stage], phase);
                fence_proxy_async();                                  // the producers' generic stores -> async proxy
                tc_fence_after();
                const uint32_t d = tmem_base + WT_COL_D + acc * 128;  // main | corr
                const uint64_t xs = desc_advance(x0, stage * stage_bytes);
                for (int p = 0; p < kchunks; ++p) {
#pragma unroll
                    for (int ks = 0; ks < KC / 8; ++ks) {
                        const uint64_t b = xs + (uint64_t)(p * (WT_ATOM >> 4) + 2 * ks);
                        const uint32_t ka = p * KC + ks * 8;
                        umma_tf32_tmem_a(d, w_hi + ka, b, idesc128, (p | ks) != 0);       // main | corr (+)= W_hi x [x_hi | x_lo]
                        umma_tf32_tmem_a(d + WT_ROWS, w_lo + ka, b, idesc64, 1);          // corr += W_lo x x_hi
                    }
                }
                umma_commit(&empty_bar[stage]);                        // frees the stage when the MMAs retire
                umma_commit(&tfull_bar[acc]);                          // accumulators of this tile complete
                if (++stage == (uint32_t)STAGES) { stage = 0; phase ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // ================================ epilogue ================================
        // 8 warps: q = warp % 4 = TMEM lane quarter (32 output features, thread = feature), h = 32-row half of the tile.
        // accumulators -> registers -> [32 rows x 32 features] shared tile -> 128-byte runs of the output rows (8 lanes x 16 B
        // per row, 4 rows per store).  Row ids / gate words of the NEXT tile are loaded before this one is awaited.
        const int ew = warp - WT_PRODUCER_WARPS;
        const int q = warp & 3, h = ew >> 2;
        const bool active = q * 32 < g.n;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        float* tbuf = epi_buf + ew * (32 * 32);
        const int nw = g.n >> 5;
        int32_t rid_n = -1; uint32_t gw_n = 0u;
        auto load_meta = [&](int tile) {                              // this lane's row (32 h + lane) of `tile`
            rid_n = -1; gw_n = 0u;
            if (tile < g.num_tiles && active) {
                const int64_t gi = (int64_t)tile * WT_ROWS + 32 * h + lane;
                rid_n = gi < g.m ? (g.rows ? __ldg(g.rows + gi) : (int32_t)gi) : -1;
                if (GATE_BITS && rid_n >= 0) gw_n = __ldg(g.gate_bits + (int64_t)rid_n * nw + q);
            }
        };
        load_meta(blockIdx.x);
        // bias (per output feature = per lane) and the ReLU epilogue are run-time options: one add / one max per element.  With a
        // row scale the bias is added AFTER scaling in the contract (out = scale * (a . B) + bias is NOT what gemm_tc.cu computes:
        // it computes scale * (a . B + bias)), so a bias together with a row scale stays on gemm_tc.cu (rows_wt_supported).
        const float bias_f = (g.bias && active && q * 32 + lane < g.n) ? __ldg(g.bias + q * 32 + lane) : 0.f;
        const bool relu_out = g.relu_out != 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
            const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
            const int32_t rid = rid_n;
            const uint32_t gw = gw_n;
            load_meta(tile + gridDim.x);
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            if (active) {
                const uint32_t tD = t_lane + WT_COL_D + acc * 128 + 32 * h;
                uint32_t posword = 0;
#pragma unroll
                for (int c = 0; c < 2; ++c) {                         // 16 rows at a time (register budget of an 800-thread CTA)
                    uint32_t pm[16], pc[16];
                    tmem_ld16_nowait(tD + 16 * c, pm);
                    tmem_ld16_nowait(tD + WT_ROWS + 16 * c, pc);
                    tmem_wait_ld();
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int jj = 16 * c + e;
                        float x = __uint_as_float(pm[e]) + __uint_as_float(pc[e]) + bias_f;
                        if (relu_out) x = fmaxf(x, 0.f);
                        if (MASK_OUT) {
                            const uint32_t v = __ballot_sync(0xffffffffu, x > 0.f);    // the 32 features of row jj = one mask word
                            if (lane == jj) posword = v;
                        }
                        tbuf[jj * 32 + lane] = x;
                    }
                }
                // the accumulator half is in registers / shared memory: hand the buffer back before the global stores
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                // gate bits are applied after the transposition (a lane then holds four features of ONE row: the row's gate word
                // comes with one shuffle; transposing the 32 x 32 bit block with ballots cost 10 us per 200 k rows); the mask word
                // of a gated row is the AND of both
                if (MASK_OUT) { if (rid >= 0) g.relu_mask_out[(int64_t)rid * nw + q] = GATE_BITS ? (posword & gw) : posword; }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int jj = (lane >> 3) + 4 * i;
                    float4 o = lds128(tbuf + jj * 32 + (lane & 7) * 4);
                    const int32_t rr = __shfl_sync(0xffffffffu, rid, jj);
                    if (GATE_BITS) {
                        const uint32_t nib = __shfl_sync(0xffffffffu, gw, jj) >> ((lane & 7) * 4);
                        if (!(nib & 1u)) o.x = 0.f;
                        if (!(nib & 2u)) o.y = 0.f;
                        if (!(nib & 4u)) o.z = 0.f;
                        if (!(nib & 8u)) o.w = 0.f;
                    }
                    if (rr >= 0) *reinterpret_cast<float4*>(g.out + (int64_t)rr * g.ldo + q * 32 + (lane & 7) * 4) = o;
                }
                __syncwarp();
            } else {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            }
        }
    }
    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == WT_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(WT_TMEM_COLS));
    }
}

constexpr size_t WT_SMEM_LIMIT = 227 * 1024 - 2048;   // static barriers + alignment slack
static int wt_stages(int k) {
    const size_t stage = (size_t)(k / KC) * WT_ATOM;
    return (int)std::min<size_t>(WT_MAX_STAGES, (WT_SMEM_LIMIT - 1024 - WT_EPI_BYTES) / stage);
}

bool rows_wt_supported(const Args& g) {
    if (g.k <= 0 || g.k > 128 || g.k % KC != 0 || g.n <= 0 || g.n > 128 || g.n % 32 != 0) return false;
    if (g.gate) return false;                                          // fp32 gate rows: gemm_tc.cu
    if (g.bias && g.out_scale) return false;                           // bias is added BEFORE the row scale there; here the scale sits on the input row
    if (g.lda % 4 != 0 || g.ldo % 4 != 0) return false;
    return wt_stages(g.k) >= 2;
}

int launch_rows_wt(const Args& g_in, cudaStream_t stream) {
    Args g = g_in;
    g.num_tiles = (int)ceil_div<int64_t>(g.m, WT_ROWS);
    g.stages = wt_stages(g.k);
    const size_t smem = 1024 + (size_t)g.stages * (g.k / KC) * WT_ATOM + WT_EPI_BYTES;
    const int grid = std::min(g.num_tiles, kNumSMs);
    auto launch = [&](auto kern) -> int {
        GD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GD_CUDA(launch_pdl(kern, grid, WT_THREADS, smem, stream, g));
        GD_LAUNCH_CHECK();
        return GD_OK;
    };
    const int v = (g.out_scale ? 1 : 0) | (g.relu_mask_out ? 2 : 0) | (g.gate_bits ? 4 : 0);
    auto by_k = [&](auto sc, auto mk, auto gt) -> int {
        constexpr bool S = decltype(sc)::value, M = decltype(mk)::value, G = decltype(gt)::value;
        switch (g.k / KC) {
            case 1: return launch(gemm_rows_wt_kernel<S, M, G, 1>);
            case 2: return launch(gemm_rows_wt_kernel<S, M, G, 2>);
            case 3: return launch(gemm_rows_wt_kernel<S, M, G, 3>);
            default: return launch(gemm_rows_wt_kernel<S, M, G, 4>);
        }
    };
    using T = std::true_type; using F = std::false_type;
    switch (v) {
        case 0: return by_k(F{}, F{}, F{});
        case 1: return by_k(T{}, F{}, F{});
        case 2: return by_k(F{}, T{}, F{});
        case 3: return by_k(T{}, T{}, F{});
        case 4: return by_k(F{}, F{}, T{});
        case 5: return by_k(T{}, F{}, T{});
        case 6: return by_k(F{}, T{}, T{});
        default: return by_k(T{}, T{}, T{});
    }
}

}  // namespace tc
}  // namespace gd
