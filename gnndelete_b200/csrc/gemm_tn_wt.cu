// Weight-gradient contraction with the A^T operand in TENSOR MEMORY (the second tcgen05 design of gemm_tn_tc_kernel).
//
//   c[k1, n2] = sum_i a_scale[r(i)] * pro(a[r(i), :k1])^T (x) g[r(i), :n2]        (dW_del = x[S]^T . dOut[S])
//
// The contraction index is the ROW.  gemm_tc.cu stages A^T and G^T (hi and lo each) in shared memory: 64 KB per 32 rows
// at k1 = n2 = 128, so a stage is 32 rows, the ring 3 deep, and per k-step the tensor core re-reads 20 KB of operands -
// its shared-memory data pipe is as busy at the HBM roofline as the memory system (0.23-0.31 of the roofline measured).
// Here the tensor core reads A^T from TENSOR MEMORY: the rows of `a` pass through a raw row-major shared-memory tile (128-bit
// loads and stores, no hi / lo duplication), are read back per feature (lane = feature = TMEM lane), split hi / lo in
// registers and written with tcgen05.st as the A operand [M = k1 lanes x K = 64 rows]; only G is staged in operand form
// (MN-major, [g_hi | g_lo] adjacent = ONE N = 2 n2 instruction).  A stage is 64 rows, and per k-step the tensor core reads
// 12 KB from shared memory.  (Loading the columns straight from global memory - one 32-bit load per row and lane - kept too
// few bytes in flight per register: 179 us against 105 us for gemm_tc.cu on the 128 x 128 case.)
//   warps 0-7   A producers   (warp % 4 = TMEM lane quarter = 32 features, warp / 4 = 32-row half of the stage)
//   warps 8-15  G producers   (16-byte chunks into the SWIZZLE_128B_BASE32B MN-major tiles, as gemm_tc.cu)
//   warp  16    MMA issue     (A from TMEM, 2 MMAs per k-step: main | corr (+)= a_hi x [g_hi | g_lo], corr += a_lo x g_hi)
//   warps 17-20 epilogue      (every TW_FLUSH stages the accumulators are added to the CTA's partial in global memory)
// The A loaders keep one stage of rows in flight in registers, the G loaders rely on the ring depth.  (Bulk L2 prefetches of
// the rows two stages ahead, which help the row GEMM, cost 8-25 % here and in gemm_tc.cu's weight-gradient kernel.)
// Partials are reduced in CTA order by tn_reduce_kernel (deterministic, no atomics).
#include "tc_common.cuh"

namespace gd {
namespace tc {

constexpr int TW_ROWS = 64;                      // rows per stage = 8 k-steps
constexpr int TW_A_WARPS = 8, TW_G_WARPS = 8, TW_EPI_WARPS = 4;
constexpr int TW_MMA_WARP = TW_A_WARPS + TW_G_WARPS;
constexpr int TW_THREADS = (TW_MMA_WARP + 1 + TW_EPI_WARPS) * 32;     // 672
constexpr int TW_FLUSH_MAX = 8;                  // stages per accumulator flush: 4 (256 rows = 32 k-steps, the chain length of gemm_tc.cu) with two
                                                 // accumulator buffers; 8 when n2 > 64 leaves room for one buffer only (the MMAs wait for its drain)
constexpr int TW_MAX_STAGES = 4;
constexpr int TW_ATOM_COL = 4096;                // bytes of one 32-feature atom column of a 32-row block: 8 k-atoms (4 rows x 128 B)
constexpr int TW_EPI_LD = 36;
constexpr int TW_EPI_BYTES = TW_EPI_WARPS * 32 * TW_EPI_LD * 4;
constexpr uint32_t TW_COL_A = 256;               // TMEM: accumulators in columns [0, 256), two A^T stages (hi 64 | lo 64) from 256
constexpr int TW_TMEM_COLS = 512;

__global__ void __launch_bounds__(TW_THREADS, 1) gemm_tn_wt_kernel(const TnArgs t) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // by offset: keeps the shared address space
    const int g_tile = t.n2 * 128;                                  // bytes of one [n2 x 128 B] tile = 32 rows of g_hi (or g_lo)
    const int stage_bytes = 4 * g_tile;                             // 2 blocks of 32 rows x [g_hi | g_lo]
    const int GS = t.stages;
    float* epi_buf = reinterpret_cast<float*>(smem + GS * stage_bytes);
    __shared__ uint64_t gfull[TW_MAX_STAGES], gempty[TW_MAX_STAGES], afull[2], aempty[2], tfull[2], tempty[2];
    __shared__ uint32_t tmem_base_smem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nbuf = t.n2 <= 64 ? 2 : 1;                            // accumulator buffers (main | corr = 2 n2 columns each)
    const int TW_FLUSH = nbuf == 2 ? 4 : TW_FLUSH_MAX;
    if (tid == 0) {
        for (int s = 0; s < TW_MAX_STAGES; ++s) { mbar_init(&gfull[s], TW_G_WARPS); mbar_init(&gempty[s], 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&afull[s], TW_A_WARPS); mbar_init(&aempty[s], 1);
            mbar_init(&tfull[s], 1); mbar_init(&tempty[s], TW_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TW_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"(TW_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    const int64_t r_beg = (int64_t)blockIdx.x * t.rows_per_cta;
    const int64_t r_end = min(t.m, r_beg + t.rows_per_cta);
    const int nstages = r_end > r_beg ? (int)((r_end - r_beg + TW_ROWS - 1) / TW_ROWS) : 0;
    const int ngroups = (nstages + TW_FLUSH - 1) / TW_FLUSH;
    auto row_id = [&](int64_t i) -> int32_t { return (i >= r_beg && i < r_end) ? (t.rows ? __ldg(t.rows + i) : (int32_t)i) : -1; };

    if (warp < TW_A_WARPS) {
        // ------------------------------ A producers: rows of `a` -> A^T in tensor memory ------------------------------
        // Two phases per stage.  (1) loader mapping (8 lanes x 16 B per row: coalesced 128-bit loads kept one stage ahead in
        // registers) -> raw row-major tile in shared memory [64 rows][k1], ReLU / row scale applied on the way.  (2) feature
        // mapping: thread = feature f = TMEM lane reads its column of the tile (consecutive lanes = consecutive words:
        // conflict free), splits hi / lo and writes 16 rows per tcgen05.st.  The raw tile is double buffered; ONE named
        // barrier of the 8 A warps per stage separates the phases (a warp reaches the barrier of stage s + 1 only after its
        // reads of stage s, so the buffer of stage s is free again when stage s + 2 is stored).
        const int q = warp & 3, half = warp >> 2;
        const int f = q * 32 + lane;                                // feature = TMEM lane
        const bool active = q * 32 < t.k1;                          // accumulator rows >= k1 are never read: nothing to write for them
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + TW_COL_A + 32 * half;
        const int lr = tid >> 3, cj = tid & 7;                      // loader mapping: rows lr and lr + 32 of the stage, chunks cj + 8 i
        const int a4 = t.k1 >> 2;
        float* raw = reinterpret_cast<float*>(smem + GS * stage_bytes + TW_EPI_BYTES);    // 2 x [64][k1] floats
        float4 xa[2][4];
        float xs[2];
        int32_t rid_nx[2];                                          // row ids of the stage whose loads are issued next (loaded a stage
                                                                    // earlier: an issue is never two dependent latencies)
        auto fetch_rids = [&](int s) {
#pragma unroll
            for (int b = 0; b < 2; ++b) rid_nx[b] = s < nstages ? row_id(r_beg + (int64_t)s * TW_ROWS + 32 * b + lr) : -1;
        };
        auto issue = [&](int s) {
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int32_t r = rid_nx[b];
                xs[b] = (t.a_scale && r >= 0) ? __ldg(t.a_scale + r) : 1.0f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    xa[b][i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r >= 0 && cj + 8 * i < a4) xa[b][i] = __ldg(reinterpret_cast<const float4*>(t.a + (int64_t)r * t.lda) + cj + 8 * i);
                }
            }
            fetch_rids(s + 1);
        };
        fetch_rids(0);
        issue(0);
        for (int s = 0; s < nstages; ++s) {
            const uint32_t ta = s & 1;
            float* rt = raw + (s & 1) * (TW_ROWS * t.k1);
            // ---- phase 1: registers -> raw tile
#pragma unroll
            for (int b = 0; b < 2; ++b) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (cj + 8 * i < a4) {
                        float4 x = xa[b][i];
                        if (t.relu_a) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
                        if (t.a_scale) { const float sc = xs[b]; x.x *= sc; x.y *= sc; x.z *= sc; x.w *= sc; }
                        *reinterpret_cast<float4*>(rt + (32 * b + lr) * t.k1 + (cj + 8 * i) * 4) = x;
                    }
                }
            }
            issue(s + 1);                                            // next stage's rows in flight during phase 2
            asm volatile("bar.sync 1, %0;" ::"r"(TW_A_WARPS * 32) : "memory");
            // ---- phase 2: raw tile column -> tensor memory (every A warp stays in step with the TMEM stage ring)
            mbar_wait(&aempty[ta], ((uint32_t)(s >> 1) & 1u) ^ 1u);   // the MMAs of stage s - 2 have read this TMEM stage
            tc_fence_after();
            if (active) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const float x = f < t.k1 ? rt[(32 * half + 16 * c + e) * t.k1 + f] : 0.f;
                        float h, l;
                        split_tf32(x, h, l);
                        hi[e] = __float_as_uint(h); lo[e] = __float_as_uint(l);
                    }
                    tmem_st16(t_lane + ta * 128 + 16 * c, hi);
                    tmem_st16(t_lane + ta * 128 + 64 + 16 * c, lo);
                }
                tmem_wait_st();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&afull[ta]);
        }
    } else if (warp < TW_MMA_WARP) {
        // ------------------------------ G producers: MN-major stage fill ------------------------------
        // row-major G[rows, n2] IS the MN-major operand layout (features contiguous, one k-row per matrix row): every lane
        // copies 16-byte chunks (SWIZZLE_128B_BASE32B placement); 8 lanes cover the 128 B of one k-row of one 32-feature atom,
        // a warp covers 4 rows, the 8 warps the 32 rows of a block, two blocks per stage
        const int gw = warp - TW_A_WARPS;
        const int rr = gw * 4 + (lane >> 3), cj = lane & 7;
        const int g4 = t.n2 >> 2;
        const uint32_t row_off = (uint32_t)((rr >> 2) * 512 + (rr & 3) * 128 + ((((cj >> 1) ^ (rr & 3)) << 5) | ((cj & 1) << 4)));
        uint32_t stage = 0, phase = 0;
        int32_t nrid0 = row_id(r_beg + rr), nrid1 = row_id(r_beg + rr + 32);       // row ids one stage ahead of the loads
        for (int s = 0; s < nstages; ++s) {
            const int32_t rid0 = nrid0, rid1 = nrid1;
            {
                const int64_t i1 = r_beg + (int64_t)(s + 1) * TW_ROWS + rr;
                nrid0 = row_id(i1); nrid1 = row_id(i1 + 32);
            }
            float4 x[2][4];
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int32_t r = b ? rid1 : rid0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    x[b][i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r >= 0 && cj + 8 * i < g4) x[b][i] = __ldg(reinterpret_cast<const float4*>(t.g + (int64_t)r * t.ldg) + cj + 8 * i);
                }
            }
            mbar_wait(&gempty[stage], phase ^ 1);
            uint8_t* st = smem + stage * stage_bytes + row_off;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {                        // atom i = features 32 i .. 32 i + 31, 4 KB apart
                    if (cj + 8 * i < g4) {
                        float4 hi, lo;
                        split4(x[b][i], hi, lo);
                        *reinterpret_cast<float4*>(st + b * 2 * g_tile + i * TW_ATOM_COL) = hi;
                        *reinterpret_cast<float4*>(st + b * 2 * g_tile + g_tile + i * TW_ATOM_COL) = lo;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&gfull[stage]);               // (no proxy fence on this side: the MMA thread fences after acquiring)
            if (++stage == (uint32_t)GS) { stage = 0; phase ^= 1; }
        }
    } else if (warp == TW_MMA_WARP) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc_mn(t.n2) & ~(1u << 15), idesc2 = make_idesc_mn(2 * t.n2) & ~(1u << 15);   // B MN-major, A from TMEM
            const uint64_t g0 = make_desc_mn(smem_u32(smem), TW_ATOM_COL, 512);
            uint32_t stage = 0, phase = 0;
            for (int s = 0; s < nstages; ++s) {
                const int grp = s / TW_FLUSH;
                const uint32_t acc = grp % nbuf, use = grp / nbuf;
                const bool first = (s % TW_FLUSH) == 0;
                const uint32_t ta = s & 1;
                if (first) {
                    mbar_wait(&tempty[acc], (use & 1u) ^ 1u);        // the epilogue has drained this accumulator buffer
                    tc_fence_after();
                }
                mbar_wait(&gfull[stage], phase);
                mbar_wait(&afull[ta], (uint32_t)(s >> 1) & 1u);
                fence_proxy_async();                                  // the G producers' generic stores -> async proxy
                tc_fence_after();
                const uint32_t d = tmem_base + acc * 2 * t.n2;
                const uint32_t a_hi = tmem_base + TW_COL_A + ta * 128, a_lo = a_hi + 64;
                const uint64_t gs = desc_advance(g0, stage * stage_bytes);
#pragma unroll
                for (int ks = 0; ks < TW_ROWS / 8; ++ks) {           // 8 rows per k-step: block ks / 4, two 4-row k-atoms (1 KB) inside it
                    const uint64_t b = desc_advance(gs, (ks >> 2) * 2 * g_tile + (ks & 3) * 1024);
                    umma_tf32_tmem_a(d, a_hi + 8 * ks, b, idesc2, !(first && ks == 0));     // main | corr (+)= a_hi x [g_hi | g_lo]
                    umma_tf32_tmem_a(d + t.n2, a_lo + 8 * ks, b, idesc, 1);                 // corr += a_lo x g_hi
                }
                umma_commit(&gempty[stage]);
                umma_commit(&aempty[ta]);
                if ((s % TW_FLUSH) == TW_FLUSH - 1 || s == nstages - 1) umma_commit(&tfull[acc]);
                if (++stage == (uint32_t)GS) { stage = 0; phase ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------ epilogue: add a flush group to the CTA's partial ------------------------------
        const int q = warp & 3;
        float* tbuf = epi_buf + (warp - TW_MMA_WARP - 1) * (32 * TW_EPI_LD);
        const int cl = lane & 7, rl = lane >> 3;
        float* part = t.partial + (int64_t)blockIdx.x * t.k1 * t.n2;
        for (int grp = 0; grp < ngroups; ++grp) {
            const uint32_t acc = grp % nbuf, use = grp / nbuf;
            mbar_wait(&tfull[acc], use & 1u);
            tc_fence_after();
            for (int c0 = 0; c0 < t.n2; c0 += 32) {
                const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 2 * t.n2 + c0;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {                     // 16 columns at a time (register budget of a 672-thread CTA)
                    uint32_t v[16], vc[16];
                    tmem_ld16_nowait(t0 + 16 * hh, v);
                    tmem_ld16_nowait(t0 + t.n2 + 16 * hh, vc);
                    tmem_wait_ld();
#pragma unroll
                    for (int e = 0; e < 16; e += 4)
                        *reinterpret_cast<float4*>(tbuf + lane * TW_EPI_LD + 16 * hh + e) =
                            make_float4(__uint_as_float(v[e]) + __uint_as_float(vc[e]), __uint_as_float(v[e + 1]) + __uint_as_float(vc[e + 1]),
                                        __uint_as_float(v[e + 2]) + __uint_as_float(vc[e + 2]), __uint_as_float(v[e + 3]) + __uint_as_float(vc[e + 3]));
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = q * 32 + i * 4 + rl;
                    if (row < t.k1) {
                        float4 o = *reinterpret_cast<const float4*>(tbuf + (i * 4 + rl) * TW_EPI_LD + cl * 4);
                        // every element of the CTA's partial is updated by this one thread, group after group: a vector reduction
                        // is deterministic here and, unlike a read-modify-write, does not wait for the old value (with one
                        // accumulator buffer at n2 = 128 the MMAs wait for this drain)
                        float* dst = part + (int64_t)row * t.n2 + c0 + cl * 4;
                        if (grp > 0) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
                        else *reinterpret_cast<float4*>(dst) = o;
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
        if (ngroups == 0) {                                         // CTA without rows: its partial is zero
            const int f = q * 32 + lane;
            if (f < t.k1) for (int c = 0; c < t.n2; ++c) part[(int64_t)f * t.n2 + c] = 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TW_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TW_TMEM_COLS));
    }
}

constexpr size_t TW_SMEM_LIMIT = 227 * 1024 - 2048;
static size_t tw_raw_bytes(int k1) { return (size_t)2 * TW_ROWS * k1 * sizeof(float); }      // double-buffered raw A tile
static int tw_stages(int k1, int n2) {
    return (int)std::min<size_t>(TW_MAX_STAGES, (TW_SMEM_LIMIT - 1024 - TW_EPI_BYTES - tw_raw_bytes(k1)) / ((size_t)n2 * 512));
}

bool tn_wt_supported(const TnArgs& t) {
    if (t.k1 <= 0 || t.k1 > 128 || t.k1 % 4 != 0 || t.n2 <= 0 || t.n2 > 128 || t.n2 % 32 != 0) return false;
    if (t.lda % 4 != 0 || t.ldg % 4 != 0) return false;
    return tw_stages(t.k1, t.n2) >= 2;
}

int launch_tn_wt(const TnArgs& t_in, cudaStream_t stream, int* nparts) {
    TnArgs t = t_in;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(kNumSMs, ceil_div<int64_t>(t.m, TW_ROWS * 4)));
    t.rows_per_cta = ceil_div<int64_t>(ceil_div<int64_t>(t.m, grid), TW_ROWS) * TW_ROWS;
    t.stages = tw_stages(t.k1, t.n2);
    const size_t smem = 1024 + (size_t)t.stages * t.n2 * 512 + TW_EPI_BYTES + tw_raw_bytes(t.k1);
    GD_CUDA(cudaFuncSetAttribute(gemm_tn_wt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GD_CUDA(launch_pdl(gemm_tn_wt_kernel, grid, TW_THREADS, smem, stream, t));
    GD_LAUNCH_CHECK();
    *nparts = grid;
    return GD_OK;
}

}  // namespace tc
}  // namespace gd
