// Dense-block Neighbourhood-Influence loss of train_fullbatch (gnndelete.py:163-193, 239-241) on the 5th-generation
// tensor cores: the one genuinely compute-bound contraction of the path (S2 x S2 x 64, 158 M node pairs at the Cora
// shape; the fp32 CUDA-core kernel in dense_ni.cu needs 4.7 ms per epoch for it, 30x the rest of that epoch).
//
//   loss_l = mean_{(i,j) in M} ( sigmoid(<z_i, z_j>) - sigmoid(logits_ori[i, j]) )^2,   dz_i = sum_j c_ij z_j
//
// One CTA owns 128 rows I of z_S and sweeps the 64-row column blocks J (flash-attention shaped, nothing N x N is formed):
//   MMA1   P[128 x 64]   = Z_I . Z_J^T            tcgen05.mma kind::tf32, M = 128, N = 64, K = 64, accumulators in TMEM
//   epilogue (thread = row i, 16 columns): tcgen05.ld P, s = sigmoid(P), residual against the packed target tile,
//            coefficient c = (s - t) s (1 - s), loss partial; c is split hi / lo and written back over P (tcgen05.st)
//   MMA2   dZ_I[128 x 64] += C[128 x 64] . Z_J     A operand from tensor memory, accumulated in TMEM over the whole sweep
// Both contractions keep fp32 accuracy with the 3xTF32 operand split of gemm_tc.cu (main + correction accumulators);
// hi and lo of the B operand sit next to each other along N, so a k-step is two MMAs (N = 128 and N = 64), not three.
// The logit tile is double buffered in TMEM: MMA1 of block t + 1 is issued before the epilogue of block t starts.
// Every operand is K-major (the layout the row GEMM has exercised since round 1): Z_J is staged twice, as [j][d] for
// MMA1 and transposed as [d][j] for MMA2 (4-byte conflict-free stores, lanes = j).
//
// The target block sigmoid(logits_ori[S][:, S]) is PACKED once per plan (gd_dense_ni_tc_pack_target): per (128 x 64)
// tile, 16-byte chunk c of row i at float4 index c * 128 + i, so a warp reads 512 contiguous bytes per load; excluded
// pairs (Df, both orders), the diagonal and the padding carry the sentinel -1 (a sigmoid is never negative), so the
// kernel reads no bitmap.  Every unordered pair is visited from both sides: a CTA writes only its own rows of dz -
// deterministic, no atomics.  Small S: the J sweep is split over `jsplit` CTAs per row block and a second kernel
// adds the parts in order.
#include <type_traits>

#include "tc_common.cuh"

namespace gd {
namespace tc {

constexpr int NI_BI = 128;                 // rows of z_S per CTA (UMMA M)
constexpr int NI_BJ = 64;                  // column block: UMMA N of the logit tile, K of the gradient contraction
constexpr int NI_D = 64;                   // embedding width (out_dim of the reference's models)
constexpr int NI_WORKER_WARPS = 16;         // 4 per TMEM lane quarter, 16 logit columns per thread
constexpr int NI_THREADS = (NI_WORKER_WARPS + 1) * 32;   // + the MMA-issuing warp
constexpr int NI_TILE = NI_BI * NI_BJ;     // floats of one packed target tile
// shared-memory map (bytes; every operand tile 1024-aligned; "atom" = 32 k values = one 128-byte swizzle row)
constexpr int NI_ZI_HI = 0;                        // Z_I   [128 rows i][k = d]   2 atoms x 16 KB
constexpr int NI_ZI_LO = NI_ZI_HI + 2 * 16384;
constexpr int NI_ZJ_HI = NI_ZI_LO + 2 * 16384;     // Z_J   2 atoms x [64 rows j hi | 64 rows j lo][k = d] (16 KB each)   B of MMA1
constexpr int NI_ZT = NI_ZJ_HI + 2 * 16384;        // Z_J^T 2 buffers x 2 atoms x [64 rows d hi | 64 rows d lo][k = j]     B of MMA2
constexpr int NI_ZT_BUF = 2 * 16384;
constexpr int NI_SMEM = NI_ZT + 2 * NI_ZT_BUF;     // 163840
constexpr int NI_TMEM_COLS = 512;                  // P[0] main | corr, P[1] main | corr, dZ main | corr (64 columns each)
constexpr uint32_t NI_COL_P = 0, NI_COL_DZ = 256;

struct NiArgs {
    const float* zs; int64_t ldz; int64_t n_s;
    const float* packed; int n_ib, n_jb;
    float scale2;                                   // 2 * weight / |M|
    float* dz_out; int64_t lddz;                    // jsplit == 1: dzs [n_s][lddz];  else parts [jsplit][n_ib * 128][64]
    float* partial;                                 // [gridDim.x] sums of squared residuals over i > j
    int jsplit;
};

// The coefficient tile is split for the 3xTF32 contraction WITHOUT rounding instructions: kind::tf32 reads the upper 19
// bits of a 32-bit operand, so the raw fp32 value serves as "hi" (= c truncated to tf32) and lo = c - trunc(c) is exact
// (<= 13 significant bits, of which the tensor core keeps 11: error 2^-21 relative).  2 ALU operations per element
// instead of 5 on the kernel's busiest pipe.  NI_TRUNC_SPLIT = false restores the rounded split of gemm_tc.cu.
constexpr bool NI_TRUNC_SPLIT = true;
__device__ __forceinline__ void split_coef(float c, float& hi, float& lo) {
    if (NI_TRUNC_SPLIT) {
        hi = c;
        lo = c - __uint_as_float(__float_as_uint(c) & 0xffffe000u);
    } else {
        split_tf32(c, hi, lo);
    }
}

// The logit tile is computed as P' = (-log2(e) Z_I) . Z_J^T (the factor is folded into the Z_I operand when it is staged),
// so sigmoid(p) = 1 / (1 + 2^P') costs two MUFU operations and one add (ex2.approx / rcp.approx: 2^-22 relative each).
// Saturates cleanly: 2^x = inf -> 1 / inf = 0, 2^x = 0 -> 1.
constexpr float NI_NEG_LOG2E = -1.4426950408889634f;
__device__ __forceinline__ float sigmoid_from_scaled(float x) {
    float e, s;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(1.0f + e));
    return s;
}

// Software pipeline of one CTA.  Warps 0-15 (workers) stage the operands and run the epilogue; warp 16 issues the MMAs
// (one elected lane) and prefetches the target tiles into L2.  The issuing warp is separate because the generic->async
// proxy fence that must precede a tcgen05.mma on freshly written shared memory compiles to MEMBAR.ALL.CTA, which waits
// for the executing thread's own outstanding global loads - on a worker it exposed the latency of the prefetched target
// tile and Z_J rows in every block (0.84 ms for the Cora block); warp 16 has no loads in flight.  The tensor pipe
// executes in issue order:
//   block t:  workers wait P[t] (also frees the Z_J buffer), stage Z_J(t+1)      | sync A |  warp 16: MMA1(t+1) -> P[(t+1)&1]
//             workers: TMEM -> coefficients; wait MMA2(t-1); store C(t), Z_J^T(t) | sync B |  warp 16: MMA2(t) -> dZ
//
// The coefficient tile never touches shared memory: the epilogue writes C (hi = the raw fp32 value, lo = c - trunc(c)) with
// tcgen05.st IN PLACE over the logit tile it has just read (P main -> C hi, P corr -> C lo) and the gradient contraction
// takes its A operand from tensor memory.  Against a C tile in shared memory (the first version, 0.80 ms) that removes
// 64 KB of stores and 96 KB of A-operand reads per block from the shared-memory data pipe - the kernel's busiest unit,
// profiles/r2_dense_ni_tc_ncu_full.md - and the wait on the previous contraction: Z_J^T is double buffered instead.
__global__ void __launch_bounds__(NI_THREADS, 1) dense_ni_tc_kernel(const NiArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // by offset: keeps the shared address space
    __shared__ uint64_t bar_p, bar_end;
    __shared__ uint32_t tmem_base_smem;
    __shared__ float red[NI_WORKER_WARPS];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ib = blockIdx.x / a.jsplit, split = blockIdx.x - ib * a.jsplit;
    const int per = (a.n_jb + a.jsplit - 1) / a.jsplit;
    const int jb0 = split * per, jb1 = min(a.n_jb, jb0 + per);
    const int T = max(jb1 - jb0, 0);
    const bool worker = warp < NI_WORKER_WARPS;
    const float* tile0 = a.packed + ((int64_t)ib * a.n_jb + jb0) * NI_TILE;       // this CTA's target tiles are contiguous

    if (tid == 0) {
        mbar_init(&bar_p, 1);
        mbar_init(&bar_end, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (!worker) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"(NI_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // Z_J staging role: thread = (row j = 32 (warp % 2) + lane, 32 bytes = chunks 2 (warp / 2), + 1 of the row)
    const int sj = (warp & 1) * 32 + lane, sc0 = ((warp >> 1) & 7) * 2;
    auto load_zj = [&](int jb, float4 (&z)[2]) {
        const int64_t gj = (int64_t)jb * NI_BJ + sj;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            z[e] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (jb < jb1 && gj < a.n_s) z[e] = __ldg(reinterpret_cast<const float4*>(a.zs + gj * a.ldz) + sc0 + e);
        }
    };
    auto stage_k = [&](const float4 (&z)[2]) {                       // K-major [j][d]: operand B of the logit tile
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            float4 hi, lo;
            split4(z[e], hi, lo);
            const int c = sc0 + e;
            const uint32_t o = (c >> 3) * 16384 + swz(sj, c & 7);     // atom: [64 rows hi | 64 rows lo], one N = 128 operand
            *reinterpret_cast<float4*>(smem + NI_ZJ_HI + o) = hi;
            *reinterpret_cast<float4*>(smem + NI_ZJ_HI + 8192 + o) = lo;
        }
    };
    const uint32_t zt_base = (sj >> 5) * 16384 + ((sj & 3) << 2);     // transposed tile: k = j -> atom sj / 32, word sj % 32 of row d
    const uint32_t zt_chunk = (sj & 31) >> 2;
    auto stage_t = [&](const float4 (&z)[2], int buf) {              // transposed [d][j]: operand B of the gradient contraction
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            float4 hi, lo;
            split4(z[e], hi, lo);
            const float hv[4] = {hi.x, hi.y, hi.z, hi.w}, lv[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int d = 4 * (sc0 + e) + u;                      // d & 7 == 4 e + u (sc0 is even)
                const uint32_t ot = zt_base + d * 128 + ((zt_chunk ^ (uint32_t)(4 * e + u)) << 4);
                *reinterpret_cast<float*>(smem + NI_ZT + buf * NI_ZT_BUF + ot) = hv[u];
                *reinterpret_cast<float*>(smem + NI_ZT + buf * NI_ZT_BUF + 8192 + ot) = lv[u];
            }
        }
    };

    float4 zc[2], zn[2];
    if (worker) {
        // ---- Z_I once, scaled by -log2(e): thread = (row r = 32 (warp % 4) + lane, 64 bytes = chunks 4 (warp / 4) .. + 3 of the row)
        const int r = (warp & 3) * 32 + lane, c0 = (warp >> 2) * 4;
        const int64_t gr = (int64_t)ib * NI_BI + r;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = c0 + e;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gr < a.n_s) v = __ldg(reinterpret_cast<const float4*>(a.zs + gr * a.ldz) + c);
            v.x *= NI_NEG_LOG2E; v.y *= NI_NEG_LOG2E; v.z *= NI_NEG_LOG2E; v.w *= NI_NEG_LOG2E;
            float4 hi, lo;
            split4(v, hi, lo);
            const uint32_t o = (c >> 3) * 16384 + swz(r, c & 7);
            *reinterpret_cast<float4*>(smem + NI_ZI_HI + o) = hi;
            *reinterpret_cast<float4*>(smem + NI_ZI_LO + o) = lo;
        }
        load_zj(jb0, zc);
        if (T > 0) stage_k(zc);
    }
    tc_fence_before();
    __syncthreads();                                                  // sync 0: Z_I, Z_J(0), barriers, TMEM address
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (!worker) {
        // ================================ MMA issuer (warp 16) ================================
        // The whole warp takes part in the named barriers (bar.sync is warp-aligned); elect.sync picks the same lane
        // every time, so all MMAs and commits come from one thread.
        // hi and lo of a B operand are adjacent along N ([64 rows hi | 64 rows lo] per k-atom) and so are the main and the
        // correction accumulator in TMEM: a_hi x [b_hi | b_lo] is ONE N = 128 instruction, a k-step costs two MMAs, not three
        const uint32_t idesc64 = make_idesc(64), idesc128 = make_idesc(128);
        const uint64_t zi_hi = make_desc(smem_u32(smem + NI_ZI_HI)), zi_lo = make_desc(smem_u32(smem + NI_ZI_LO));
        const uint64_t zj = make_desc(smem_u32(smem + NI_ZJ_HI)), zt = make_desc(smem_u32(smem + NI_ZT));
        auto logits = [&](int buf) {                                  // P[buf] = (-log2 e Z_I) . Z_J^T
            const uint32_t d = tmem_base + NI_COL_P + 128 * buf;
            fence_proxy_async();                                      // the workers' generic stores -> async proxy
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < 8; ++s) {                             // k-step: atom s / 4 (16 KB), 32 bytes per step inside it
                const uint32_t o = (s >> 2) * (16384 >> 4) + (s & 3) * 2;
                umma_tf32(d, zi_hi + o, zj + o, idesc128, s != 0);    // main | corr (+)= a_hi x [b_hi | b_lo]
                umma_tf32(d + 64, zi_lo + o, zj + o, idesc64, 1);     // corr += a_lo x b_hi
            }
            umma_commit(&bar_p);
        };
        auto gradient = [&](int t) {                                  // dZ += C(t) . Z_J(t); A = C in tensor memory over P[t & 1]
            const uint32_t d = tmem_base + NI_COL_DZ, c_hi = tmem_base + NI_COL_P + 128 * (t & 1), c_lo = c_hi + 64;
            const uint64_t b = zt + (uint64_t)((t & 1) * (NI_ZT_BUF >> 4));
            fence_proxy_async();
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < 8; ++s) {                             // k-step: 8 columns of C, atom s / 4 (+ 32 bytes per step) of Z_J^T
                const uint32_t o = (s >> 2) * (16384 >> 4) + (s & 3) * 2;
                umma_tf32_tmem_a(d, c_hi + 8 * s, b + o, idesc128, !(t == 0 && s == 0));
                umma_tf32_tmem_a(d + 64, c_lo + 8 * s, b + o, idesc64, 1);
            }
        };
        auto l2_prefetch_tile = [&](int t) {                          // whole 32 KB target tile of block t -> L2
            if (t < T) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(tile0 + (int64_t)t * NI_TILE), "r"(NI_TILE * 4) : "memory");
        };
        if (elect_one()) {
            if (T > 0) logits(0);
            l2_prefetch_tile(1);
            l2_prefetch_tile(2);
        }
        __syncwarp();
        for (int t = 0; t < T; ++t) {
            asm volatile("bar.sync 1, %0;" ::"r"(NI_THREADS) : "memory");              // sync A(t): Z_J(t+1) staged, P[(t+1)&1] drained
            if (elect_one()) {
                if (t + 1 < T) logits((t + 1) & 1);
                l2_prefetch_tile(t + 3);
            }
            __syncwarp();
            asm volatile("bar.sync 2, %0;" ::"r"(NI_THREADS) : "memory");              // sync B(t): C(t), Z_J^T(t) stored
            if (elect_one()) {
                gradient(t);
                if (t == T - 1) umma_commit(&bar_end);
            }
            __syncwarp();
        }
    } else {
        // ================================ workers (warps 0-15) ================================
        load_zj(jb0 + 1, zn);
        // epilogue role: thread = (row i = 32 (warp % 4) + lane, logit columns 16 (warp / 4) .. + 15)
        const int q = warp & 3, cg = warp >> 2;
        const int ei = q * 32 + lane;
        const int64_t gi = (int64_t)ib * NI_BI + ei;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        float loss = 0.f;

        for (int t = 0; t < T; ++t) {
            const int jb = jb0 + t;
            // this block's target chunks (L2 hits: the tile was prefetched three blocks ago)
            const float4* tg_tile = reinterpret_cast<const float4*>(tile0 + (int64_t)t * NI_TILE);
            float4 tg[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) tg[c] = __ldg(tg_tile + (4 * cg + c) * NI_BI + ei);
            // ---- logit tile of block t complete (the tensor pipe is in order: so is everything issued before it)
            mbar_wait(&bar_p, (uint32_t)(t & 1));
            tc_fence_after();
            // ---- next block's Z_J; its logit tile runs under this block's epilogue
            if (t + 1 < T) stage_k(zn);
            asm volatile("bar.sync 1, %0;" ::"r"(NI_THREADS) : "memory");              // sync A(t)
            // ---- epilogue: logits -> coefficients (registers)
            uint32_t pm[16], pc[16];
            const uint32_t tP = t_lane + NI_COL_P + 128 * (t & 1) + 16 * cg;
            tmem_ld16_nowait(tP, pm);
            tmem_ld16_nowait(tP + 64, pc);
            tmem_wait_ld();
            // pairs of this tile that count towards the loss (i > j): all of them, none, or (2 of the sweep's blocks) a band
            const int64_t j_lo = (int64_t)jb * NI_BJ, i_lo = (int64_t)ib * NI_BI;
            const int tri = (j_lo + NI_BJ - 1 < i_lo) ? 1 : ((j_lo > i_lo + NI_BI - 1) ? 0 : 2);
            const int lim = (int)max((int64_t)-1, min((int64_t)NI_BJ, gi - (j_lo + 16 * cg)));   // element k of my 16 counts iff lim > k
            float cv[16];
            float sq = 0.f;
            auto coefficients = [&](auto band) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float tv[4] = {tg[c].x, tg[c].y, tg[c].z, tg[c].w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float s = sigmoid_from_scaled(__uint_as_float(pm[4 * c + u]) + __uint_as_float(pc[4 * c + u]));
                        const float r = tv[u] >= 0.f ? s - tv[u] : 0.f;              // sentinel -1: pair not in M
                        cv[4 * c + u] = r * fmaf(-s, s, s);                          // (s - t) s (1 - s); 2 w / |M| is applied to dz at the end
                        if (!decltype(band)::value || lim > 4 * c + u) sq = fmaf(r, r, sq);
                    }
                }
            };
            if (tri == 2) coefficients(std::true_type{}); else coefficients(std::false_type{});
            if (tri != 0) loss += sq;
            // ---- C in place over the logit tile (tensor memory); Z_J^T into the buffer MMA2(t-2) has released
            //      (it was issued before MMA1(t), whose completion this iteration has already awaited)
            uint32_t ch[16], cl[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                float hi, lo;
                split_coef(cv[k], hi, lo);
                ch[k] = __float_as_uint(hi); cl[k] = __float_as_uint(lo);
            }
            tmem_st16(tP, ch);
            tmem_st16(tP + 64, cl);
            stage_t(zc, t & 1);
            tmem_wait_st();
            tc_fence_before();
            asm volatile("bar.sync 2, %0;" ::"r"(NI_THREADS) : "memory");              // sync B(t)
            zc[0] = zn[0]; zc[1] = zn[1];
            load_zj(jb + 2, zn);
        }
        // ---- rows of dz owned by this CTA: thread = (row i, 16 embedding columns), scaled by 2 w / |M|
        float* orow = a.jsplit == 1 ? a.dz_out + gi * a.lddz + 16 * cg
                                    : a.dz_out + (((int64_t)split * a.n_ib + ib) * NI_BI + ei) * NI_D + 16 * cg;
        const bool row_ok = a.jsplit == 1 ? gi < a.n_s : true;
        uint32_t dm[16], dc[16];
        if (T > 0) {
            mbar_wait(&bar_end, 0);
            tc_fence_after();
            tmem_ld16_nowait(t_lane + NI_COL_DZ + 16 * cg, dm);
            tmem_ld16_nowait(t_lane + NI_COL_DZ + 64 + 16 * cg, dc);
            tmem_wait_ld();
        } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) { dm[e] = 0; dc[e] = 0; }
        }
        if (row_ok) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float4 o;
                o.x = a.scale2 * (__uint_as_float(dm[4 * c]) + __uint_as_float(dc[4 * c]));
                o.y = a.scale2 * (__uint_as_float(dm[4 * c + 1]) + __uint_as_float(dc[4 * c + 1]));
                o.z = a.scale2 * (__uint_as_float(dm[4 * c + 2]) + __uint_as_float(dc[4 * c + 2]));
                o.w = a.scale2 * (__uint_as_float(dm[4 * c + 3]) + __uint_as_float(dc[4 * c + 3]));
                reinterpret_cast<float4*>(orow)[c] = o;
            }
        }
        // ---- deterministic block reduction of the loss
        loss = warp_sum(loss);
        if (lane == 0) red[warp] = loss;
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int w = 0; w < NI_WORKER_WARPS; ++w) s += red[w];
        a.partial[blockIdx.x] = s;
    }
    if (!worker) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(NI_TMEM_COLS));
    }
}

// loss_sum = sum of the CTA partials in CTA order; with a split J sweep also dzs = sum of the parts in split order
__global__ void __launch_bounds__(256) dense_ni_tc_finish_kernel(const float* __restrict__ partial, int nparts, float* __restrict__ loss_sum,
                                                                 const float* __restrict__ parts, int jsplit, int64_t rows_pad, int64_t n_s,
                                                                 float* __restrict__ dzs, int64_t lddz) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < nparts; ++i) s += partial[i];
        *loss_sum = s;
    }
    if (jsplit > 1) {
        const int64_t total = n_s * (NI_D / 4);
        for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < total; u += (int64_t)gridDim.x * blockDim.x) {
            const int64_t i = u / (NI_D / 4);
            const int c = (int)(u - i * (NI_D / 4));
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int s = 0; s < jsplit; ++s) add4(acc, __ldg(reinterpret_cast<const float4*>(parts + ((int64_t)s * rows_pad + i) * NI_D) + c));
            reinterpret_cast<float4*>(dzs + i * lddz)[c] = acc;
        }
    }
}

// one float4 of the packed target per thread (setup path, once per plan)
__global__ void __launch_bounds__(256) dense_ni_tc_pack_kernel(const float* __restrict__ tgt, int64_t ldt, const uint32_t* __restrict__ excl,
                                                               int64_t n_s, int n_jb, int64_t total_f4, float4* __restrict__ packed) {
    for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < total_f4; u += (int64_t)gridDim.x * blockDim.x) {
        const int64_t tile = u / (NI_TILE / 4);
        const int w = (int)(u - tile * (NI_TILE / 4));
        const int c = w / NI_BI, i = w - c * NI_BI;
        const int64_t ib = tile / n_jb, jb = tile - ib * n_jb;
        const int64_t gi = ib * NI_BI + i;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int64_t gj = jb * NI_BJ + 4 * c + e;
            v[e] = -1.0f;
            if (gi < n_s && gj < n_s && gi != gj) {
                const int64_t bit = gi * n_s + gj;
                if (!((__ldg(excl + (bit >> 5)) >> (bit & 31)) & 1u)) v[e] = __ldg(tgt + gi * ldt + gj);
            }
        }
        packed[u] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

static int ni_jsplit(int64_t n_ib, int64_t n_jb) {
    return (int)std::max<int64_t>(1, std::min<int64_t>(n_jb, kNumSMs / std::max<int64_t>(1, n_ib)));
}

}  // namespace tc
}  // namespace gd

using namespace gd;

extern "C" int gd_dense_ni_tc_supported(int32_t dim) { return dim == tc::NI_D ? 1 : 0; }

extern "C" size_t gd_dense_ni_tc_target_bytes(int64_t n_s) {
    const int64_t n_ib = ceil_div<int64_t>(n_s, tc::NI_BI), n_jb = ceil_div<int64_t>(n_s, tc::NI_BJ);
    return (size_t)(n_ib * n_jb) * tc::NI_TILE * sizeof(float);
}

extern "C" int gd_dense_ni_tc_pack_target(const float* tgt_sig, int64_t ldt, const uint32_t* excl_bits, int64_t n_s, float* packed,
                                          gd_stream_t stream_) {
    if (n_s == 0) return GD_OK;
    GD_CHECK_ARG(tgt_sig && excl_bits && packed, "null pointer");
    GD_CHECK_ARG(ldt >= n_s && ((uintptr_t)packed % 16) == 0, "bad target layout");
    const int64_t n_ib = ceil_div<int64_t>(n_s, tc::NI_BI), n_jb = ceil_div<int64_t>(n_s, tc::NI_BJ);
    const int64_t total_f4 = n_ib * n_jb * (tc::NI_TILE / 4);
    const int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(total_f4, 256), kNumSMs * 32);
    tc::dense_ni_tc_pack_kernel<<<blocks, 256, 0, as_stream(stream_)>>>(tgt_sig, ldt, excl_bits, n_s, (int)n_jb, total_f4,
                                                                       reinterpret_cast<float4*>(packed));
    GD_LAUNCH_CHECK();
    return GD_OK;
}

extern "C" size_t gd_dense_ni_tc_workspace_bytes(int64_t n_s) {
    const int64_t n_ib = ceil_div<int64_t>(n_s, tc::NI_BI), n_jb = ceil_div<int64_t>(n_s, tc::NI_BJ);
    const int js = tc::ni_jsplit(n_ib, n_jb);
    size_t bytes = align_up((size_t)(n_ib * js + 1) * sizeof(float));
    if (js > 1) bytes += (size_t)js * n_ib * tc::NI_BI * tc::NI_D * sizeof(float);
    return bytes;
}

extern "C" int gd_dense_ni_tc_fwd_bwd(const float* zs, int64_t ldz, int64_t n_s, const float* packed_target, float coef_scale,
                                      float* dzs, int64_t lddz, float* loss_sum, void* workspace, size_t workspace_bytes,
                                      gd_stream_t stream_) {
    cudaStream_t stream = as_stream(stream_);
    GD_CHECK_ARG(loss_sum != nullptr, "null loss_sum");
    if (n_s == 0) { GD_CUDA(cudaMemsetAsync(loss_sum, 0, sizeof(float), stream)); return GD_OK; }
    GD_CHECK_ARG(zs && packed_target && dzs, "null pointer");
    GD_CHECK_ARG(ldz % 4 == 0 && lddz % 4 == 0 && (((uintptr_t)zs | (uintptr_t)dzs | (uintptr_t)packed_target) % 16) == 0,
                 "operands must be 16-byte aligned with leading dimensions that are multiples of 4");
    if (!workspace || workspace_bytes < gd_dense_ni_tc_workspace_bytes(n_s))
        return fail(GD_ERR_WORKSPACE, "gd_dense_ni_tc_fwd_bwd: workspace too small");
    const int64_t n_ib = ceil_div<int64_t>(n_s, tc::NI_BI), n_jb = ceil_div<int64_t>(n_s, tc::NI_BJ);
    const int js = tc::ni_jsplit(n_ib, n_jb);
    float* partial = static_cast<float*>(workspace);
    float* parts = js > 1 ? reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + align_up((size_t)(n_ib * js + 1) * sizeof(float))) : nullptr;
    tc::NiArgs a{zs, ldz, n_s, packed_target, (int)n_ib, (int)n_jb, 2.0f * coef_scale, js > 1 ? parts : dzs, lddz, partial, js};
    const int grid = (int)(n_ib * js);
    const size_t smem = tc::NI_SMEM + 1024;
    GD_CUDA(cudaFuncSetAttribute(tc::dense_ni_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc::dense_ni_tc_kernel<<<grid, tc::NI_THREADS, smem, stream>>>(a);
    GD_LAUNCH_CHECK();
    const int fblocks = js > 1 ? (int)std::min<int64_t>(ceil_div<int64_t>(n_s * (tc::NI_D / 4), 256), kNumSMs * 8) : 1;
    tc::dense_ni_tc_finish_kernel<<<fblocks, 256, 0, stream>>>(partial, grid, loss_sum, parts, js, n_ib * tc::NI_BI, n_s, dzs, lddz);
    GD_LAUNCH_CHECK();
    return GD_OK;
}
