// Dense-block Neighbourhood-Influence loss of train_fullbatch (gnndelete.py:163-193, 239-241) on the 5th-generation
// tensor cores: the one genuinely compute-bound contraction of the path (S2 x S2 x 64, 158 M node pairs at the Cora
// shape; the fp32 CUDA-core kernel in dense_ni.cu needs 4.7 ms per epoch for it, 30x the rest of that epoch).
//
//   loss_l = mean_{(i,j) in M} ( sigmoid(<z_i, z_j>) - sigmoid(logits_ori[i, j]) )^2,   dz_i = sum_j c_ij z_j
//
// One CTA owns 128 rows I of z_S and sweeps the 64-row column blocks J (flash-attention shaped, nothing N x N is formed):
//   MMA1   P[128 x 64]   = Z_I . Z_J^T            tcgen05.mma kind::tf32, M = 128, N = 64, K = 64, accumulators in TMEM
//   epilogue (thread = row i, 32 columns each): tcgen05.ld P, s = sigmoid(P), residual against the packed target tile,
//            coefficient c = 2 w/|M| (s - t) s (1 - s), loss partial; c is split hi / lo and stored as the K-major
//            SWIZZLE_128B A operand of the second contraction
//   MMA2   dZ_I[128 x 64] += C[128 x 64] . Z_J     same instruction shape, accumulated in TMEM over the whole sweep
// Both contractions keep fp32 accuracy with the 3xTF32 operand split of gemm_tc.cu (main + correction accumulators).
// Every operand is K-major (the layout the row GEMM has exercised since round 1): Z_J is staged twice, as [j][d] for
// MMA1 and transposed as [d][j] for MMA2 (4-byte conflict-free stores, lanes = j).
//
// The target block sigmoid(logits_ori[S][:, S]) is PACKED once per plan (gd_dense_ni_tc_pack_target): per (128 x 64)
// tile, 16-byte chunk c of row i at float4 index c * 128 + i, so a warp reads 512 contiguous bytes per load; excluded
// pairs (Df, both orders), the diagonal and the padding carry the sentinel -1 (a sigmoid is never negative), so the
// kernel reads no bitmap.  Every unordered pair is visited from both sides: a CTA writes only its own rows of dz -
// deterministic, no atomics.  Small S: the J sweep is split over `jsplit` CTAs per row block and a second kernel
// adds the parts in order.
#include "tc_common.cuh"

namespace gd {
namespace tc {

constexpr int NI_BI = 128;                 // rows of z_S per CTA (UMMA M)
constexpr int NI_BJ = 64;                  // column block: UMMA N of the logit tile, K of the gradient contraction
constexpr int NI_D = 64;                   // embedding width (out_dim of the reference's models)
constexpr int NI_THREADS = 256;
constexpr int NI_TILE = NI_BI * NI_BJ;     // floats of one packed target tile
// shared-memory map (bytes; every operand tile 1024-aligned; "atom" = 32 k values = one 128-byte swizzle row)
constexpr int NI_ZI_HI = 0;                        // Z_I   [128 rows i][k = d]   2 atoms x 16 KB
constexpr int NI_ZI_LO = NI_ZI_HI + 2 * 16384;
constexpr int NI_ZJ_HI = NI_ZI_LO + 2 * 16384;     // Z_J   [64 rows j][k = d]    2 atoms x 8 KB     B of MMA1
constexpr int NI_ZJ_LO = NI_ZJ_HI + 2 * 8192;
constexpr int NI_ZT_HI = NI_ZJ_LO + 2 * 8192;      // Z_J^T [64 rows d][k = j]    2 atoms x 8 KB     B of MMA2
constexpr int NI_ZT_LO = NI_ZT_HI + 2 * 8192;
constexpr int NI_C_HI = NI_ZT_LO + 2 * 8192;       // C     [128 rows i][k = j]   2 atoms x 16 KB    A of MMA2
constexpr int NI_C_LO = NI_C_HI + 2 * 16384;
constexpr int NI_SMEM = NI_C_LO + 2 * 16384;       // 196608
constexpr int NI_TMEM_COLS = 256;                  // P main | P corr | dZ main | dZ corr, 64 columns each

struct NiArgs {
    const float* zs; int64_t ldz; int64_t n_s;
    const float* packed; int n_ib, n_jb;
    float scale2;                                   // 2 * weight / |M|
    float* dz_out; int64_t lddz;                    // jsplit == 1: dzs [n_s][lddz];  else parts [jsplit][n_ib * 128][64]
    float* partial;                                 // [gridDim.x] sums of squared residuals over i > j
    int jsplit;
};

// sigmoid(p) with two MUFU ops.  ex2.approx is accurate to 2^-22 but the fp32 product p * log2(e) is not (its rounding
// error grows with |p|), so the product's exact residual and the low part of log2(e) are folded back in.
__device__ __forceinline__ float sigmoid_mufu(float p) {
    const float L2E = 1.4426950408889634f;
    float x = -p * L2E;
    float err = fmaf(-p, L2E, -x);
    err = fmaf(-p, 1.925963033500e-8f, err);
    x = fminf(x, 126.0f);                           // keeps e finite (sigmoid underflows to ~1e-38 there)
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x));
    e = fmaf(e, err * 0.6931471805599453f, e);
    float s;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(1.0f + e));
    return s;
}

__global__ void __launch_bounds__(NI_THREADS, 1) dense_ni_tc_kernel(const NiArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // by offset: keeps the shared address space
    __shared__ uint64_t bar_p, bar_d;
    __shared__ uint32_t tmem_base_smem;
    __shared__ float red[NI_THREADS / 32];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ib = blockIdx.x / a.jsplit, split = blockIdx.x - ib * a.jsplit;
    const int per = (a.n_jb + a.jsplit - 1) / a.jsplit;
    const int jb0 = split * per, jb1 = min(a.n_jb, jb0 + per);

    if (tid == 0) {
        mbar_init(&bar_p, 1);
        mbar_init(&bar_d, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"(NI_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // ---- Z_I once: thread = (row r = 32 (warp % 4) + lane, atom = warp / 4): 128 contiguous bytes of its row
    {
        const int r = (warp & 3) * 32 + lane, atom = warp >> 2;
        const int64_t gi = (int64_t)ib * NI_BI + r;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gi < a.n_s) v = __ldg(reinterpret_cast<const float4*>(a.zs + gi * a.ldz + atom * 32) + c);
            float4 hi, lo;
            split4(v, hi, lo);
            const uint32_t o = atom * 16384 + swz(r, c);
            *reinterpret_cast<float4*>(smem + NI_ZI_HI + o) = hi;
            *reinterpret_cast<float4*>(smem + NI_ZI_LO + o) = lo;
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    // Z_J staging role: thread = (row j = 32 (warp % 2) + lane, 64 bytes = chunks 4 (warp / 2) .. + 3 of the row)
    const int sj = (warp & 1) * 32 + lane, sc0 = (warp >> 1) * 4;
    float4 zj[4];
    auto load_zj = [&](int jb) {
        const int64_t gj = (int64_t)jb * NI_BJ + sj;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            zj[e] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (jb < jb1 && gj < a.n_s) zj[e] = __ldg(reinterpret_cast<const float4*>(a.zs + gj * a.ldz) + sc0 + e);
        }
    };
    // epilogue role: thread = (row i = 32 (warp % 4) + lane, columns 32 (warp / 4) .. + 31 = atom warp / 4 of the C tile)
    const int q = warp & 3, h = warp >> 2;
    const int ei = q * 32 + lane;
    const int64_t gi = (int64_t)ib * NI_BI + ei;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t idesc = make_idesc(NI_BJ);
    float loss = 0.f;

    load_zj(jb0);
    int t = 0;
    for (int jb = jb0; jb < jb1; ++jb, ++t) {
        // ---- (1) the previous block's second contraction has read Z_J^T and C
        if (t > 0) mbar_wait(&bar_d, (uint32_t)((t - 1) & 1));
        // ---- (2) stage Z_J: K-major [j][d] (128-bit stores) and transposed [d][j] (32-bit stores, lanes = j)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float4 hi, lo;
            split4(zj[e], hi, lo);
            const int c = sc0 + e;                                    // 16-byte chunk of the row: atom c / 8, chunk c % 8
            const uint32_t o = (c >> 3) * 8192 + swz(sj, c & 7);
            *reinterpret_cast<float4*>(smem + NI_ZJ_HI + o) = hi;
            *reinterpret_cast<float4*>(smem + NI_ZJ_LO + o) = lo;
            const float hv[4] = {hi.x, hi.y, hi.z, hi.w}, lv[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int d = 4 * c + u;                              // row of the transposed tile; k = j: atom sj / 32, word sj % 32
                const uint32_t ot = (sj >> 5) * 8192 + d * 128 + ((((sj & 31) >> 2) ^ (d & 7)) << 4) + ((sj & 3) << 2);
                *reinterpret_cast<float*>(smem + NI_ZT_HI + ot) = hv[u];
                *reinterpret_cast<float*>(smem + NI_ZT_LO + ot) = lv[u];
            }
        }
        fence_proxy_async();
        __syncthreads();
        // ---- (3) logit tile
        if (warp == 0 && elect_one()) {
            tc_fence_after();
            const uint64_t a_hi = make_desc(smem_u32(smem + NI_ZI_HI)), a_lo = make_desc(smem_u32(smem + NI_ZI_LO));
            const uint64_t b_hi = make_desc(smem_u32(smem + NI_ZJ_HI)), b_lo = make_desc(smem_u32(smem + NI_ZJ_LO));
            const uint32_t dP = tmem_base, dPc = tmem_base + 64;
#pragma unroll
            for (int s = 0; s < 8; ++s) {                             // k-step: atom s / 4, 32 bytes per step inside it
                const uint32_t ao = (s >> 2) * (16384 >> 4) + (s & 3) * 2, bo = (s >> 2) * (8192 >> 4) + (s & 3) * 2;
                umma_tf32(dPc, a_lo + ao, b_hi + bo, idesc, s != 0);
                umma_tf32(dPc, a_hi + ao, b_lo + bo, idesc, 1);
                umma_tf32(dP, a_hi + ao, b_hi + bo, idesc, s != 0);
            }
            umma_commit(&bar_p);
        }
        __syncwarp();
        // ---- (4) while the tensor core works: this block's target chunks and the next block's rows
        const float4* tg_tile = reinterpret_cast<const float4*>(a.packed + ((int64_t)ib * a.n_jb + jb) * NI_TILE);
        float4 tg[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) tg[c] = __ldg(tg_tile + (8 * h + c) * NI_BI + ei);
        load_zj(jb + 1);
        // ---- (5) epilogue: logits -> coefficients
        mbar_wait(&bar_p, (uint32_t)(t & 1));
        tc_fence_after();
        uint32_t pm[32], pc[32];
        tmem_ld32_nowait(t_lane + 32 * h, pm);
        tmem_ld32_nowait(t_lane + 64 + 32 * h, pc);
        tmem_wait_ld();
        const int64_t gj0 = (int64_t)jb * NI_BJ + 32 * h;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float tv[4] = {tg[c].x, tg[c].y, tg[c].z, tg[c].w};
            float cv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float p = __uint_as_float(pm[4 * c + u]) + __uint_as_float(pc[4 * c + u]);
                const float s = sigmoid_mufu(p);
                const float r = s - tv[u];
                const bool on = tv[u] >= 0.f;
                cv[u] = on ? a.scale2 * r * s * (1.0f - s) : 0.f;
                if (on && gi > gj0 + 4 * c + u) loss = fmaf(r, r, loss);
            }
            float4 hi, lo;
            split4(make_float4(cv[0], cv[1], cv[2], cv[3]), hi, lo);
            const uint32_t o = h * 16384 + swz(ei, c);
            *reinterpret_cast<float4*>(smem + NI_C_HI + o) = hi;
            *reinterpret_cast<float4*>(smem + NI_C_LO + o) = lo;
        }
        tc_fence_before();
        fence_proxy_async();
        __syncthreads();
        // ---- (6) gradient contraction, accumulated over the sweep
        if (warp == 0 && elect_one()) {
            tc_fence_after();
            const uint64_t a_hi = make_desc(smem_u32(smem + NI_C_HI)), a_lo = make_desc(smem_u32(smem + NI_C_LO));
            const uint64_t b_hi = make_desc(smem_u32(smem + NI_ZT_HI)), b_lo = make_desc(smem_u32(smem + NI_ZT_LO));
            const uint32_t dD = tmem_base + 128, dDc = tmem_base + 192;
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const uint32_t ao = (s >> 2) * (16384 >> 4) + (s & 3) * 2, bo = (s >> 2) * (8192 >> 4) + (s & 3) * 2;
                const uint32_t accum = (t != 0) || (s != 0);
                umma_tf32(dDc, a_lo + ao, b_hi + bo, idesc, accum);
                umma_tf32(dDc, a_hi + ao, b_lo + bo, idesc, 1);
                umma_tf32(dD, a_hi + ao, b_hi + bo, idesc, accum);
            }
            umma_commit(&bar_d);
        }
        __syncwarp();
    }
    // ---- rows of dz owned by this CTA
    {
        float* orow = a.jsplit == 1 ? a.dz_out + gi * a.lddz + 32 * h
                                    : a.dz_out + (((int64_t)split * a.n_ib + ib) * NI_BI + ei) * NI_D + 32 * h;
        const bool row_ok = a.jsplit == 1 ? gi < a.n_s : true;
        uint32_t dm[32], dc[32];
        if (t > 0) {
            mbar_wait(&bar_d, (uint32_t)((t - 1) & 1));
            tc_fence_after();
            tmem_ld32_nowait(t_lane + 128 + 32 * h, dm);
            tmem_ld32_nowait(t_lane + 192 + 32 * h, dc);
            tmem_wait_ld();
        } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) { dm[e] = 0; dc[e] = 0; }
        }
        if (row_ok) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
                reinterpret_cast<float4*>(orow)[c] =
                    make_float4(__uint_as_float(dm[4 * c]) + __uint_as_float(dc[4 * c]), __uint_as_float(dm[4 * c + 1]) + __uint_as_float(dc[4 * c + 1]),
                                __uint_as_float(dm[4 * c + 2]) + __uint_as_float(dc[4 * c + 2]), __uint_as_float(dm[4 * c + 3]) + __uint_as_float(dc[4 * c + 3]));
        }
    }
    // ---- deterministic block reduction of the loss
    loss = warp_sum(loss);
    if (lane == 0) red[warp] = loss;
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int w = 0; w < NI_THREADS / 32; ++w) s += red[w];
        a.partial[blockIdx.x] = s;
    }
    if (warp == 0) {
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
    const size_t smem = tc::NI_SMEM + 1024;
    GD_CUDA(cudaFuncSetAttribute(tc::dense_ni_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)(n_ib * js);
    tc::dense_ni_tc_kernel<<<grid, tc::NI_THREADS, smem, stream>>>(a);
    GD_LAUNCH_CHECK();
    const int fblocks = js > 1 ? (int)std::min<int64_t>(ceil_div<int64_t>(n_s * (tc::NI_D / 4), 256), kNumSMs * 8) : 1;
    tc::dense_ni_tc_finish_kernel<<<fblocks, 256, 0, stream>>>(partial, grid, loss_sum, parts, js, n_ib * tc::NI_BI, n_s, dzs, lddz);
    GD_LAUNCH_CHECK();
    return GD_OK;
}
