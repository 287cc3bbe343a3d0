// GATConv(heads=1) edge-softmax aggregation, forward and backward (gat.py:11-12; PyG
// semantics in SURVEY.md §9.2):  e_ik = LeakyReLU(a_src[k] + a_dst[i]),
// alpha_ik = exp(e_ik - max_i) / (sum_i exp(e - max_i) + 1e-16),  out_i = sum_k alpha_ik h_k + b.
// One sub-warp (C/4 lanes) per destination row: a scalar pass for the row max, then one
// gather pass that accumulates exp-weights and the weighted feature sum together — the
// [nnz, C] message tensor and the three scatter passes of PyG's softmax are never
// materialised.  Backward = one destination-major SDDMM pass (d e per edge, written in
// transposed-CSR order) + one source-major gather pass; deterministic, no atomics.
#include "common.cuh"

namespace gd {

__device__ __forceinline__ float lrelu(float x, float slope) { return x > 0.f ? x : slope * x; }

template <int LANES>
__device__ __forceinline__ unsigned sub_mask(int sub) {
    return (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << (sub * LANES));
}
template <int LANES>
__device__ __forceinline__ float sub_sum(float v, unsigned mask) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, LANES);
    return v;
}
template <int LANES>
__device__ __forceinline__ float sub_max(float v, unsigned mask) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(mask, v, o, LANES));
    return v;
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}

template <int LANES>
__global__ void __launch_bounds__(256) gat_scores_kernel(const float* __restrict__ h, int64_t ldh, int64_t n,
                                                         const float* __restrict__ att_src,
                                                         const float* __restrict__ att_dst,
                                                         float* __restrict__ a_src, float* __restrict__ a_dst) {
    constexpr int PER_WARP = 32 / LANES;
    const int lane = threadIdx.x & 31, sub = lane / LANES, sl = lane % LANES;
    const unsigned mask = sub_mask<LANES>(sub);
    const int64_t row = ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5) * PER_WARP + sub;
    if (row >= n) return;
    const float4 v = ldg4(h + row * ldh + sl * 4);
    const float s = sub_sum<LANES>(dot4(v, ldg4(att_src + sl * 4)), mask);
    const float d = sub_sum<LANES>(dot4(v, ldg4(att_dst + sl * 4)), mask);
    if (sl == 0) { a_src[row] = s; a_dst[row] = d; }
}

struct GatFwdArgs {
    const int32_t* rowptr; const int32_t* col;
    const float* h; int64_t ldh;
    const float* a_src; const float* a_dst; const float* bias;
    float slope;
    float* out; int64_t ldo;
    float* rowmax; float* rowden;
    int64_t n;
};

template <int LANES>
__global__ void __launch_bounds__(256) gat_fwd_kernel(const GatFwdArgs a) {
    constexpr int PER_WARP = 32 / LANES;
    const int lane = threadIdx.x & 31, sub = lane / LANES, sl = lane % LANES;
    const unsigned mask = sub_mask<LANES>(sub);
    const int64_t row = ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5) * PER_WARP + sub;
    if (row >= a.n) return;
    const int beg = __ldg(a.rowptr + row), end = __ldg(a.rowptr + row + 1);
    const float ad = __ldg(a.a_dst + row);
    // pass 1: row max of the leaky-relu scores (scalar gathers only)
    float m = -INFINITY;
    for (int k = beg + sl; k < end; k += LANES) m = fmaxf(m, lrelu(__ldg(a.a_src + __ldg(a.col + k)) + ad, a.slope));
    m = sub_max<LANES>(m, mask);
    // pass 2: exp weights + weighted feature sum
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float den = 0.f;
    for (int base = beg; base < end; base += LANES) {
        const int k = base + sl;
        int c = 0; float w = 0.f;
        if (k < end) { c = __ldg(a.col + k); w = __expf(lrelu(__ldg(a.a_src + c) + ad, a.slope) - m); }
        den += w;
        const int cnt = min(LANES, end - base);
        for (int j = 0; j < cnt; j += 4) {
            float4 v[4]; float wj[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = j + u;
                const int cj = __shfl_sync(mask, c, idx & (LANES - 1), LANES);
                wj[u] = __shfl_sync(mask, w, idx & (LANES - 1), LANES);
                v[u] = idx < cnt ? ldg4(a.h + (int64_t)cj * a.ldh + sl * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx >= cnt) wj[u] = 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) fma4(acc, wj[u], v[u]);
        }
    }
    den = sub_sum<LANES>(den, mask) + 1e-16f;
    const float inv = 1.0f / den;
    acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
    if (a.bias) add4(acc, ldg4(a.bias + sl * 4));
    stg4(a.out + row * a.ldo + sl * 4, acc);
    if (sl == 0) { a.rowmax[row] = m; a.rowden[row] = den; }
}

struct GatBwdDstArgs {
    const int32_t* rowptr; const int32_t* col; const int32_t* tinv;
    const float* h; int64_t ldh;
    const float* a_src; const float* a_dst; const float* rowmax; const float* rowden;
    const float* gout; int64_t ldg;
    const float* out; int64_t ldo; const float* bias;
    float slope;
    float* alpha_t; float* dpre_t; float* da_dst;
    int64_t n;
};

// destination-major: per edge alpha and d(pre-activation score), written at the entry's
// position in the TRANSPOSED CSR (tinv) so the source-major pass streams them
template <int LANES>
__global__ void __launch_bounds__(256) gat_bwd_dst_kernel(const GatBwdDstArgs a) {
    constexpr int PER_WARP = 32 / LANES;
    const int lane = threadIdx.x & 31, sub = lane / LANES, sl = lane % LANES;
    const unsigned mask = sub_mask<LANES>(sub);
    const int64_t row = ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5) * PER_WARP + sub;
    if (row >= a.n) return;
    const int beg = __ldg(a.rowptr + row), end = __ldg(a.rowptr + row + 1);
    const float ad = __ldg(a.a_dst + row), m = __ldg(a.rowmax + row), inv = 1.0f / __ldg(a.rowden + row);
    const float4 g = ldg4(a.gout + row * a.ldg + sl * 4);
    float4 o = ldg4(a.out + row * a.ldo + sl * 4);
    if (a.bias) { const float4 b = ldg4(a.bias + sl * 4); o.x -= b.x; o.y -= b.y; o.z -= b.z; o.w -= b.w; }
    const float t = sub_sum<LANES>(dot4(g, o), mask);           // g_i . (sum_k alpha_ik h_k)
    float dad = 0.f;
    for (int base = beg; base < end; base += LANES) {
        const int k = base + sl;
        const int c = k < end ? __ldg(a.col + k) : 0;
        const int cnt = min(LANES, end - base);
        float mydot = 0.f;
        for (int j = 0; j < cnt; ++j) {
            const int cj = __shfl_sync(mask, c, j, LANES);
            const float d = sub_sum<LANES>(dot4(g, ldg4(a.h + (int64_t)cj * a.ldh + sl * 4)), mask);
            if (j == sl) mydot = d;
        }
        if (k < end) {
            const float pre = __ldg(a.a_src + c) + ad;
            const float alpha = __expf(lrelu(pre, a.slope) - m) * inv;
            const float dpre = alpha * (mydot - t) * (pre > 0.f ? 1.0f : a.slope);
            const int p = __ldg(a.tinv + k);
            a.alpha_t[p] = alpha;
            a.dpre_t[p] = dpre;
            dad += dpre;
        }
    }
    dad = sub_sum<LANES>(dad, mask);
    if (sl == 0) a.da_dst[row] = dad;
}

struct GatBwdSrcArgs {
    const int32_t* rowptr; const int32_t* col;      // transposed CSR: row = source k, col = destination i
    const float* alpha_t; const float* dpre_t;
    const float* gout; int64_t ldg;
    const float* att_src; const float* att_dst; const float* da_dst;
    float* dh; int64_t lddh; float* da_src;
    int64_t n;
};

template <int LANES>
__global__ void __launch_bounds__(256) gat_bwd_src_kernel(const GatBwdSrcArgs a) {
    constexpr int PER_WARP = 32 / LANES;
    const int lane = threadIdx.x & 31, sub = lane / LANES, sl = lane % LANES;
    const unsigned mask = sub_mask<LANES>(sub);
    const int64_t row = ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5) * PER_WARP + sub;
    if (row >= a.n) return;
    const int beg = __ldg(a.rowptr + row), end = __ldg(a.rowptr + row + 1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float das = 0.f;
    for (int base = beg; base < end; base += LANES) {
        const int k = base + sl;
        int c = 0; float w = 0.f;
        if (k < end) { c = __ldg(a.col + k); w = __ldg(a.alpha_t + k); das += __ldg(a.dpre_t + k); }
        const int cnt = min(LANES, end - base);
        for (int j = 0; j < cnt; j += 4) {
            float4 v[4]; float wj[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = j + u;
                const int cj = __shfl_sync(mask, c, idx & (LANES - 1), LANES);
                wj[u] = __shfl_sync(mask, w, idx & (LANES - 1), LANES);
                v[u] = idx < cnt ? ldg4(a.gout + (int64_t)cj * a.ldg + sl * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx >= cnt) wj[u] = 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) fma4(acc, wj[u], v[u]);
        }
    }
    das = sub_sum<LANES>(das, mask);
    // d h_k = sum_i alpha_ik g_i + d a_src[k] * att_src + d a_dst[k] * att_dst
    fma4(acc, das, ldg4(a.att_src + sl * 4));
    fma4(acc, __ldg(a.da_dst + row), ldg4(a.att_dst + sl * 4));
    stg4(a.dh + row * a.lddh + sl * 4, acc);
    if (sl == 0) a.da_src[row] = das;
}

static unsigned gat_grid(int64_t n, int lanes) {
    return (unsigned)ceil_div<int64_t>(ceil_div<int64_t>(n, 32 / lanes), 8);
}

}  // namespace gd

using namespace gd;

#define GAT_DISPATCH(C, KERNEL, N, ...)                                                          \
    do {                                                                                         \
        if ((C) == 128) KERNEL<32><<<gat_grid(N, 32), 256, 0, stream>>>(__VA_ARGS__);            \
        else if ((C) == 64) KERNEL<16><<<gat_grid(N, 16), 256, 0, stream>>>(__VA_ARGS__);        \
        else if ((C) == 32) KERNEL<8><<<gat_grid(N, 8), 256, 0, stream>>>(__VA_ARGS__);          \
        else return fail(GD_ERR_INVALID, std::string(__func__) + ": out_channels must be 32, 64 or 128"); \
        GD_LAUNCH_CHECK();                                                                       \
    } while (0)

extern "C" int gd_gat_scores(const float* h, int64_t ldh, int64_t n, int32_t c, const float* att_src,
                             const float* att_dst, float* a_src, float* a_dst, gd_stream_t stream_) {
    cudaStream_t stream = as_stream(stream_);
    if (n == 0) return GD_OK;
    GD_CHECK_ARG(h && att_src && att_dst && a_src && a_dst, "null pointer");
    GD_CHECK_ARG(ldh % 4 == 0 && ldh >= c, "ldh must be a multiple of 4");
    GAT_DISPATCH(c, gat_scores_kernel, n, h, ldh, n, att_src, att_dst, a_src, a_dst);
    return GD_OK;
}

extern "C" int gd_gat_fwd(const gd_csr_t* csr, const float* h, int64_t ldh, int32_t c, const float* a_src,
                          const float* a_dst, const float* bias, float negative_slope, float* out, int64_t ldo,
                          float* rowmax, float* rowden, gd_stream_t stream_) {
    cudaStream_t stream = as_stream(stream_);
    GD_CHECK_ARG(csr != nullptr, "null csr");
    if (csr->num_rows == 0) return GD_OK;
    GD_CHECK_ARG(csr->rowptr && csr->col && h && a_src && a_dst && out && rowmax && rowden, "null pointer");
    GD_CHECK_ARG(ldh % 4 == 0 && ldo % 4 == 0, "leading dimensions must be multiples of 4");
    GatFwdArgs a{csr->rowptr, csr->col, h, ldh, a_src, a_dst, bias, negative_slope, out, ldo, rowmax, rowden, csr->num_rows};
    GAT_DISPATCH(c, gat_fwd_kernel, a.n, a);
    return GD_OK;
}

extern "C" int gd_gat_bwd_dst(const gd_csr_t* csr, const int32_t* tinv, const float* h, int64_t ldh, int32_t c,
                              const float* a_src, const float* a_dst, const float* rowmax, const float* rowden,
                              const float* gout, int64_t ldg, const float* out, int64_t ldo, const float* bias,
                              float negative_slope, float* alpha_t, float* dpre_t, float* da_dst,
                              gd_stream_t stream_) {
    cudaStream_t stream = as_stream(stream_);
    GD_CHECK_ARG(csr != nullptr, "null csr");
    if (csr->num_rows == 0) return GD_OK;
    GD_CHECK_ARG(csr->rowptr && csr->col && tinv && h && a_src && a_dst && rowmax && rowden && gout && out &&
                 alpha_t && dpre_t && da_dst, "null pointer");
    GD_CHECK_ARG(ldh % 4 == 0 && ldg % 4 == 0 && ldo % 4 == 0, "leading dimensions must be multiples of 4");
    GatBwdDstArgs a{csr->rowptr, csr->col, tinv, h, ldh, a_src, a_dst, rowmax, rowden, gout, ldg, out, ldo, bias,
                    negative_slope, alpha_t, dpre_t, da_dst, csr->num_rows};
    GAT_DISPATCH(c, gat_bwd_dst_kernel, a.n, a);
    return GD_OK;
}

extern "C" int gd_gat_bwd_src(const gd_csr_t* csr_t, const float* alpha_t, const float* dpre_t, const float* gout,
                              int64_t ldg, int32_t c, const float* att_src, const float* att_dst,
                              const float* da_dst, float* dh, int64_t lddh, float* da_src, gd_stream_t stream_) {
    cudaStream_t stream = as_stream(stream_);
    GD_CHECK_ARG(csr_t != nullptr, "null csr");
    if (csr_t->num_rows == 0) return GD_OK;
    GD_CHECK_ARG(csr_t->rowptr && csr_t->col && alpha_t && dpre_t && gout && att_src && att_dst && da_dst && dh && da_src,
                 "null pointer");
    GD_CHECK_ARG(ldg % 4 == 0 && lddh % 4 == 0, "leading dimensions must be multiples of 4");
    GatBwdSrcArgs a{csr_t->rowptr, csr_t->col, alpha_t, dpre_t, gout, ldg, att_src, att_dst, da_dst, dh, lddh, da_src,
                    csr_t->num_rows};
    GAT_DISPATCH(c, gat_bwd_src_kernel, a.n, a);
    return GD_OK;
}
