// GATConv(heads=1) edge-softmax aggregation, forward and backward (gat.py:11-12; PyG
// semantics in SURVEY.md §9.2):  e_ik = LeakyReLU(a_src[k] + a_dst[i]),
// alpha_ik = exp(e_ik - max_i) / (sum_i exp(e - max_i) + 1e-16),  out_i = sum_k alpha_ik h_k + b.
// One sub-warp (C/4 lanes) per destination row: a scalar pass for the row max, then one
// gather pass that accumulates exp-weights and the weighted feature sum together — the
// [nnz, C] message tensor and the three scatter passes of PyG's softmax are never
// materialised.  Backward = one destination-major SDDMM pass (d e per edge, written in
// transposed-CSR order) + one source-major gather pass; deterministic, no float atomics.
// Rows longer than the CSR's split length (gd_spmm_plan_build, 128 entries) are processed as SEGMENTS by separate
// sub-warps: a segment leaves its partial result (for the softmax: running max, exp-sum and weighted sum relative to
// that max) in scratch, and the sub-warp that completes the row (ticket counter) merges the segments in order - a
// 10,000-neighbour hub is ~80 sub-warps of work instead of one.
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

// split plan of the CSR (all null / zero: no splitting)
struct GatSplit {
    int seg_len, num_seg;
    const int32_t* heavy_seg_beg; const int32_t* heavy_nseg; const int32_t* seg_row; const int32_t* seg_beg;
    const int32_t* seg_heavy; int32_t* heavy_ticket;
    float* scratch;
};
static GatSplit make_split(const gd_csr_t* csr, float* scratch) {
    GatSplit sp{0, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, scratch};
    if (csr->seg_len > 0 && csr->num_seg > 0 && scratch) {
        sp.seg_len = csr->seg_len; sp.num_seg = csr->num_seg;
        sp.heavy_seg_beg = csr->heavy_seg_beg; sp.heavy_nseg = csr->heavy_nseg; sp.seg_row = csr->seg_row;
        sp.seg_beg = csr->seg_beg; sp.seg_heavy = csr->seg_heavy; sp.heavy_ticket = csr->heavy_ticket;
    }
    return sp;
}
// Work item w of a launch over n rows + num_seg segments -> (row, [beg, end), seg).  seg = -1: a whole (light) row;
// returns false when the item has nothing to do (a heavy row's own slot, or past the end).
__device__ __forceinline__ bool gat_work(int64_t w, int64_t n, const int32_t* rowptr, const GatSplit& sp, int64_t& row,
                                         int& beg, int& end, int& seg) {
    if (w < n) {
        row = w; seg = -1;
        beg = __ldg(rowptr + row); end = __ldg(rowptr + row + 1);
        return !(sp.num_seg > 0 && end - beg > sp.seg_len);
    }
    seg = (int)(w - n);
    if (seg >= sp.num_seg) return false;
    row = __ldg(sp.seg_row + seg);
    beg = __ldg(sp.seg_beg + seg);
    end = min(beg + sp.seg_len, __ldg(rowptr + row + 1));
    return true;
}
// last-arriving segment of heavy row h (fixed merge order afterwards); re-arms the counter
template <int LANES>
__device__ __forceinline__ bool gat_last_segment(const GatSplit& sp, int h, int sl, unsigned mask) {
    __threadfence();
    int ticket = 0;
    if (sl == 0) ticket = atomicAdd(sp.heavy_ticket + h, 1);
    ticket = __shfl_sync(mask, ticket, 0, LANES);
    const bool last = ticket == __ldg(sp.heavy_nseg + h) - 1;
    if (last) { __threadfence(); if (sl == 0) sp.heavy_ticket[h] = 0; }
    return last;
}

struct GatFwdArgs {
    const int32_t* rowptr; const int32_t* col;
    const float* h; int64_t ldh;
    const float* a_src; const float* a_dst; const float* bias;
    float slope;
    float* out; int64_t ldo;
    float* rowmax; float* rowden;
    int64_t n;
    GatSplit sp;
};

template <int LANES>
__global__ void __launch_bounds__(256) gat_fwd_kernel(const GatFwdArgs a) {
    constexpr int PER_WARP = 32 / LANES, C = 4 * LANES;
    const int lane = threadIdx.x & 31, sub = lane / LANES, sl = lane % LANES;
    const unsigned mask = sub_mask<LANES>(sub);
    const int64_t w = ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5) * PER_WARP + sub;
    int64_t row; int beg, end, seg;
    if (!gat_work(w, a.n, a.rowptr, a.sp, row, beg, end, seg)) return;
    const float ad = __ldg(a.a_dst + row);
    // pass 1: max of the leaky-relu scores of this row / segment (scalar gathers only)
    float m = -INFINITY;
    for (int k = beg + sl; k < end; k += LANES) m = fmaxf(m, lrelu(__ldg(a.a_src + __ldg(a.col + k)) + ad, a.slope));
    m = sub_max<LANES>(m, mask);
    // pass 2: exp weights + weighted feature sum
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float den = 0.f;
    for (int base = beg; base < end; base += LANES) {
        const int k = base + sl;
        int c = 0; float wgt = 0.f;
        if (k < end) { c = __ldg(a.col + k); wgt = __expf(lrelu(__ldg(a.a_src + c) + ad, a.slope) - m); }
        den += wgt;
        const int cnt = min(LANES, end - base);
        for (int j = 0; j < cnt; j += 4) {
            float4 v[4]; float wj[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = j + u;
                const int cj = __shfl_sync(mask, c, idx & (LANES - 1), LANES);
                wj[u] = __shfl_sync(mask, wgt, idx & (LANES - 1), LANES);
                v[u] = idx < cnt ? ldg4(a.h + (int64_t)cj * a.ldh + sl * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx >= cnt) wj[u] = 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) fma4(acc, wj[u], v[u]);
        }
    }
    den = sub_sum<LANES>(den, mask);
    if (seg >= 0) {                          // a segment of a long row: partial (acc, m, den) -> scratch, the last one merges
        float* sp = a.sp.scratch + (int64_t)seg * (C + 4);
        stg4(sp + sl * 4, acc);
        if (sl == 0) { sp[C] = m; sp[C + 1] = den; }
        const int h = __ldg(a.sp.seg_heavy + seg);
        if (!gat_last_segment<LANES>(a.sp, h, sl, mask)) return;
        const int s0 = __ldg(a.sp.heavy_seg_beg + h), ns = __ldg(a.sp.heavy_nseg + h);
        m = -INFINITY;
        for (int q = 0; q < ns; ++q) m = fmaxf(m, __ldcg(a.sp.scratch + (int64_t)(s0 + q) * (C + 4) + C));
        acc = make_float4(0.f, 0.f, 0.f, 0.f); den = 0.f;
        for (int q = 0; q < ns; ++q) {
            const float* p = a.sp.scratch + (int64_t)(s0 + q) * (C + 4);
            const float f = __expf(__ldcg(p + C) - m);
            den = fmaf(f, __ldcg(p + C + 1), den);
            fma4(acc, f, __ldcg(reinterpret_cast<const float4*>(p) + sl));
        }
    }
    den += 1e-16f;
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
    GatSplit sp;
};

// destination-major: per edge alpha and d(pre-activation score), written at the entry's
// position in the TRANSPOSED CSR (tinv) so the source-major pass streams them
template <int LANES>
__global__ void __launch_bounds__(256) gat_bwd_dst_kernel(const GatBwdDstArgs a) {
    constexpr int PER_WARP = 32 / LANES;
    const int lane = threadIdx.x & 31, sub = lane / LANES, sl = lane % LANES;
    const unsigned mask = sub_mask<LANES>(sub);
    const int64_t w = ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5) * PER_WARP + sub;
    int64_t row; int beg, end, seg;
    if (!gat_work(w, a.n, a.rowptr, a.sp, row, beg, end, seg)) return;
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
    if (seg >= 0) {                          // partial d a_dst of a long row's segment; the last one adds them in order
        if (sl == 0) a.sp.scratch[seg] = dad;
        const int h = __ldg(a.sp.seg_heavy + seg);
        if (!gat_last_segment<LANES>(a.sp, h, sl, mask)) return;
        const int s0 = __ldg(a.sp.heavy_seg_beg + h), ns = __ldg(a.sp.heavy_nseg + h);
        dad = 0.f;
        for (int q = 0; q < ns; ++q) dad += __ldcg(a.sp.scratch + s0 + q);
    }
    if (sl == 0) a.da_dst[row] = dad;
}

struct GatBwdSrcArgs {
    const int32_t* rowptr; const int32_t* col;      // transposed CSR: row = source k, col = destination i
    const float* alpha_t; const float* dpre_t;
    const float* gout; int64_t ldg;
    const float* att_src; const float* att_dst; const float* da_dst;
    float* dh; int64_t lddh; float* da_src;
    int64_t n;
    GatSplit sp;
};

template <int LANES>
__global__ void __launch_bounds__(256) gat_bwd_src_kernel(const GatBwdSrcArgs a) {
    constexpr int PER_WARP = 32 / LANES, C = 4 * LANES;
    const int lane = threadIdx.x & 31, sub = lane / LANES, sl = lane % LANES;
    const unsigned mask = sub_mask<LANES>(sub);
    const int64_t w = ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5) * PER_WARP + sub;
    int64_t row; int beg, end, seg;
    if (!gat_work(w, a.n, a.rowptr, a.sp, row, beg, end, seg)) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float das = 0.f;
    for (int base = beg; base < end; base += LANES) {
        const int k = base + sl;
        int c = 0; float wgt = 0.f;
        if (k < end) { c = __ldg(a.col + k); wgt = __ldg(a.alpha_t + k); das += __ldg(a.dpre_t + k); }
        const int cnt = min(LANES, end - base);
        for (int j = 0; j < cnt; j += 4) {
            float4 v[4]; float wj[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = j + u;
                const int cj = __shfl_sync(mask, c, idx & (LANES - 1), LANES);
                wj[u] = __shfl_sync(mask, wgt, idx & (LANES - 1), LANES);
                v[u] = idx < cnt ? ldg4(a.gout + (int64_t)cj * a.ldg + sl * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx >= cnt) wj[u] = 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) fma4(acc, wj[u], v[u]);
        }
    }
    das = sub_sum<LANES>(das, mask);
    if (seg >= 0) {                          // partial (sum alpha g, sum dpre) of a long row's segment
        float* sp = a.sp.scratch + (int64_t)seg * (C + 4);
        stg4(sp + sl * 4, acc);
        if (sl == 0) sp[C] = das;
        const int h = __ldg(a.sp.seg_heavy + seg);
        if (!gat_last_segment<LANES>(a.sp, h, sl, mask)) return;
        const int s0 = __ldg(a.sp.heavy_seg_beg + h), ns = __ldg(a.sp.heavy_nseg + h);
        acc = make_float4(0.f, 0.f, 0.f, 0.f); das = 0.f;
        for (int q = 0; q < ns; ++q) {
            const float* p = a.sp.scratch + (int64_t)(s0 + q) * (C + 4);
            das += __ldcg(p + C);
            add4(acc, __ldcg(reinterpret_cast<const float4*>(p) + sl));
        }
    }
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

extern "C" size_t gd_gat_scratch_floats(const gd_csr_t* csr, int32_t channels) {
    return csr && csr->num_seg > 0 ? (size_t)csr->num_seg * (size_t)(channels + 4) : 0;
}

extern "C" int gd_gat_fwd(const gd_csr_t* csr, const float* h, int64_t ldh, int32_t c, const float* a_src,
                          const float* a_dst, const float* bias, float negative_slope, float* out, int64_t ldo,
                          float* rowmax, float* rowden, float* scratch, gd_stream_t stream_) {
    cudaStream_t stream = as_stream(stream_);
    GD_CHECK_ARG(csr != nullptr, "null csr");
    if (csr->num_rows == 0) return GD_OK;
    GD_CHECK_ARG(csr->rowptr && csr->col && h && a_src && a_dst && out && rowmax && rowden, "null pointer");
    GD_CHECK_ARG(ldh % 4 == 0 && ldo % 4 == 0, "leading dimensions must be multiples of 4");
    GatFwdArgs a{csr->rowptr, csr->col, h, ldh, a_src, a_dst, bias, negative_slope, out, ldo, rowmax, rowden, csr->num_rows,
                 make_split(csr, scratch)};
    GAT_DISPATCH(c, gat_fwd_kernel, a.n + a.sp.num_seg, a);
    return GD_OK;
}

extern "C" int gd_gat_bwd_dst(const gd_csr_t* csr, const int32_t* tinv, const float* h, int64_t ldh, int32_t c,
                              const float* a_src, const float* a_dst, const float* rowmax, const float* rowden,
                              const float* gout, int64_t ldg, const float* out, int64_t ldo, const float* bias,
                              float negative_slope, float* alpha_t, float* dpre_t, float* da_dst, float* scratch,
                              gd_stream_t stream_) {
    cudaStream_t stream = as_stream(stream_);
    GD_CHECK_ARG(csr != nullptr, "null csr");
    if (csr->num_rows == 0) return GD_OK;
    GD_CHECK_ARG(csr->rowptr && csr->col && tinv && h && a_src && a_dst && rowmax && rowden && gout && out &&
                 alpha_t && dpre_t && da_dst, "null pointer");
    GD_CHECK_ARG(ldh % 4 == 0 && ldg % 4 == 0 && ldo % 4 == 0, "leading dimensions must be multiples of 4");
    GatBwdDstArgs a{csr->rowptr, csr->col, tinv, h, ldh, a_src, a_dst, rowmax, rowden, gout, ldg, out, ldo, bias,
                    negative_slope, alpha_t, dpre_t, da_dst, csr->num_rows, make_split(csr, scratch)};
    GAT_DISPATCH(c, gat_bwd_dst_kernel, a.n + a.sp.num_seg, a);
    return GD_OK;
}

extern "C" int gd_gat_bwd_src(const gd_csr_t* csr_t, const float* alpha_t, const float* dpre_t, const float* gout,
                              int64_t ldg, int32_t c, const float* att_src, const float* att_dst,
                              const float* da_dst, float* dh, int64_t lddh, float* da_src, float* scratch,
                              gd_stream_t stream_) {
    cudaStream_t stream = as_stream(stream_);
    GD_CHECK_ARG(csr_t != nullptr, "null csr");
    if (csr_t->num_rows == 0) return GD_OK;
    GD_CHECK_ARG(csr_t->rowptr && csr_t->col && alpha_t && dpre_t && gout && att_src && att_dst && da_dst && dh && da_src,
                 "null pointer");
    GD_CHECK_ARG(ldg % 4 == 0 && lddh % 4 == 0, "leading dimensions must be multiples of 4");
    GatBwdSrcArgs a{csr_t->rowptr, csr_t->col, alpha_t, dpre_t, gout, ldg, att_src, att_dst, da_dst, dh, lddh, da_src,
                    csr_t->num_rows, make_split(csr_t, scratch)};
    GAT_DISPATCH(c, gat_bwd_src_kernel, a.n + a.sp.num_seg, a);
    return GD_OK;
}
