// CSR aggregation (SpMM): the GCN / GIN neighbourhood sum, its transpose-backward and
// the loss backward gather.  HBM/L2-bound gather kernel:
//   * a sub-warp of F/4 lanes owns one destination row and reads each source row as
//     one coalesced run of 128-bit loads (F=64: 256 B, F=128: 512 B);
//   * column ids of a row are fetched with one coalesced load per LANES entries and
//     broadcast by shuffle, 8 independent row gathers are kept in flight per lane;
//   * rows longer than seg_len are cut into segments (gd_spmm_plan_build) that are
//     scheduled first and reduced by a small finalize kernel, so a power-law hub
//     never serialises on one warp and the result stays deterministic (no atomics).
#include "common.cuh"

namespace gd {

struct SpmmArgs {
    const int32_t* rowptr;
    const int32_t* col;
    const float* val;
    const float* col_scale;
    const float* row_scale;
    const float* x;
    const float* bias;
    float* out;
    float* scratch;
    const int32_t* seg_row;
    const int32_t* seg_beg;
    const int32_t* heavy_row;
    const int32_t* heavy_seg_beg;
    const int32_t* heavy_nseg;
    int64_t ldx, ldo;
    int64_t num_rows;
    int32_t num_seg, num_heavy, seg_len, feat;
    float self_coef;
};

constexpr int kBatch = 8;   // independent source-row gathers in flight per lane

template <int LANES, bool WEIGHTED>
__device__ __forceinline__ float4 gather_range(const SpmmArgs& a, int beg, int end, int sl, unsigned mask) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* xb = a.x + sl * 4;
    for (int base = beg; base < end; base += LANES) {
        int k = base + sl;
        int c = 0;
        float w = 0.f;
        if (k < end) {
            c = __ldg(a.col + k);
            if (WEIGHTED) {
                w = a.val ? __ldg(a.val + k) : 1.0f;
                if (a.col_scale) w *= __ldg(a.col_scale + c);
            }
        }
        int cnt = min(LANES, end - base);
        for (int j = 0; j < cnt; j += kBatch) {
            float4 v[kBatch];
            float wj[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                int idx = j + u;
                int cj = __shfl_sync(mask, c, idx & (LANES - 1), LANES);
                if (WEIGHTED) wj[u] = __shfl_sync(mask, w, idx & (LANES - 1), LANES);
                v[u] = (idx < cnt) ? ldg4(xb + (int64_t)cj * a.ldx) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                if (WEIGHTED) fma4(acc, wj[u], v[u]); else add4(acc, v[u]);
            }
        }
    }
    return acc;
}

// items [0, num_seg) are long-row segments (scheduled first), items [num_seg, num_seg+N) are rows
template <int LANES, bool WEIGHTED>
__global__ void __launch_bounds__(256) spmm_vec_kernel(const SpmmArgs a) {
    constexpr int PER_WARP = 32 / LANES;
    const int lane = threadIdx.x & 31;
    const int sub = lane / LANES, sl = lane % LANES;
    const unsigned mask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << (sub * LANES));
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t item = warp * PER_WARP + sub;
    if (item >= a.num_seg + a.num_rows) return;
    if (item < a.num_seg) {
        int row = a.seg_row[item];
        int beg = a.seg_beg[item];
        int end = min(beg + a.seg_len, __ldg(a.rowptr + row + 1));
        float4 acc = gather_range<LANES, WEIGHTED>(a, beg, end, sl, mask);
        stg4(a.scratch + item * (int64_t)a.feat + sl * 4, acc);
        return;
    }
    const int64_t row = item - a.num_seg;
    int beg = __ldg(a.rowptr + row), end = __ldg(a.rowptr + row + 1);
    if (a.seg_len > 0 && end - beg > a.seg_len) return;   // finalised from segments
    float4 acc = gather_range<LANES, WEIGHTED>(a, beg, end, sl, mask);
    if (a.row_scale) { float s = __ldg(a.row_scale + row); acc.x *= s; acc.y *= s; acc.z *= s; acc.w *= s; }
    if (a.self_coef != 0.f) fma4(acc, a.self_coef, ldg4(a.x + row * a.ldx + sl * 4));
    if (a.bias) add4(acc, ldg4(a.bias + sl * 4));
    stg4(a.out + row * a.ldo + sl * 4, acc);
}

template <int LANES>
__global__ void __launch_bounds__(256) spmm_vec_finalize_kernel(const SpmmArgs a) {
    constexpr int PER_WARP = 32 / LANES;
    const int lane = threadIdx.x & 31;
    const int sub = lane / LANES, sl = lane % LANES;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t h = warp * PER_WARP + sub;
    if (h >= a.num_heavy) return;
    const int64_t row = a.heavy_row[h];
    const int s0 = a.heavy_seg_beg[h], ns = a.heavy_nseg[h];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < ns; ++s) add4(acc, ldg4(a.scratch + (int64_t)(s0 + s) * a.feat + sl * 4));
    if (a.row_scale) { float s = __ldg(a.row_scale + row); acc.x *= s; acc.y *= s; acc.z *= s; acc.w *= s; }
    if (a.self_coef != 0.f) fma4(acc, a.self_coef, ldg4(a.x + row * a.ldx + sl * 4));
    if (a.bias) add4(acc, ldg4(a.bias + sl * 4));
    stg4(a.out + row * a.ldo + sl * 4, acc);
}

// any feature width: one warp per item, 32-wide scalar strips (coalesced 128 B per source row strip)
__global__ void __launch_bounds__(256) spmm_generic_kernel(const SpmmArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t item = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (item >= a.num_seg + a.num_rows) return;
    const bool is_seg = item < a.num_seg;
    int64_t row;
    int beg, end;
    if (is_seg) {
        row = a.seg_row[item];
        beg = a.seg_beg[item];
        end = min(beg + a.seg_len, a.rowptr[row + 1]);
    } else {
        row = item - a.num_seg;
        beg = a.rowptr[row]; end = a.rowptr[row + 1];
        if (a.seg_len > 0 && end - beg > a.seg_len) return;
    }
    for (int f0 = 0; f0 < a.feat; f0 += 128) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int base = beg; base < end; base += 32) {
            int k = base + lane;
            int c = 0; float w = 0.f;
            if (k < end) {
                c = a.col[k];
                w = a.val ? a.val[k] : 1.0f;
                if (a.col_scale) w *= a.col_scale[c];
            }
            int cnt = min(32, end - base);
            for (int j = 0; j < cnt; ++j) {
                int cj = __shfl_sync(0xffffffffu, c, j);
                float wj = __shfl_sync(0xffffffffu, w, j);
                const float* xr = a.x + (int64_t)cj * a.ldx;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    int f = f0 + t * 32 + lane;
                    if (f < a.feat) acc[t] = fmaf(wj, __ldg(xr + f), acc[t]);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            int f = f0 + t * 32 + lane;
            if (f >= a.feat) continue;
            if (is_seg) { a.scratch[item * (int64_t)a.feat + f] = acc[t]; continue; }
            float r = acc[t];
            if (a.row_scale) r *= a.row_scale[row];
            if (a.self_coef != 0.f) r = fmaf(a.self_coef, a.x[row * a.ldx + f], r);
            if (a.bias) r += a.bias[f];
            a.out[row * a.ldo + f] = r;
        }
    }
}

__global__ void __launch_bounds__(256) spmm_generic_finalize_kernel(const SpmmArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t h = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (h >= a.num_heavy) return;
    const int64_t row = a.heavy_row[h];
    const int s0 = a.heavy_seg_beg[h], ns = a.heavy_nseg[h];
    for (int f = lane; f < a.feat; f += 32) {
        float r = 0.f;
        for (int s = 0; s < ns; ++s) r += a.scratch[(int64_t)(s0 + s) * a.feat + f];
        if (a.row_scale) r *= a.row_scale[row];
        if (a.self_coef != 0.f) r = fmaf(a.self_coef, a.x[row * a.ldx + f], r);
        if (a.bias) r += a.bias[f];
        a.out[row * a.ldo + f] = r;
    }
}

template <int LANES>
static int launch_vec(const SpmmArgs& a, bool weighted, cudaStream_t stream) {
    constexpr int PER_WARP = 32 / LANES;
    int64_t items = a.num_seg + a.num_rows;
    int64_t warps = ceil_div<int64_t>(items, PER_WARP);
    int64_t blocks = ceil_div<int64_t>(warps, 8);
    if (blocks > 0) {
        if (weighted) spmm_vec_kernel<LANES, true><<<(unsigned)blocks, 256, 0, stream>>>(a);
        else spmm_vec_kernel<LANES, false><<<(unsigned)blocks, 256, 0, stream>>>(a);
        GD_LAUNCH_CHECK();
    }
    if (a.num_heavy > 0) {
        int64_t hb = ceil_div<int64_t>(ceil_div<int64_t>(a.num_heavy, PER_WARP), 8);
        spmm_vec_finalize_kernel<LANES><<<(unsigned)hb, 256, 0, stream>>>(a);
        GD_LAUNCH_CHECK();
    }
    return GD_OK;
}

}  // namespace gd

using namespace gd;

extern "C" int gd_spmm(const gd_csr_t* csr, const float* val, const float* col_scale, const float* row_scale,
                       const float* x, int64_t ldx, int32_t feat, float self_coef, const float* bias,
                       float* out, int64_t ldo, float* scratch, gd_stream_t stream_) {
    cudaStream_t stream = as_stream(stream_);
    GD_CHECK_ARG(csr != nullptr, "null csr");
    GD_CHECK_ARG(feat > 0, "feat must be positive");
    if (csr->num_rows == 0) return GD_OK;
    GD_CHECK_ARG(csr->rowptr && x && out, "null pointer");
    GD_CHECK_ARG(csr->nnz == 0 || csr->col, "null col");
    GD_CHECK_ARG(ldx >= feat && ldo >= feat, "leading dimension smaller than feat");
    GD_CHECK_ARG(csr->num_seg == 0 || (scratch && csr->seg_row && csr->seg_beg && csr->seg_len > 0), "split plan without scratch");
    GD_CHECK_ARG(csr->num_heavy == 0 || (csr->heavy_row && csr->heavy_seg_beg && csr->heavy_nseg), "incomplete split plan");
    SpmmArgs a;
    a.rowptr = csr->rowptr; a.col = csr->col; a.val = val; a.col_scale = col_scale; a.row_scale = row_scale;
    a.x = x; a.bias = bias; a.out = out; a.scratch = scratch;
    a.seg_row = csr->seg_row; a.seg_beg = csr->seg_beg; a.heavy_row = csr->heavy_row;
    a.heavy_seg_beg = csr->heavy_seg_beg; a.heavy_nseg = csr->heavy_nseg;
    a.ldx = ldx; a.ldo = ldo; a.num_rows = csr->num_rows;
    a.num_seg = csr->num_seg; a.num_heavy = csr->num_heavy; a.seg_len = csr->seg_len; a.feat = feat;
    a.self_coef = self_coef;
    const bool weighted = val != nullptr || col_scale != nullptr;
    const bool vec_ok = (ldx % 4 == 0) && (ldo % 4 == 0) && (((uintptr_t)x | (uintptr_t)out | (uintptr_t)scratch | (uintptr_t)bias) % 16 == 0);
    if (vec_ok && feat == 128) return launch_vec<32>(a, weighted, stream);
    if (vec_ok && feat == 64) return launch_vec<16>(a, weighted, stream);
    if (vec_ok && feat == 32) return launch_vec<8>(a, weighted, stream);
    int64_t items = a.num_seg + a.num_rows;
    int64_t blocks = ceil_div<int64_t>(items, 8);
    spmm_generic_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a);
    GD_LAUNCH_CHECK();
    if (a.num_heavy > 0) {
        spmm_generic_finalize_kernel<<<(unsigned)ceil_div<int64_t>(a.num_heavy, 8), 256, 0, stream>>>(a);
        GD_LAUNCH_CHECK();
    }
    return GD_OK;
}
