// CSR aggregation (SpMM): the GCN / GIN neighbourhood sum, its transpose-backward and
// the loss backward gather.  Gather kernel bound by L2 / HBM bandwidth once the instruction
// stream is thin enough, so the design minimises warp-instructions per gathered row:
//   * LANES lanes own one destination row and each lane reads VPL consecutive 128-bit words
//     of every source row (F=64: 4 lanes x 64 B, F=128: 8 lanes x 64 B), so a warp works on
//     8 / 4 destination rows at once and one index shuffle + one address computation are
//     amortised over VPL loads;
//   * column ids are fetched LANES at a time with one coalesced load and broadcast by
//     shuffle; 4 edges x VPL loads (16 x 16 B) are in flight per lane;
//   * rows can be visited through a (windowed, degree-sorted) permutation so that the rows
//     sharing a warp have similar length;
//   * rows longer than seg_len are cut into segments (gd_spmm_plan_build) that are scheduled
//     first; the sub-warp that completes the LAST segment of a row (ticket counter) adds the
//     partial sums in segment order, so there is no separate finalize pass, no float atomics
//     and the result is bitwise reproducible.
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
    const int32_t* row_perm;
    const int32_t* seg_row;
    const int32_t* seg_beg;
    const int32_t* seg_heavy;
    const int32_t* heavy_seg_beg;
    const int32_t* heavy_nseg;
    int32_t* heavy_ticket;
    int64_t ldx, ldo;
    int64_t num_rows;
    int32_t num_seg, num_heavy, seg_len, feat;
    float self_coef;
};

template <int VPL>
__device__ __forceinline__ void zero_acc(float4 (&acc)[VPL]) {
#pragma unroll
    for (int q = 0; q < VPL; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
}

template <int LANES, int VPL, bool WEIGHTED>
__device__ __forceinline__ void gather_range(const SpmmArgs& a, int beg, int end, int sl, unsigned mask,
                                             float4 (&acc)[VPL]) {
    const float4* xb = reinterpret_cast<const float4*>(a.x) + sl * VPL;
    const unsigned ld4 = (unsigned)(a.ldx >> 2);
    for (int base = beg; base < end; base += LANES) {
        const int k = base + sl;
        unsigned off = 0;
        float w = 0.f;
        if (k < end) {
            const int c = __ldg(a.col + k);
            off = (unsigned)c * ld4;
            if (WEIGHTED) {
                w = a.val ? __ldg(a.val + k) : 1.0f;
                if (a.col_scale) w *= __ldg(a.col_scale + c);
            }
        }
        const int cnt = min(LANES, end - base);
#pragma unroll
        for (int j0 = 0; j0 < LANES; j0 += 4) {
            if (j0 < cnt) {
                float4 v[4][VPL];
                float wj[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int idx = j0 + u;                       // < LANES when LANES >= 4
                    const unsigned oj = __shfl_sync(mask, off, idx & (LANES - 1), LANES);
                    if (WEIGHTED) wj[u] = __shfl_sync(mask, w, idx & (LANES - 1), LANES);
                    const bool valid = idx < cnt;
#pragma unroll
                    for (int q = 0; q < VPL; ++q)
                        v[u][q] = valid ? __ldg(xb + oj + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int q = 0; q < VPL; ++q) {
                        if (WEIGHTED) fma4(acc[q], wj[u], v[u][q]); else add4(acc[q], v[u][q]);
                    }
            }
        }
    }
}

template <int VPL>
__device__ __forceinline__ void finish_row(const SpmmArgs& a, int64_t row, int sl, float4 (&acc)[VPL]) {
    if (a.row_scale) {
        const float s = __ldg(a.row_scale + row);
#pragma unroll
        for (int q = 0; q < VPL; ++q) { acc[q].x *= s; acc[q].y *= s; acc[q].z *= s; acc[q].w *= s; }
    }
    if (a.self_coef != 0.f) {
        const float4* xs = reinterpret_cast<const float4*>(a.x + row * a.ldx) + sl * VPL;
#pragma unroll
        for (int q = 0; q < VPL; ++q) fma4(acc[q], a.self_coef, __ldg(xs + q));
    }
    if (a.bias) {
        const float4* bp = reinterpret_cast<const float4*>(a.bias) + sl * VPL;
#pragma unroll
        for (int q = 0; q < VPL; ++q) add4(acc[q], __ldg(bp + q));
    }
    float4* op = reinterpret_cast<float4*>(a.out + row * a.ldo) + sl * VPL;
#pragma unroll
    for (int q = 0; q < VPL; ++q) op[q] = acc[q];
}

// items [0, num_seg) are long-row segments (scheduled first), items [num_seg, num_seg+N) are rows
template <int LANES, int VPL, bool WEIGHTED>
__global__ void __launch_bounds__(256) spmm_vec_kernel(const SpmmArgs a) {
    constexpr int PER_WARP = 32 / LANES;
    const int lane = threadIdx.x & 31;
    const int sub = lane / LANES, sl = lane % LANES;
    const unsigned mask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << (sub * LANES));
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t item = warp * PER_WARP + sub;
    if (item >= a.num_seg + a.num_rows) return;
    float4 acc[VPL];
    zero_acc<VPL>(acc);
    if (item < a.num_seg) {
        const int row = __ldg(a.seg_row + item);
        const int beg = __ldg(a.seg_beg + item);
        const int end = min(beg + a.seg_len, __ldg(a.rowptr + row + 1));
        gather_range<LANES, VPL, WEIGHTED>(a, beg, end, sl, mask, acc);
        float4* sp = reinterpret_cast<float4*>(a.scratch + item * (int64_t)a.feat) + sl * VPL;
#pragma unroll
        for (int q = 0; q < VPL; ++q) sp[q] = acc[q];
        // the sub-warp that stores the last segment of the row reduces all of them, in order
        const int h = __ldg(a.seg_heavy + item);
        const int ns = __ldg(a.heavy_nseg + h);
        __threadfence();
        int ticket = 0;
        if (sl == 0) ticket = atomicAdd(a.heavy_ticket + h, 1);
        ticket = __shfl_sync(mask, ticket, 0, LANES);
        if (ticket != ns - 1) return;
        __threadfence();
        if (sl == 0) a.heavy_ticket[h] = 0;                      // re-arm for the next launch
        const int s0 = __ldg(a.heavy_seg_beg + h);
        zero_acc<VPL>(acc);
        const float4* sb = reinterpret_cast<const float4*>(a.scratch) + sl * VPL;
        const int f4 = a.feat >> 2;
        for (int s = 0; s < ns; ++s) {
            const float4* p = sb + (int64_t)(s0 + s) * f4;
#pragma unroll
            for (int q = 0; q < VPL; ++q) add4(acc[q], __ldcg(p + q));   // L2: written by other SMs
        }
        finish_row<VPL>(a, row, sl, acc);
        return;
    }
    const int64_t i = item - a.num_seg;
    const int64_t row = a.row_perm ? __ldg(a.row_perm + i) : i;
    const int beg = __ldg(a.rowptr + row), end = __ldg(a.rowptr + row + 1);
    if (a.seg_len > 0 && end - beg > a.seg_len) return;          // finished from its segments
    gather_range<LANES, VPL, WEIGHTED>(a, beg, end, sl, mask, acc);
    finish_row<VPL>(a, row, sl, acc);
}

// any feature width / alignment: one warp per item, 32-wide scalar strips (coalesced 128 B)
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

__global__ void __launch_bounds__(256) spmm_generic_finalize_kernel(const SpmmArgs a, const int32_t* heavy_row) {
    const int lane = threadIdx.x & 31;
    const int64_t h = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (h >= a.num_heavy) return;
    const int64_t row = heavy_row[h];
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

template <int LANES, int VPL>
static int launch_vec(const SpmmArgs& a, bool weighted, cudaStream_t stream) {
    constexpr int PER_WARP = 32 / LANES;
    const int64_t items = a.num_seg + a.num_rows;
    const int64_t blocks = ceil_div<int64_t>(ceil_div<int64_t>(items, PER_WARP), 8);
    if (blocks > 0) {
        if (weighted) spmm_vec_kernel<LANES, VPL, true><<<(unsigned)blocks, 256, 0, stream>>>(a);
        else spmm_vec_kernel<LANES, VPL, false><<<(unsigned)blocks, 256, 0, stream>>>(a);
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
    GD_CHECK_ARG(csr->num_seg == 0 || (scratch && csr->seg_row && csr->seg_beg && csr->seg_heavy && csr->heavy_ticket &&
                                       csr->seg_len > 0), "split plan without scratch / ticket arrays");
    GD_CHECK_ARG(csr->num_heavy == 0 || (csr->heavy_row && csr->heavy_seg_beg && csr->heavy_nseg), "incomplete split plan");
    SpmmArgs a;
    a.rowptr = csr->rowptr; a.col = csr->col; a.val = val; a.col_scale = col_scale; a.row_scale = row_scale;
    a.x = x; a.bias = bias; a.out = out; a.scratch = scratch; a.row_perm = csr->row_perm;
    a.seg_row = csr->seg_row; a.seg_beg = csr->seg_beg; a.seg_heavy = csr->seg_heavy;
    a.heavy_seg_beg = csr->heavy_seg_beg; a.heavy_nseg = csr->heavy_nseg; a.heavy_ticket = csr->heavy_ticket;
    a.ldx = ldx; a.ldo = ldo; a.num_rows = csr->num_rows;
    a.num_seg = csr->num_seg; a.num_heavy = csr->num_heavy; a.seg_len = csr->seg_len; a.feat = feat;
    a.self_coef = self_coef;
    const bool weighted = val != nullptr || col_scale != nullptr;
    const bool vec_ok = (ldx % 4 == 0) && (ldo % 4 == 0) &&
                        (((uintptr_t)x | (uintptr_t)out | (uintptr_t)scratch | (uintptr_t)bias) % 16 == 0) &&
                        ((double)csr->num_rows * (double)(ldx / 4) < 4.0e9);   // 32-bit float4 offsets
    if (vec_ok && feat == 128) return launch_vec<8, 4>(a, weighted, stream);
    if (vec_ok && feat == 64) return launch_vec<4, 4>(a, weighted, stream);
    if (vec_ok && feat == 32) return launch_vec<4, 2>(a, weighted, stream);
    const int64_t items = a.num_seg + a.num_rows;
    spmm_generic_kernel<<<(unsigned)ceil_div<int64_t>(items, 8), 256, 0, stream>>>(a);
    GD_LAUNCH_CHECK();
    if (a.num_heavy > 0) {
        spmm_generic_finalize_kernel<<<(unsigned)ceil_div<int64_t>(a.num_heavy, 8), 256, 0, stream>>>(a, csr->heavy_row);
        GD_LAUNCH_CHECK();
    }
    return GD_OK;
}
