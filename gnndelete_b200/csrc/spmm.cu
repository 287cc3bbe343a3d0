// CSR aggregation (SpMM): the GCN / GIN neighbourhood sum, its transpose-backward and the loss
// backward gather.  A gather kernel: every non-zero reads one 256 B / 512 B source row that is
// (mostly) L2 resident, so the bound is how many row gathers the SMs keep in flight.
//
// Measured on B200 (tools/gather_bench.cu): random 256-byte row gathers from an L2-resident
// matrix sustain ~18 TB/s (7 TB/s from HBM) when ~250 KB of loads are in flight per SM.
// Kernels in this file, Collab shape, F = 64 (tools/spmm_bench.py):
//   spmm_pipe_kernel   123 us  5.4 TB/s gathered   row-pipelined persistent sub-warps
// (the default aggregation is the batched kernel of spmm_batched.cu; this file is the plain-CSR path for per-entry
// values in CSR order, widths the batch plan does not cover and CSRs that are rewritten in place)
// Variants that were measured and dropped: a cp.async ring in shared memory over row groups (152 us), sub-warp per row without cross-row prefetch (131 us +
// 22 us finalize pass), 4 lanes x 4 loads per row (144 us), register streaming over row groups
// (184 us, 100+ registers), forcing 5 CTAs/SM on the pipeline (spills, 133 us).
// Rows longer than seg_len are cut into segments (gd_spmm_plan_build) that are scheduled first;
// the sub-warp that completes the LAST segment of a row (ticket counter) adds the partial sums in
// segment order: no separate finalize pass, no float atomics, bitwise reproducible results.
#include <cstdlib>

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
    int accumulate;          // out += result instead of out = result
};

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
            if (a.accumulate) r += a.out[row * a.ldo + f];
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
        if (a.accumulate) r += a.out[row * a.ldo + f];
        a.out[row * a.ldo + f] = r;
    }
}

// =====================================================================================
// Row-pipelined variant (default for F = 32 / 64 / 128).
//
// Measured on B200 (tools/gather_bench.cu): random 256-byte row gathers from an L2-resident
// matrix sustain ~18 TB/s when ~250 KB of loads are in flight per SM; the kernels above reach
// ~5 TB/s because every row walks rowptr -> col -> gather serially and only the last step has
// many loads outstanding.  Here each sub-warp (F/4 lanes, one 128-bit load per lane and edge) is
// persistent and software-pipelined ACROSS rows: while the gathers of row i are in flight it
// has already issued the column-id load of row i+1 and the rowptr loads of row i+2, so in
// steady state a row costs one memory latency instead of three and the resident warps keep the
// gather queue full.  Long rows are cut into segments (scheduled first, ticket finalize).
template <int LANES, bool WEIGHTED>
__global__ void __launch_bounds__(256) spmm_pipe_kernel(   // (256, 5) would spill at 48 regs and is slower (measured)
    const SpmmArgs a) {
    constexpr int PER_WARP = 32 / LANES;
    constexpr int B = 8;                                       // gathers in flight per lane and batch
    const int lane = threadIdx.x & 31;
    const int sub = lane / LANES, sl = lane % LANES;
    const unsigned mask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << (sub * LANES));
    const int64_t items = (int64_t)a.num_seg + a.num_rows;
    const int64_t G = (int64_t)gridDim.x * (blockDim.x >> 5) * PER_WARP;
    const int64_t gid = ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5) * PER_WARP + sub;
    const float4* xb = reinterpret_cast<const float4*>(a.x) + sl;
    const unsigned ld4 = (unsigned)(a.ldx >> 2);

    // item -> (row, beg, end); row = -1: nothing to do (out of range, or a long row handled by its segments)
    auto fetch = [&](int64_t it, int& row, int& beg, int& end) {
        row = -1; beg = 0; end = 0;
        if (it >= items) return;
        if (it < a.num_seg) {
            row = __ldg(a.seg_row + it);
            beg = __ldg(a.seg_beg + it);
            end = min(beg + a.seg_len, __ldg(a.rowptr + row + 1));
        } else {
            const int r = (int)(it - a.num_seg);
            beg = __ldg(a.rowptr + r);
            end = __ldg(a.rowptr + r + 1);
            row = (a.seg_len > 0 && end - beg > a.seg_len) ? -1 : r;
        }
    };
    auto load_ids = [&](int beg, int end, unsigned& off, float& w) {
        const int k = beg + sl;
        off = 0; w = 0.f;
        if (k < end) {
            const int c = __ldg(a.col + k);
            off = (unsigned)c * ld4;
            if (WEIGHTED) {
                w = a.val ? __ldg(a.val + k) : 1.0f;
                if (a.col_scale) w *= __ldg(a.col_scale + c);
            }
        }
    };

    int64_t it0 = gid, it1 = gid + G;
    int r0, b0, e0, r1, b1, e1;
    fetch(it0, r0, b0, e0);
    fetch(it1, r1, b1, e1);
    unsigned off0; float w0;
    load_ids(b0, r0 >= 0 ? e0 : b0, off0, w0);

    while (it0 < items) {
        // ---- prefetch: rowptr of the item after next, column ids of the next item
        const int64_t it2 = it1 + G;
        int r2, b2, e2;
        fetch(it2, r2, b2, e2);
        unsigned off1; float w1;
        load_ids(b1, r1 >= 0 ? e1 : b1, off1, w1);

        if (r0 >= 0) {
            // ---- gather row / segment it0
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            unsigned off = off0; float w = w0;
            for (int base = b0; base < e0; base += LANES) {
                if (base != b0) load_ids(base, e0, off, w);           // rows longer than LANES: ids on demand
                const int cnt = min(LANES, e0 - base);
                for (int j = 0; j < cnt; j += B) {
                    float4 v[B]; float wj[B];
#pragma unroll
                    for (int u = 0; u < B; ++u) {
                        const int idx = j + u;
                        const unsigned oj = __shfl_sync(mask, off, idx & (LANES - 1), LANES);
                        if (WEIGHTED) wj[u] = __shfl_sync(mask, w, idx & (LANES - 1), LANES);
                        v[u] = idx < cnt ? __ldg(xb + oj) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int u = 0; u < B; ++u) { if (WEIGHTED) fma4(acc, wj[u], v[u]); else add4(acc, v[u]); }
                }
            }
            if (it0 < a.num_seg) {
                // segment: partial to scratch; the sub-warp that arrives last reduces the row's segments in order
                stg4(a.scratch + it0 * (int64_t)a.feat + sl * 4, acc);
                const int h = __ldg(a.seg_heavy + it0);
                const int ns = __ldg(a.heavy_nseg + h);
                __threadfence();
                int ticket = 0;
                if (sl == 0) ticket = atomicAdd(a.heavy_ticket + h, 1);
                ticket = __shfl_sync(mask, ticket, 0, LANES);
                if (ticket == ns - 1) {
                    __threadfence();
                    if (sl == 0) a.heavy_ticket[h] = 0;
                    const int s0 = __ldg(a.heavy_seg_beg + h);
                    acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int s = 0; s < ns; ++s)
                        add4(acc, __ldcg(reinterpret_cast<const float4*>(a.scratch + (int64_t)(s0 + s) * a.feat) + sl));
                } else {
                    r0 = -1;                                            // not the finisher: nothing to store
                }
            }
            if (r0 >= 0) {
                if (a.row_scale) { const float s = __ldg(a.row_scale + r0); acc.x *= s; acc.y *= s; acc.z *= s; acc.w *= s; }
                if (a.self_coef != 0.f) fma4(acc, a.self_coef, __ldg(reinterpret_cast<const float4*>(a.x + (int64_t)r0 * a.ldx) + sl));
                if (a.bias) add4(acc, __ldg(reinterpret_cast<const float4*>(a.bias) + sl));
                float* op = a.out + (int64_t)r0 * a.ldo + sl * 4;
                if (a.accumulate) add4(acc, *reinterpret_cast<const float4*>(op));
                stg4(op, acc);
            }
        }
        // ---- rotate the pipeline
        it0 = it1; r0 = r1; b0 = b1; e0 = e1; off0 = off1; w0 = w1;
        it1 = it2; r1 = r2; b1 = b2; e1 = e2;
    }
}

template <int LANES>
static int launch_pipe(const SpmmArgs& a, bool weighted, cudaStream_t stream) {
    constexpr int PER_WARP = 32 / LANES;
    const int64_t items = (int64_t)a.num_seg + a.num_rows;
    if (items == 0) return GD_OK;
    // persistent grid: all resident CTAs of the chip (occupancy is register bound), capped by the work
    int per_sm = 0;
    if (weighted) GD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spmm_pipe_kernel<LANES, true>, 256, 0));
    else GD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spmm_pipe_kernel<LANES, false>, 256, 0));
    const int64_t want = ceil_div<int64_t>(ceil_div<int64_t>(items, PER_WARP), 8);
    const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)kNumSMs * std::max(per_sm, 1)));
    if (weighted) spmm_pipe_kernel<LANES, true><<<blocks, 256, 0, stream>>>(a);
    else spmm_pipe_kernel<LANES, false><<<blocks, 256, 0, stream>>>(a);
    GD_LAUNCH_CHECK();
    return GD_OK;
}

}  // namespace gd

using namespace gd;

extern "C" int gd_spmm(const gd_csr_t* csr, const float* val, const float* col_scale, const float* row_scale,
                       const float* x, int64_t ldx, int32_t feat, float self_coef, const float* bias,
                       float* out, int64_t ldo, float* scratch, gd_stream_t stream_) {
    return gd_spmm_acc(csr, val, col_scale, row_scale, x, ldx, feat, self_coef, bias, out, ldo, scratch, 0, stream_);
}

extern "C" int gd_spmm_acc(const gd_csr_t* csr, const float* val, const float* col_scale, const float* row_scale,
                           const float* x, int64_t ldx, int32_t feat, float self_coef, const float* bias,
                           float* out, int64_t ldo, float* scratch, int32_t accumulate, gd_stream_t stream_) {
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
    a.accumulate = accumulate;
    const bool weighted = val != nullptr || col_scale != nullptr;
    const bool vec_ok = (ldx % 4 == 0) && (ldo % 4 == 0) &&
                        (((uintptr_t)x | (uintptr_t)out | (uintptr_t)scratch | (uintptr_t)bias) % 16 == 0) &&
                        ((double)csr->num_rows * (double)(ldx / 4) < 4.0e9);   // 32-bit float4 offsets
    if (vec_ok) {
        if (feat == 128) return launch_pipe<32>(a, weighted, stream);
        if (feat == 64) return launch_pipe<16>(a, weighted, stream);
        if (feat == 32) return launch_pipe<8>(a, weighted, stream);
    }
    const int64_t items = a.num_seg + a.num_rows;
    spmm_generic_kernel<<<(unsigned)ceil_div<int64_t>(items, 8), 256, 0, stream>>>(a);
    GD_LAUNCH_CHECK();
    if (a.num_heavy > 0) {
        spmm_generic_finalize_kernel<<<(unsigned)ceil_div<int64_t>(a.num_heavy, 8), 256, 0, stream>>>(a, csr->heavy_row);
        GD_LAUNCH_CHECK();
    }
    return GD_OK;
}
