// CSR aggregation (SpMM): the GCN / GIN neighbourhood sum, its transpose-backward and the loss
// backward gather.  A gather kernel: every non-zero reads one 256 B / 512 B source row that is
// (mostly) L2 resident, so the bound is how many row gathers the SMs keep in flight.
//
// Measured on B200 (tools/gather_bench.cu): random 256-byte row gathers from an L2-resident
// matrix sustain ~18 TB/s (7 TB/s from HBM) when ~250 KB of loads are in flight per SM.
// Kernels in this file, Collab shape, F = 64 (tools/spmm_bench.py):
//   spmm_pipe_kernel   (default)  123 us  5.4 TB/s gathered   row-pipelined persistent sub-warps
//   spmm_stream_kernel (GD_SPMM=stream) 152 us              cp.async ring in shared memory
// Variants that were measured and dropped: sub-warp per row without cross-row prefetch (131 us +
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
// Streaming variant (default when the CSR carries a row-group plan).
//
// The sub-warp-per-row kernel above is latency bound: every row walks the dependent chain
// rowptr -> col -> source rows, and only the last link has many loads in flight (ncu: L2 at
// 22 %, DRAM at 13 %, all warps on the long scoreboard).  Here a warp owns a GROUP of
// consecutive rows (~128 non-zeros, built at plan time) and treats its column ids as one
// contiguous stream:
//   * column ids are fetched 32 (16) at a time with one coalesced load, one chunk ahead;
//   * every source row is copied global -> shared memory with cp.async (16 B per lane, no
//     register staging), two chunks (2 x 8 KB) per warp in flight, 12 warps per SM;
//   * the warp then walks the chunk in shared memory, one conflict-free LDS per edge, flushing
//     a destination row whenever the stream crosses a row boundary (row ends held in a
//     register block, broadcast by shuffle).
// Long rows are still cut into segments that are scheduled first and reduced by the warp
// that finishes last (ticket counter), so the result is deterministic.
template <int F>
struct StreamCfg {
    static constexpr int LPE = F / 4;                    // lanes per edge in the async copy (16 B each)
    static constexpr int EPI = 32 / LPE;                 // edges per cp.async instruction
    static constexpr int CHUNK = F >= 128 ? 8 : 16;      // edges per chunk (4 KB for F = 64 / 128)
    static constexpr int VPT = F / 32;                   // floats per lane in the consume phase
    static constexpr int WARPS = 8;
    static constexpr int RING = 2;
    static constexpr int SMEM_BYTES = WARPS * RING * CHUNK * F * 4;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int F, bool WEIGHTED>
__global__ void __launch_bounds__(256) spmm_stream_kernel(const SpmmArgs a, const int32_t* __restrict__ grp_row, int num_grp) {
    using C = StreamCfg<F>;
    extern __shared__ __align__(16) float smem_f[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* ring = smem_f + warp * (C::RING * C::CHUNK * F);
    const int64_t item = (int64_t)blockIdx.x * C::WARPS + warp;
    if (item >= (int64_t)a.num_seg + num_grp) return;

    const bool is_seg = item < a.num_seg;
    int r0, r1, e0, e1;
    if (is_seg) {
        r0 = __ldg(a.seg_row + item); r1 = r0 + 1;
        e0 = __ldg(a.seg_beg + item);
        e1 = min(e0 + a.seg_len, __ldg(a.rowptr + r0 + 1));
    } else {
        const int g = (int)(item - a.num_seg);
        r0 = __ldg(grp_row + g); r1 = __ldg(grp_row + g + 1);
        e0 = __ldg(a.rowptr + r0); e1 = __ldg(a.rowptr + r1);
        if (a.seg_len > 0 && r1 == r0 + 1 && e1 - e0 > a.seg_len) return;      // long row: done by its segments
    }
    const int nchunks = (e1 - e0 + C::CHUNK - 1) / C::CHUNK;

    // column ids (and weights) of a chunk live in one register per lane
    auto load_ids = [&](int c, int& cid, float& w) {
        const int k = e0 + c * C::CHUNK + lane;
        cid = 0; w = 0.f;
        if (lane < C::CHUNK && k < e1) {
            cid = __ldg(a.col + k);
            if (WEIGHTED) {
                w = a.val ? __ldg(a.val + k) : 1.0f;
                if (a.col_scale) w *= __ldg(a.col_scale + cid);
            }
        }
    };
    auto issue = [&](int c, int cid) {                       // async copies of chunk c into ring slot c % RING
        const int cnt = min(C::CHUNK, e1 - (e0 + c * C::CHUNK));
        float* slot = ring + (c % C::RING) * (C::CHUNK * F);
        const int sub = lane / C::LPE, sl = lane % C::LPE;
#pragma unroll
        for (int i = 0; i < C::CHUNK / C::EPI; ++i) {
            const int j = i * C::EPI + sub;
            const int src = __shfl_sync(0xffffffffu, cid, j);
            if (j < cnt) cp_async16(slot + j * F + sl * 4, a.x + (int64_t)src * a.ldx + sl * 4);
        }
        cp_async_commit();
    };

    int cid_cur, cid_nxt = 0;
    float w_cur = 0.f, w_nxt = 0.f, w_r0 = 0.f, w_r1 = 0.f;     // weights of the chunks in ring slot 0 / 1
    load_ids(0, cid_cur, w_cur);
    if (nchunks > 0) { issue(0, cid_cur); w_r0 = w_cur; }
    if (nchunks > 1) { load_ids(1, cid_cur, w_cur); issue(1, cid_cur); w_r1 = w_cur; }
    if (nchunks > 2) load_ids(2, cid_nxt, w_nxt);

    // row ends and row scales of up to 32 rows at a time live in one register per lane, so that
    // finishing a row needs no dependent global load
    int rb = r0;
    int rend = (rb + lane < r1) ? __ldg(a.rowptr + rb + lane + 1) : e1;
    float rscale = (a.row_scale && rb + lane < r1) ? __ldg(a.row_scale + rb + lane) : 1.0f;
    float bias_r[C::VPT];
#pragma unroll
    for (int v = 0; v < C::VPT; ++v) bias_r[v] = a.bias ? __ldg(a.bias + lane * C::VPT + v) : 0.f;
    int row = r0;
    int cur_end = __shfl_sync(0xffffffffu, rend, 0);
    float acc[C::VPT];
#pragma unroll
    for (int v = 0; v < C::VPT; ++v) acc[v] = 0.f;

    auto flush = [&]() {                                     // finish destination row `row`
        if (!is_seg) {
            float o[C::VPT];
            const float s = __shfl_sync(0xffffffffu, rscale, row - rb);
#pragma unroll
            for (int v = 0; v < C::VPT; ++v) {
                o[v] = acc[v] * s;
                if (a.self_coef != 0.f) o[v] = fmaf(a.self_coef, __ldg(a.x + (int64_t)row * a.ldx + lane * C::VPT + v), o[v]);
                o[v] += bias_r[v];
            }
            float* op = a.out + (int64_t)row * a.ldo + lane * C::VPT;
            if (C::VPT == 4) *reinterpret_cast<float4*>(op) = make_float4(o[0], o[1], o[2], o[3]);
            else if (C::VPT == 2) *reinterpret_cast<float2*>(op) = make_float2(o[0], o[1]);
            else op[0] = o[0];
#pragma unroll
            for (int v = 0; v < C::VPT; ++v) acc[v] = 0.f;
        }
        ++row;
        if (row - rb == 32) {
            rb = row;
            rend = (rb + lane < r1) ? __ldg(a.rowptr + rb + lane + 1) : e1;
            rscale = (a.row_scale && rb + lane < r1) ? __ldg(a.row_scale + rb + lane) : 1.0f;
        }
        cur_end = __shfl_sync(0xffffffffu, rend, row - rb);
    };

    for (int c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncwarp();
        const int base = e0 + c * C::CHUNK;
        const int cnt = min(C::CHUNK, e1 - base);
        const float* sp = ring + (c & 1) * (C::CHUNK * F) + lane * C::VPT;
        const float wreg = (c & 1) ? w_r1 : w_r0;
        int j = 0;
        while (j < cnt) {
            // edges up to the next row boundary are accumulated by a tight loop without checks
            int run = cnt - j;
            if (!is_seg) {
                run = min(run, cur_end - (base + j));
                if (run <= 0) { flush(); continue; }
            }
            for (int t0 = 0; t0 < run; t0 += 8) {
                float vv[8][C::VPT], ww[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int t = t0 + u;
                    ww[u] = (t < run) ? (WEIGHTED ? __shfl_sync(0xffffffffu, wreg, (j + t) & 31) : 1.0f) : 0.f;
                    const float* q = sp + (j + min(t, run - 1)) * F;
                    if (C::VPT == 4) {
                        const float4 v = *reinterpret_cast<const float4*>(q);
                        vv[u][0] = v.x; vv[u][1 % C::VPT] = v.y; vv[u][2 % C::VPT] = v.z; vv[u][3 % C::VPT] = v.w;
                    } else if (C::VPT == 2) {
                        const float2 v = *reinterpret_cast<const float2*>(q);
                        vv[u][0] = v.x; vv[u][1 % C::VPT] = v.y;
                    } else {
                        vv[u][0] = q[0];
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u)
#pragma unroll
                    for (int v = 0; v < C::VPT; ++v) acc[v] = fmaf(ww[u], vv[u][v], acc[v]);
            }
            j += run;
        }
        __syncwarp();
        if (c + 2 < nchunks) {                               // refill the slot just consumed
            issue(c + 2, cid_nxt);
            if (c & 1) w_r1 = w_nxt; else w_r0 = w_nxt;
            if (c + 3 < nchunks) load_ids(c + 3, cid_nxt, w_nxt);
        }
    }
    if (!is_seg) {
        while (row < r1) flush();                            // last row and trailing empty rows
        return;
    }
    // ---- segment: partial sum to scratch; the warp that arrives last reduces the row's segments in order
    float* sp = a.scratch + item * (int64_t)F + lane * C::VPT;
#pragma unroll
    for (int v = 0; v < C::VPT; ++v) sp[v] = acc[v];
    const int h = __ldg(a.seg_heavy + item);
    const int ns = __ldg(a.heavy_nseg + h);
    __threadfence();
    int ticket = 0;
    if (lane == 0) ticket = atomicAdd(a.heavy_ticket + h, 1);
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    if (ticket != ns - 1) return;
    __threadfence();
    if (lane == 0) a.heavy_ticket[h] = 0;
    const int s0 = __ldg(a.heavy_seg_beg + h);
#pragma unroll
    for (int v = 0; v < C::VPT; ++v) acc[v] = 0.f;
    for (int s = 0; s < ns; ++s) {
        const float* p = a.scratch + (int64_t)(s0 + s) * F + lane * C::VPT;
#pragma unroll
        for (int v = 0; v < C::VPT; ++v) acc[v] += __ldcg(p + v);
    }
    {
        const int64_t rw = r0;
        const float s = a.row_scale ? __ldg(a.row_scale + rw) : 1.0f;
#pragma unroll
        for (int v = 0; v < C::VPT; ++v) {
            float o = acc[v] * s;
            if (a.self_coef != 0.f) o = fmaf(a.self_coef, __ldg(a.x + rw * a.ldx + lane * C::VPT + v), o);
            if (a.bias) o += __ldg(a.bias + lane * C::VPT + v);
            a.out[rw * a.ldo + lane * C::VPT + v] = o;
        }
    }
}

template <int F>
static int launch_stream(const SpmmArgs& a, const int32_t* grp_row, int num_grp, bool weighted, cudaStream_t stream) {
    using C = StreamCfg<F>;
    const int64_t items = (int64_t)a.num_seg + num_grp;
    const unsigned blocks = (unsigned)ceil_div<int64_t>(items, C::WARPS);
    if (blocks == 0) return GD_OK;
    if (weighted) {
        GD_CUDA(cudaFuncSetAttribute(spmm_stream_kernel<F, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        spmm_stream_kernel<F, true><<<blocks, 32 * C::WARPS, C::SMEM_BYTES, stream>>>(a, grp_row, num_grp);
    } else {
        GD_CUDA(cudaFuncSetAttribute(spmm_stream_kernel<F, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        spmm_stream_kernel<F, false><<<blocks, 32 * C::WARPS, C::SMEM_BYTES, stream>>>(a, grp_row, num_grp);
    }
    GD_LAUNCH_CHECK();
    return GD_OK;
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
    // GD_SPMM = pipe (default) | stream selects the aggregation kernel (for A/B measurements)
    static const int mode = [] {
        const char* e = getenv("GD_SPMM");
        return (e && e[0] == 's') ? 1 : 0;
    }();
    if (vec_ok && mode == 0) {
        if (feat == 128) return launch_pipe<32>(a, weighted, stream);
        if (feat == 64) return launch_pipe<16>(a, weighted, stream);
        if (feat == 32) return launch_pipe<8>(a, weighted, stream);
    }
    if (vec_ok && mode == 1 && !accumulate && csr->grp_row && csr->num_grp > 0) {
        if (feat == 128) return launch_stream<128>(a, csr->grp_row, csr->num_grp, weighted, stream);
        if (feat == 64) return launch_stream<64>(a, csr->grp_row, csr->num_grp, weighted, stream);
        if (feat == 32) return launch_stream<32>(a, csr->grp_row, csr->num_grp, weighted, stream);
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
