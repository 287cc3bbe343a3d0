// COO -> CSR construction, GCN normalisation and the long-row split plan.
// One-time setup per edge set (the reference rebuilds the equivalent index state
// inside every GCNConv/GATConv call: gcn_norm with cached=False, gcn.py:11-12).
#include <cub/cub.cuh>

#include "common.cuh"

namespace gd {

// key = ((dst * R + rel) * N + src); dropped entries get key = key_max
__global__ void make_keys_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                 const int64_t* __restrict__ rel, int64_t E, int64_t N, int R,
                                 int self_loops, int64_t key_drop, int64_t* __restrict__ keys,
                                 int32_t* __restrict__ vals, int32_t* __restrict__ status) {
    int64_t total = E + ((self_loops & 1) ? N : 0);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        if (i < E) {
            int64_t s = src[i], d = dst[i];
            int64_t r = rel ? rel[i] : 0;
            bool bad = s < 0 || s >= N || d < 0 || d >= N || r < 0 || r >= R;
            if (bad) atomicAdd(&status[1], 1);
            bool drop = bad || ((self_loops & 1) && s == d);
            // bit 1 of self_loops: order rows only (key = destination): fewer radix passes; the stable sort keeps the
            // input order inside a row, so the result is still deterministic
            keys[i] = drop ? key_drop : ((self_loops & 2) ? d : ((d * R + r) * N + s));
            vals[i] = (int32_t)i;
        } else {
            int64_t v = i - E;
            keys[i] = (v * R) * N + v;
            vals[i] = (int32_t)(E + v);   // inserted self loop of node v
        }
    }
}

// sorted keys -> rowptr / col / rel ; one thread per sorted entry
__global__ void csr_fill_kernel(const int64_t* __restrict__ keys, const int32_t* __restrict__ vals,
                                const int64_t* __restrict__ src_if_rows_only,
                                int64_t total, int64_t N, int R, int64_t key_drop,
                                int32_t* __restrict__ rowptr, int32_t* __restrict__ col,
                                int32_t* __restrict__ eid, int32_t* __restrict__ rel_out,
                                int32_t* __restrict__ status) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= total;
         i += (int64_t)gridDim.x * blockDim.x) {
        // row of entry i (N for the end sentinel / dropped tail)
        int64_t row_i = N, row_p = -1;
        if (i < total) {
            int64_t k = keys[i];
            if (k != key_drop) {
                if (src_if_rows_only) {
                    row_i = k;
                    col[i] = (int32_t)src_if_rows_only[vals[i]];
                    eid[i] = vals[i];
                } else {
                    int64_t dr = k / N;
                    row_i = dr / R;
                    col[i] = (int32_t)(k - dr * N);
                    eid[i] = vals[i];
                    if (rel_out) rel_out[i] = (int32_t)(dr - row_i * R);
                }
            }
        }
        if (i > 0) {
            int64_t kp = keys[i - 1];
            row_p = (kp == key_drop) ? N : (src_if_rows_only ? kp : (kp / N) / R);
        }
        // rows (row_p, row_i] start at i
        if (row_i != row_p) {
            for (int64_t r = row_p + 1; r <= row_i; ++r) rowptr[r] = (int32_t)i;
            if (row_i == N && row_p != N) status[0] = (int32_t)i;   // nnz
        }
    }
}

__global__ void invert_perm_kernel(const int32_t* __restrict__ perm, int64_t n, int32_t* __restrict__ inv) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        int32_t p = perm[i];
        if (p >= 0) inv[p] = (int32_t)i;
    }
}

__global__ void gcn_dinv_kernel(const int32_t* __restrict__ rowptr, int64_t n, float* __restrict__ dinv) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        int deg = rowptr[i + 1] - rowptr[i];
        dinv[i] = deg > 0 ? 1.0f / sqrtf((float)deg) : 0.0f;
    }
}

// one warp per row; heavy rows reserve their segment range with one atomic
__global__ void spmm_plan_kernel(const int32_t* __restrict__ rowptr, int64_t n, int seg_len,
                                 int32_t* __restrict__ heavy_row, int32_t* __restrict__ heavy_seg_beg,
                                 int32_t* __restrict__ heavy_nseg, int32_t* __restrict__ seg_row,
                                 int32_t* __restrict__ seg_beg, int32_t* __restrict__ seg_heavy,
                                 int32_t* __restrict__ counts) {
    int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = warp * 32; base < n; base += nwarps * 32) {
        int64_t row = base + lane;
        int beg = 0, deg = 0;
        if (row < n) { beg = rowptr[row]; deg = rowptr[row + 1] - beg; }
        unsigned heavy = __ballot_sync(0xffffffffu, deg > seg_len);
        while (heavy) {
            int l = __ffs(heavy) - 1;
            heavy &= heavy - 1;
            int r_beg = __shfl_sync(0xffffffffu, beg, l);
            int r_deg = __shfl_sync(0xffffffffu, deg, l);
            int nseg = (r_deg + seg_len - 1) / seg_len;
            int h = 0, s0 = 0;
            if (lane == 0) {
                h = atomicAdd(&counts[0], 1);
                s0 = atomicAdd(&counts[1], nseg);
                heavy_row[h] = (int32_t)(base + l);
                heavy_seg_beg[h] = s0;
                heavy_nseg[h] = nseg;
            }
            s0 = __shfl_sync(0xffffffffu, s0, 0);
            h = __shfl_sync(0xffffffffu, h, 0);
            for (int s = lane; s < nseg; s += 32) {
                seg_row[s0 + s] = (int32_t)(base + l);
                seg_beg[s0 + s] = r_beg + s * seg_len;
                seg_heavy[s0 + s] = h;
            }
        }
    }
}

static int key_bits(int64_t key_drop) {
    int b = 1;
    while (b < 63 && (int64_t(1) << b) <= key_drop) ++b;
    return b;
}

}  // namespace gd

using namespace gd;

extern "C" size_t gd_csr_workspace_bytes(int64_t num_edges, int64_t num_nodes) {
    int64_t total = num_edges + num_nodes;
    size_t sort_bytes = 0;
    cub::DoubleBuffer<int64_t> k(nullptr, nullptr);
    cub::DoubleBuffer<int32_t> v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, k, v, (int)std::min<int64_t>(total, INT32_MAX));
    return align_up(sort_bytes) + 2 * align_up(total * sizeof(int64_t)) + 2 * align_up(total * sizeof(int32_t)) + 1024;
}

extern "C" int gd_csr_from_coo(const int64_t* src, const int64_t* dst, const int64_t* rel, int64_t E,
                               int64_t N, int32_t R, int32_t self_loops, int32_t* rowptr, int32_t* col,
                               int32_t* eid, int32_t* rel_out, int32_t* status, void* workspace,
                               size_t workspace_bytes, gd_stream_t stream_) {
    cudaStream_t stream = as_stream(stream_);
    GD_CHECK_ARG(E >= 0 && N >= 0, "negative size");
    GD_CHECK_ARG(R >= 1, "num_rel must be >= 1");
    GD_CHECK_ARG(rowptr && status, "null output");
    const bool rows_only = (self_loops & 2) != 0;
    GD_CHECK_ARG(!rows_only || (rel == nullptr && R == 1 && !(self_loops & 1)), "row-order-only build: no relations, no self loops");
    int64_t total = E + ((self_loops & 1) ? N : 0);
    GD_CHECK_ARG(total < INT32_MAX, "more than 2^31-1 entries");
    GD_CHECK_ARG((double)N * (double)R * (double)N < 9.0e18, "sort key overflows int64");
    if (workspace_bytes < gd_csr_workspace_bytes(E, N))
        return fail(GD_ERR_WORKSPACE, "gd_csr_from_coo: workspace too small");
    GD_CUDA(cudaMemsetAsync(status, 0, 2 * sizeof(int32_t), stream));
    if (total == 0) {
        GD_CUDA(cudaMemsetAsync(rowptr, 0, (N + 1) * sizeof(int32_t), stream));
        return GD_OK;
    }
    GD_CHECK_ARG(src && dst && col && eid, "null pointer");
    int64_t key_drop = rows_only ? N : N * (int64_t)R * N;   // strictly above every valid key

    char* p = static_cast<char*>(workspace);
    int64_t* keys0 = reinterpret_cast<int64_t*>(p); p += align_up(total * sizeof(int64_t));
    int64_t* keys1 = reinterpret_cast<int64_t*>(p); p += align_up(total * sizeof(int64_t));
    int32_t* vals0 = reinterpret_cast<int32_t*>(p); p += align_up(total * sizeof(int32_t));
    int32_t* vals1 = reinterpret_cast<int32_t*>(p); p += align_up(total * sizeof(int32_t));
    size_t sort_bytes = workspace_bytes - (p - static_cast<char*>(workspace));

    int threads = 256;
    int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(total + 1, threads), kNumSMs * 16);
    make_keys_kernel<<<blocks, threads, 0, stream>>>(src, dst, rel, E, N, R, self_loops, key_drop, keys0, vals0, status);
    GD_LAUNCH_CHECK();
    cub::DoubleBuffer<int64_t> kb(keys0, keys1);
    cub::DoubleBuffer<int32_t> vb(vals0, vals1);
    GD_CUDA(cub::DeviceRadixSort::SortPairs(p, sort_bytes, kb, vb, (int)total, 0, key_bits(key_drop), stream));
    csr_fill_kernel<<<blocks, threads, 0, stream>>>(kb.Current(), vb.Current(), rows_only ? src : nullptr, total, N, R, key_drop,
                                                    rowptr, col, eid, rel_out, status);
    GD_LAUNCH_CHECK();
    return GD_OK;
}

extern "C" int gd_invert_perm(const int32_t* perm, int64_t n, int32_t* inv, gd_stream_t stream) {
    if (n == 0) return GD_OK;
    GD_CHECK_ARG(perm && inv, "null pointer");
    int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(n, 256), kNumSMs * 16);
    invert_perm_kernel<<<blocks, 256, 0, as_stream(stream)>>>(perm, n, inv);
    GD_LAUNCH_CHECK();
    return GD_OK;
}

extern "C" int gd_gcn_dinv(const int32_t* rowptr, int64_t n, float* dinv, gd_stream_t stream) {
    if (n == 0) return GD_OK;
    GD_CHECK_ARG(rowptr && dinv, "null pointer");
    int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(n, 256), kNumSMs * 16);
    gcn_dinv_kernel<<<blocks, 256, 0, as_stream(stream)>>>(rowptr, n, dinv);
    GD_LAUNCH_CHECK();
    return GD_OK;
}

extern "C" int gd_spmm_plan_build(const int32_t* rowptr, int64_t n, int32_t seg_len, int32_t* heavy_row,
                                  int32_t* heavy_seg_beg, int32_t* heavy_nseg, int32_t* seg_row,
                                  int32_t* seg_beg, int32_t* seg_heavy, int32_t* counts, gd_stream_t stream_) {
    cudaStream_t stream = as_stream(stream_);
    GD_CHECK_ARG(counts, "null counts");
    GD_CHECK_ARG(seg_len >= 32, "seg_len must be >= 32");
    GD_CUDA(cudaMemsetAsync(counts, 0, 2 * sizeof(int32_t), stream));
    if (n == 0) return GD_OK;
    GD_CHECK_ARG(rowptr && heavy_row && heavy_seg_beg && heavy_nseg && seg_row && seg_beg && seg_heavy, "null pointer");
    int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(n, 256), kNumSMs * 16);
    spmm_plan_kernel<<<blocks, 256, 0, stream>>>(rowptr, n, seg_len, heavy_row, heavy_seg_beg, heavy_nseg,
                                                 seg_row, seg_beg, seg_heavy, counts);
    GD_LAUNCH_CHECK();
    return GD_OK;
}
