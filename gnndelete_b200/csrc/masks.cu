// Deletion-mask construction: frontier-bitmap k-hop expansion + induced edge / node masks
// (torch_geometric.utils.k_hop_subgraph as called at delete_gnn.py:128-151) and
// to_undirected + coalesce carrying integer edge attributes (delete_gnn.py:175-182).
// Integer / bit work, HBM-bound: one streaming pass over the edge list per hop, bitmaps
// (N/8 bytes) stay L2-resident.
#include <cub/cub.cuh>

#include "common.cuh"

namespace gd {

__device__ __forceinline__ bool test_bit(const uint32_t* bits, int64_t i) {
    return (bits[i >> 5] >> (i & 31)) & 1u;
}
__device__ __forceinline__ void set_bit(uint32_t* bits, int64_t i) {
    const uint32_t m = 1u << (i & 31);
    if (!(bits[i >> 5] & m)) atomicOr(&bits[i >> 5], m);
}

// seeds: endpoints of the edges selected by sel (delete_gnn.py:129: [:, df_mask].flatten().unique())
__global__ void mark_endpoints_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                      const uint8_t* __restrict__ sel, int64_t E, int64_t N,
                                      uint32_t* __restrict__ bits, uint32_t* __restrict__ bits2,
                                      int32_t* __restrict__ bad) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = src[e], d = dst[e];
        if ((uint64_t)s >= (uint64_t)N || (uint64_t)d >= (uint64_t)N) { atomicAdd(bad, 1); continue; }
        if (!sel[e]) continue;
        set_bit(bits, s); set_bit(bits, d);
        set_bit(bits2, s); set_bit(bits2, d);
    }
}

// one hop with flow='source_to_target': edges whose TARGET is in the frontier add their SOURCE
__global__ void khop_expand_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t E,
                                   int64_t N, const uint32_t* __restrict__ frontier, uint32_t* __restrict__ next,
                                   uint32_t* __restrict__ subset) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = src[e], d = dst[e];
        if ((uint64_t)s >= (uint64_t)N || (uint64_t)d >= (uint64_t)N) continue;
        if (test_bit(frontier, d)) {
            set_bit(next, s);
            set_bit(subset, s);
        }
    }
}

// induced edges of the subset + their endpoints (delete_gnn.py:144-145 node masks)
__global__ void induced_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t E,
                               int64_t N, const uint32_t* __restrict__ subset, uint8_t* __restrict__ edge_mask,
                               uint8_t* __restrict__ node_mask) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = src[e], d = dst[e];
        if ((uint64_t)s >= (uint64_t)N || (uint64_t)d >= (uint64_t)N) { edge_mask[e] = 0; continue; }
        const bool in = test_bit(subset, s) && test_bit(subset, d);
        edge_mask[e] = in ? 1 : 0;
        if (in) { node_mask[s] = 1; node_mask[d] = 1; }
    }
}

// ---- to_undirected ---------------------------------------------------------------------
__global__ void und_keys_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t E,
                                int64_t N, int64_t* __restrict__ keys, int32_t* __restrict__ vals) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < 2 * E; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i < E ? i : i - E;
        const int64_t r = i < E ? src[e] : dst[e];
        const int64_t c = i < E ? dst[e] : src[e];
        keys[i] = r * N + c;
        vals[i] = (int32_t)e;
    }
}

__global__ void und_flag_kernel(const int64_t* __restrict__ keys, int64_t total, int32_t* __restrict__ flag) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// slot[i] = inclusive_scan(flag)[i] - 1; duplicates add their attributes (coalesce reduce='add')
__global__ void und_emit_kernel(const int64_t* __restrict__ keys, const int32_t* __restrict__ vals,
                                const int32_t* __restrict__ slot_incl, int64_t total, int64_t N,
                                const int32_t* __restrict__ attr_a, const int32_t* __restrict__ attr_b,
                                int64_t* __restrict__ out_row, int64_t* __restrict__ out_col,
                                int32_t* __restrict__ out_a, int32_t* __restrict__ out_b,
                                int64_t* __restrict__ out_count) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t slot = slot_incl[i] - 1;
        const int64_t k = keys[i];
        const bool first = (i == 0) || (k != keys[i - 1]);
        if (first) { out_row[slot] = k / N; out_col[slot] = k % N; }
        const int32_t e = vals[i];
        if (attr_a) atomicAdd(&out_a[slot], attr_a[e]);
        if (attr_b) atomicAdd(&out_b[slot], attr_b[e]);
        if (i == total - 1) *out_count = slot + 1;
    }
}

static int grid_for(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div<int64_t>(n, 256), kNumSMs * 16)); }
static size_t bitmap_bytes(int64_t n) { return align_up((size_t)ceil_div<int64_t>(n, 32) * 4); }

}  // namespace gd

using namespace gd;

extern "C" size_t gd_khop_workspace_bytes(int64_t num_nodes) { return 3 * bitmap_bytes(num_nodes) + 256; }

extern "C" int gd_khop_masks(const int64_t* src, const int64_t* dst, int64_t E, int64_t N, const uint8_t* seed_edge_mask,
                             int32_t num_hops, uint8_t* edge_mask, uint8_t* node_mask, int32_t* status,
                             void* workspace, size_t workspace_bytes, gd_stream_t stream_) {
    cudaStream_t stream = as_stream(stream_);
    GD_CHECK_ARG(E >= 0 && N >= 0 && num_hops >= 0, "bad size");
    GD_CHECK_ARG(status != nullptr, "null status");
    GD_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), stream));
    if (N > 0) { GD_CHECK_ARG(node_mask, "null node_mask"); GD_CUDA(cudaMemsetAsync(node_mask, 0, N, stream)); }
    if (E == 0) return GD_OK;
    GD_CHECK_ARG(src && dst && seed_edge_mask && edge_mask, "null pointer");
    if (!workspace || workspace_bytes < gd_khop_workspace_bytes(N))
        return fail(GD_ERR_WORKSPACE, "gd_khop_masks: workspace too small");
    const size_t bb = bitmap_bytes(N);
    char* p = static_cast<char*>(workspace);
    uint32_t* subset = reinterpret_cast<uint32_t*>(p);
    uint32_t* fa = reinterpret_cast<uint32_t*>(p + bb);
    uint32_t* fb = reinterpret_cast<uint32_t*>(p + 2 * bb);
    GD_CUDA(cudaMemsetAsync(p, 0, 3 * bb, stream));
    const int grid = grid_for(E);
    mark_endpoints_kernel<<<grid, 256, 0, stream>>>(src, dst, seed_edge_mask, E, N, subset, fa, status);
    GD_LAUNCH_CHECK();
    for (int h = 0; h < num_hops; ++h) {
        GD_CUDA(cudaMemsetAsync(fb, 0, bb, stream));
        khop_expand_kernel<<<grid, 256, 0, stream>>>(src, dst, E, N, fa, fb, subset);
        GD_LAUNCH_CHECK();
        std::swap(fa, fb);
    }
    induced_kernel<<<grid, 256, 0, stream>>>(src, dst, E, N, subset, edge_mask, node_mask);
    GD_LAUNCH_CHECK();
    return GD_OK;
}

extern "C" size_t gd_to_undirected_workspace_bytes(int64_t E) {
    const int64_t total = 2 * E;
    size_t sort_bytes = 0, scan_bytes = 0;
    cub::DoubleBuffer<int64_t> k(nullptr, nullptr);
    cub::DoubleBuffer<int32_t> v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, k, v, (int)std::min<int64_t>(total, INT32_MAX));
    cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (int32_t*)nullptr, (int32_t*)nullptr, (int)std::min<int64_t>(total, INT32_MAX));
    return align_up(std::max(sort_bytes, scan_bytes)) + 2 * align_up(total * 8) + 4 * align_up(total * 4) + 1024;
}

extern "C" int gd_to_undirected(const int64_t* src, const int64_t* dst, int64_t E, int64_t N, const int32_t* attr_a,
                                const int32_t* attr_b, int64_t* out_row, int64_t* out_col, int32_t* out_a,
                                int32_t* out_b, int64_t* out_count, void* workspace, size_t workspace_bytes,
                                gd_stream_t stream_) {
    cudaStream_t stream = as_stream(stream_);
    GD_CHECK_ARG(E >= 0 && N >= 0, "bad size");
    GD_CHECK_ARG(out_count != nullptr, "null out_count");
    GD_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int64_t), stream));
    if (E == 0) return GD_OK;
    const int64_t total = 2 * E;
    GD_CHECK_ARG(total < INT32_MAX, "too many edges");
    GD_CHECK_ARG(src && dst && out_row && out_col, "null pointer");
    GD_CHECK_ARG((!attr_a || out_a) && (!attr_b || out_b), "attribute without output");
    if (!workspace || workspace_bytes < gd_to_undirected_workspace_bytes(E))
        return fail(GD_ERR_WORKSPACE, "gd_to_undirected: workspace too small");
    char* p = static_cast<char*>(workspace);
    int64_t* k0 = reinterpret_cast<int64_t*>(p); p += align_up(total * 8);
    int64_t* k1 = reinterpret_cast<int64_t*>(p); p += align_up(total * 8);
    int32_t* v0 = reinterpret_cast<int32_t*>(p); p += align_up(total * 4);
    int32_t* v1 = reinterpret_cast<int32_t*>(p); p += align_up(total * 4);
    int32_t* flag = reinterpret_cast<int32_t*>(p); p += align_up(total * 4);
    int32_t* slot = reinterpret_cast<int32_t*>(p); p += align_up(total * 4);
    size_t tmp_bytes = workspace_bytes - (p - static_cast<char*>(workspace));
    const int grid = grid_for(total);
    und_keys_kernel<<<grid, 256, 0, stream>>>(src, dst, E, N, k0, v0);
    GD_LAUNCH_CHECK();
    int bits = 1;
    while (bits < 63 && (int64_t(1) << bits) < N * N) ++bits;
    cub::DoubleBuffer<int64_t> kb(k0, k1);
    cub::DoubleBuffer<int32_t> vb(v0, v1);
    GD_CUDA(cub::DeviceRadixSort::SortPairs(p, tmp_bytes, kb, vb, (int)total, 0, bits, stream));
    und_flag_kernel<<<grid, 256, 0, stream>>>(kb.Current(), total, flag);
    GD_LAUNCH_CHECK();
    GD_CUDA(cub::DeviceScan::InclusiveSum(p, tmp_bytes, flag, slot, (int)total, stream));
    if (attr_a) GD_CUDA(cudaMemsetAsync(out_a, 0, total * sizeof(int32_t), stream));
    if (attr_b) GD_CUDA(cudaMemsetAsync(out_b, 0, total * sizeof(int32_t), stream));
    und_emit_kernel<<<grid, 256, 0, stream>>>(kb.Current(), vb.Current(), slot, total, N, attr_a, attr_b, out_row,
                                              out_col, out_a, out_b, out_count);
    GD_LAUNCH_CHECK();
    return GD_OK;
}
