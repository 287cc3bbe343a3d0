// Native builder of the batch plan of gd_spmm_batched (layout: include/gnndelete_b200.h, gd_spmm_bplan_t).
// Two calls, because the array sizes depend on the degree distribution:
//   gd_spmm_bplan_count : per-row batch counts -> exclusive scans (CUB) -> {num_batches, num_split, num_piece}
//                         on the device + the per-row scan arrays the fill pass needs (kept in the workspace);
//   gd_spmm_bplan_fill  : one warp per row writes the padded column slots, the descriptors, slot_of_entry and
//                         the split-row tables.
// One-time setup per edge set (plan time, not on the epoch path).
#include <cub/cub.cuh>

#include "common.cuh"

namespace gd {

constexpr int kSlots = 8;

// per row: number of batches, and (given the worker range length) whether the row is split and into how many pieces
__global__ void bplan_rows_kernel(const int32_t* __restrict__ rowptr, int64_t n, int32_t* __restrict__ nbr) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const int deg = rowptr[r + 1] - rowptr[r];
        nbr[r] = max((deg + kSlots - 1) / kSlots, 1);               // an empty row is one all-padding batch
    }
}

__global__ void bplan_split_kernel(const int32_t* __restrict__ bptr, int64_t n, const int32_t* __restrict__ total,
                                   int32_t num_workers, int32_t* __restrict__ split, int32_t* __restrict__ npiece) {
    const int nb = *total;
    const int w = max(1, min(num_workers, max(nb, 1)));
    const int per = nb ? (nb + w - 1) / w : 1;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const int first = bptr[r], last = bptr[r + 1] - 1;
        const int wf = first / per, wl = last / per;
        split[r] = wf != wl;
        npiece[r] = wf != wl ? wl - wf + 1 : 0;
    }
}

__global__ void bplan_sizes_kernel(const int32_t* __restrict__ bptr, const int32_t* __restrict__ hid,
                                   const int32_t* __restrict__ pbeg, int64_t n, int32_t* __restrict__ sizes) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { sizes[0] = bptr[n]; sizes[1] = hid[n]; sizes[2] = pbeg[n]; }
}

struct FillArgs {
    const int32_t* rowptr; const int32_t* col; const int32_t* bptr; const int32_t* hid; const int32_t* pbeg;
    int64_t n; int32_t per;
    int32_t* desc; int32_t* colp; int64_t* slot_of_entry;
    int32_t* piece_split; int32_t* split_row; int32_t* split_piece_beg; int32_t* split_npiece;
};

__global__ void __launch_bounds__(256) bplan_fill_kernel(const FillArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; r < a.n; r += warps) {
        const int e0 = a.rowptr[r], e1 = a.rowptr[r + 1];
        const int first = a.bptr[r], nb = a.bptr[r + 1] - first;
        const int wf = first / a.per, wl = (first + nb - 1) / a.per;
        const bool split = wf != wl;
        const int h = a.hid[r], p0 = a.pbeg[r];
        if (split && lane == 0) { a.split_row[h] = (int32_t)r; a.split_piece_beg[h] = p0; a.split_npiece[h] = wl - wf + 1; }
        for (int k = lane; k < nb; k += 32) {
            const int b = first + k;
            const int eb = e0 + kSlots * k;
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
                const int e = eb + s;
                const bool ok = e < e1;
                a.colp[(int64_t)b * kSlots + s] = ok ? a.col[e] : -1;
                if (ok) a.slot_of_entry[e] = (int64_t)b * kSlots + s;
            }
            const bool flush = k == nb - 1 || (b + 1) % a.per == 0;
            int d = 0;
            if (flush) {
                if (!split) d = (int)(0x80000000u | (uint32_t)r);
                else {
                    const int piece = p0 + (b / a.per - wf);
                    d = (int)(0x80000000u | 0x40000000u | (uint32_t)piece);
                    a.piece_split[piece] = h;
                }
            }
            a.desc[b] = d;
        }
    }
}

static size_t bplan_scan_bytes(int64_t n) {
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr, (int)std::min<int64_t>(n + 1, INT32_MAX));
    return bytes;
}

}  // namespace gd

using namespace gd;

// workspace layout: [nbr | bptr | split | hid | npiece | pbeg] (n + 1 int32 each) + CUB temp storage
extern "C" size_t gd_spmm_bplan_workspace_bytes(int64_t num_rows) {
    return 6 * align_up((size_t)(num_rows + 1) * sizeof(int32_t)) + align_up(bplan_scan_bytes(num_rows)) + 256;
}

extern "C" int gd_spmm_bplan_count(const int32_t* rowptr, int64_t num_rows, int32_t num_workers, int32_t* sizes,
                                   void* workspace, size_t workspace_bytes, gd_stream_t stream_) {
    cudaStream_t stream = as_stream(stream_);
    GD_CHECK_ARG(rowptr && sizes && workspace, "null pointer");
    GD_CHECK_ARG(num_rows > 0 && num_rows < (1 << 30), "row count out of range");
    GD_CHECK_ARG(num_workers > 0, "num_workers must be positive");
    if (workspace_bytes < gd_spmm_bplan_workspace_bytes(num_rows)) return fail(GD_ERR_WORKSPACE, "gd_spmm_bplan_count: workspace too small");
    const size_t arr = align_up((size_t)(num_rows + 1) * sizeof(int32_t));
    char* p = static_cast<char*>(workspace);
    int32_t* nbr = reinterpret_cast<int32_t*>(p);
    int32_t* bptr = reinterpret_cast<int32_t*>(p + arr);
    int32_t* split = reinterpret_cast<int32_t*>(p + 2 * arr);
    int32_t* hid = reinterpret_cast<int32_t*>(p + 3 * arr);
    int32_t* npiece = reinterpret_cast<int32_t*>(p + 4 * arr);
    int32_t* pbeg = reinterpret_cast<int32_t*>(p + 5 * arr);
    void* tmp = p + 6 * arr;
    size_t tmp_bytes = bplan_scan_bytes(num_rows);
    const int n1 = (int)(num_rows + 1);
    const int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(num_rows, 256), kNumSMs * 16);
    GD_CUDA(cudaMemsetAsync(workspace, 0, 6 * arr, stream));           // the (n + 1)-th inputs of the scans are 0
    bplan_rows_kernel<<<blocks, 256, 0, stream>>>(rowptr, num_rows, nbr);
    GD_LAUNCH_CHECK();
    GD_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, nbr, bptr, n1, stream));
    bplan_split_kernel<<<blocks, 256, 0, stream>>>(bptr, num_rows, bptr + num_rows, num_workers, split, npiece);
    GD_LAUNCH_CHECK();
    GD_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, split, hid, n1, stream));
    GD_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, npiece, pbeg, n1, stream));
    bplan_sizes_kernel<<<1, 32, 0, stream>>>(bptr, hid, pbeg, num_rows, sizes);
    GD_LAUNCH_CHECK();
    return GD_OK;
}

extern "C" int gd_spmm_bplan_fill(const int32_t* rowptr, const int32_t* col, int64_t num_rows, int32_t num_batches,
                                  int32_t batches_per_worker, int32_t* desc, int32_t* colp, int64_t* slot_of_entry,
                                  int32_t* piece_split, int32_t* split_row, int32_t* split_piece_beg, int32_t* split_npiece,
                                  const void* workspace, gd_stream_t stream_) {
    cudaStream_t stream = as_stream(stream_);
    GD_CHECK_ARG(rowptr && desc && colp && slot_of_entry && workspace, "null pointer");
    GD_CHECK_ARG(num_rows > 0 && num_batches > 0 && batches_per_worker > 0, "bad sizes");
    const size_t arr = align_up((size_t)(num_rows + 1) * sizeof(int32_t));
    const char* p = static_cast<const char*>(workspace);
    FillArgs a{rowptr, col, reinterpret_cast<const int32_t*>(p + arr), reinterpret_cast<const int32_t*>(p + 3 * arr),
               reinterpret_cast<const int32_t*>(p + 5 * arr), num_rows, batches_per_worker, desc, colp, slot_of_entry,
               piece_split, split_row, split_piece_beg, split_npiece};
    // two batches of slack after the plan: all-padding columns, descriptor 0
    GD_CUDA(cudaMemsetAsync(colp + (int64_t)num_batches * kSlots, 0xff, 2 * kSlots * sizeof(int32_t), stream));
    GD_CUDA(cudaMemsetAsync(desc + num_batches, 0, 2 * sizeof(int32_t), stream));
    const int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(num_rows, 8), kNumSMs * 32);
    bplan_fill_kernel<<<blocks, 256, 0, stream>>>(a);
    GD_LAUNCH_CHECK();
    return GD_OK;
}
