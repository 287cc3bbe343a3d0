// RGCNConv with block-diagonal relation weights (rgcn.py:17-22, num_blocks = 4 when num_edge_type > 20), aggr = 'mean':
//     out_i = sum_r mean_{k in N_r(i)} x_k . W_r + x_i . root + bias          (PyG semantics: SURVEY.md §9.4)
// as ONE pass over the edges - no [R N, out] intermediate, no dense expansion of the blocks.
//
// On a BioKG-shaped graph a (destination, relation) group holds ~1.06 edges (108 entries over 102 relations per
// node), so "aggregate per relation, then transform" saves nothing: the work IS one [1 x in/4] . [in/4 x out/4] product
// per block per edge (2048 MACs per edge for 128 -> 64).  What can be reused is the relation's weight: the plan sorts
// the edges of a TILE of destination rows by (relation, destination), a warp keeps its slice of W_r (64 registers)
// across the tile's edges of relation r, and accumulates `mean weight x (x_src . W_r)` into a per-warp shared-memory
// tile of output rows.  Lane l owns CPL output columns [l CPL, (l+1) CPL) (all in one block b) and reads the IB inputs
// of that block from x_src: the 8 lanes of a block read the same addresses (hardware broadcast), there is no
// cross-lane reduction at all; products run as packed FFMA2 over input pairs.
// A tile's edge list is cut into chunks of <= chunk edges (work items), so a hub row is spread over many warps; every
// item writes its partial tile to scratch and gd_rgcn_edge_reduce adds the chunks of a tile in order onto
// x . root + bias (written before by the tcgen05 GEMM) - no float atomics, bitwise reproducible.
// The gradient w.r.t. x is the same kernel on the transposed edge list with W_r^T blocks.
#include "common.cuh"

namespace gd {

struct RgcnEdgeArgs {
    const int32_t* item_tile; const int32_t* item_beg; const int32_t* item_end;
    const int32_t* ent_src; const int32_t* ent_meta; const float* ent_w;
    const float* x; int64_t ldx;
    const float* weight;            // [R, 4, IB, OB]
    float* scratch;                 // [num_items, T, 4 * OB]
    int32_t num_items;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// IB / OB: rows / columns of one weight block (4 blocks).  CPL: output columns per lane and pass; a pass covers 32 CPL
// columns, PASSES = 4 OB / (32 CPL).  T: destination rows per tile (meta = relation << 5 | row in tile).
// The source rows of the next G edges are staged in shared memory with cp.async while the current G edges are
// multiplied (the first version loaded x_src into registers and used it at once: one edge in flight per warp, 16 warps
// per SM at 128 registers -> latency bound at ~4 G edges/s).
template <int IB, int OB, int CPL, int T>
__global__ void __launch_bounds__(256, 2) rgcn_edge_kernel(const RgcnEdgeArgs a) {
    constexpr int IN = 4 * IB, OUT = 4 * OB, PASSES = OUT / (32 * CPL), NX = IB / 4, G = 4;
    constexpr int WARP_FLOATS = T * OUT + 2 * G * IN;
    static_assert(IB * CPL == 64, "a lane keeps 64 weights");
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* tile = smem + warp * WARP_FLOATS;
    float* xs = tile + T * OUT;                                   // [2][G][IN]
    const int warps_total = (gridDim.x * blockDim.x) >> 5;
    for (int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < a.num_items; item += warps_total) {
        const int e0 = __ldg(a.item_beg + item), e1 = __ldg(a.item_end + item);
        for (int i = lane; i < T * OUT / 4; i += 32) reinterpret_cast<float4*>(tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
        const int ngroups = (e1 - e0 + G - 1) / G;
        auto stage_rows = [&](int g) {                            // source rows of edge group g -> xs[g & 1]
            if (g < ngroups && lane < IN / 4) {
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    const int e = e0 + g * G + j;
                    if (e < e1) cp_async16(xs + ((g & 1) * G + j) * IN + 4 * lane, a.x + (int64_t)__ldg(a.ent_src + e) * a.ldx + 4 * lane);
                }
            }
            cp_async_commit();
        };
#pragma unroll 1
        for (int pass = 0; pass < PASSES; ++pass) {
            const int c0 = pass * 32 * CPL + lane * CPL;          // first output column of this lane
            const int b = c0 / OB;                                // its block
            const float* wl = a.weight + (int64_t)b * IB * OB + (c0 - b * OB);
            unsigned long long wreg[IB / 2][CPL];                 // packed over input pairs: {W[2q][c], W[2q+1][c]}
            int cur_rel = -1;
            stage_rows(0);
#pragma unroll 1
            for (int g = 0; g < ngroups; ++g) {
                stage_rows(g + 1);
                int metas[G]; float ws[G];
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    const int e = min(e0 + g * G + j, e1 - 1);
                    metas[j] = __ldg(a.ent_meta + e); ws[j] = __ldg(a.ent_w + e);
                }
                cp_async_wait<1>();
                __syncwarp();
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    if (e0 + g * G + j >= e1) break;
                    const int rel = metas[j] >> 5, dl = metas[j] & 31;
                    if (rel != cur_rel) {                         // warp-uniform: new relation group, reload the weight slice
                        cur_rel = rel;
                        const float* wr = wl + (int64_t)rel * 4 * IB * OB;
#pragma unroll
                        for (int q = 0; q < IB / 2; ++q) {
                            float lo[CPL], hi[CPL];
                            if (CPL == 2) {
                                const float2 u = __ldg(reinterpret_cast<const float2*>(wr + (2 * q) * OB));
                                const float2 v = __ldg(reinterpret_cast<const float2*>(wr + (2 * q + 1) * OB));
                                lo[0] = u.x; lo[1] = u.y; hi[0] = v.x; hi[1] = v.y;
                            } else {
                                const float4 u = __ldg(reinterpret_cast<const float4*>(wr + (2 * q) * OB));
                                const float4 v = __ldg(reinterpret_cast<const float4*>(wr + (2 * q + 1) * OB));
                                lo[0] = u.x; lo[1] = u.y; lo[2 % CPL] = u.z; lo[3 % CPL] = u.w;
                                hi[0] = v.x; hi[1] = v.y; hi[2 % CPL] = v.z; hi[3 % CPL] = v.w;
                            }
#pragma unroll
                            for (int c = 0; c < CPL; ++c) asm("mov.b64 %0, {%1, %2};" : "=l"(wreg[q][c]) : "f"(lo[c]), "f"(hi[c]));
                        }
                    }
                    // the IB inputs of this lane's block (the 8 lanes of a block read the same words: broadcast)
                    const float4* xp = reinterpret_cast<const float4*>(xs + ((g & 1) * G + j) * IN + b * IB);
                    unsigned long long acc[CPL];
#pragma unroll
                    for (int c = 0; c < CPL; ++c) acc[c] = 0ull;
#pragma unroll
                    for (int q = 0; q < NX; ++q) {
                        const float4 xv = xp[q];
                        unsigned long long x01, x23;
                        asm("mov.b64 %0, {%1, %2};" : "=l"(x01) : "f"(xv.x), "f"(xv.y));
                        asm("mov.b64 %0, {%1, %2};" : "=l"(x23) : "f"(xv.z), "f"(xv.w));
#pragma unroll
                        for (int c = 0; c < CPL; ++c) {
                            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[c]) : "l"(x01), "l"(wreg[2 * q][c]));
                            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[c]) : "l"(x23), "l"(wreg[2 * q + 1][c]));
                        }
                    }
                    float* tp = tile + dl * OUT + c0;
#pragma unroll
                    for (int c = 0; c < CPL; ++c) {
                        float s0, s1;
                        asm("mov.b64 {%0, %1}, %2;" : "=f"(s0), "=f"(s1) : "l"(acc[c]));
                        tp[c] = fmaf(ws[j], s0 + s1, tp[c]);
                    }
                }
                __syncwarp();                                     // the staging buffer of this group is refilled two groups on
            }
            cp_async_wait<0>();
        }
        __syncwarp();
        float4* sp = reinterpret_cast<float4*>(a.scratch + (int64_t)item * (T * OUT));
        for (int i = lane; i < T * OUT / 4; i += 32) sp[i] = reinterpret_cast<const float4*>(tile)[i];
        __syncwarp();
    }
}

// out[row, :] += sum over the chunks of row's tile (in chunk order) of scratch[chunk, row in tile, :]
__global__ void __launch_bounds__(256) rgcn_edge_reduce_kernel(const int32_t* __restrict__ tile_item_ptr, int64_t num_rows,
                                                               int tile_rows, int out_dim, const float* __restrict__ scratch,
                                                               float* __restrict__ out, int64_t ldo) {
    const int per_row = out_dim >> 2;
    const int64_t total = num_rows * per_row;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / per_row;
        const int c4 = (int)(i - row * per_row);
        const int64_t t = row / tile_rows;
        const int rl = (int)(row - t * tile_rows);
        const int k0 = __ldg(tile_item_ptr + t), k1 = __ldg(tile_item_ptr + t + 1);
        float4* op = reinterpret_cast<float4*>(out + row * ldo) + c4;
        float4 o = *op;
        for (int k = k0; k < k1; ++k)
            add4(o, __ldg(reinterpret_cast<const float4*>(scratch + ((int64_t)k * tile_rows + rl) * out_dim) + c4));
        *op = o;
    }
}

template <int IB, int OB, int CPL, int T>
static int launch_rgcn_edge(const RgcnEdgeArgs& a, cudaStream_t stream) {
    constexpr int smem = 8 * (T * 4 * OB + 2 * 4 * 4 * IB) * (int)sizeof(float);
    static bool configured = false;
    if (!configured) {
        GD_CUDA(cudaFuncSetAttribute(rgcn_edge_kernel<IB, OB, CPL, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    const int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(a.num_items, 8), (int64_t)kNumSMs * 2);
    rgcn_edge_kernel<IB, OB, CPL, T><<<blocks, 256, smem, stream>>>(a);
    GD_LAUNCH_CHECK();
    return GD_OK;
}

}  // namespace gd

using namespace gd;

extern "C" int32_t gd_rgcn_edge_tile_rows(int32_t in_block, int32_t out_block) {
    if (in_block == 32 && out_block == 16) return 32;
    if ((in_block == 16 || in_block == 32) && out_block == 32) return 16;
    return 0;                                                   // shape not covered
}

extern "C" int gd_rgcn_edge_conv(const int32_t* item_tile, const int32_t* item_beg, const int32_t* item_end,
                                 int64_t num_items, const int32_t* ent_src, const int32_t* ent_meta, const float* ent_w,
                                 const float* x, int64_t ldx, const float* weight, int32_t in_block, int32_t out_block,
                                 float* scratch, gd_stream_t stream_) {
    cudaStream_t stream = as_stream(stream_);
    if (num_items == 0) return GD_OK;
    GD_CHECK_ARG(item_tile && item_beg && item_end && ent_src && ent_meta && ent_w && x && weight && scratch, "null pointer");
    GD_CHECK_ARG(num_items < INT32_MAX && ldx % 4 == 0 && ldx >= 4 * in_block && (((uintptr_t)x | (uintptr_t)weight | (uintptr_t)scratch) % 16) == 0,
                 "rows must be 16-byte aligned");
    RgcnEdgeArgs a{item_tile, item_beg, item_end, ent_src, ent_meta, ent_w, x, ldx, weight, scratch, (int32_t)num_items};
    if (in_block == 32 && out_block == 16) return launch_rgcn_edge<32, 16, 2, 32>(a, stream);
    if (in_block == 16 && out_block == 32) return launch_rgcn_edge<16, 32, 4, 16>(a, stream);
    if (in_block == 32 && out_block == 32) return launch_rgcn_edge<32, 32, 2, 16>(a, stream);
    return fail(GD_ERR_INVALID, "gd_rgcn_edge_conv: block shape must be 32x16, 16x32 or 32x32");
}

extern "C" int gd_rgcn_edge_reduce(const int32_t* tile_item_ptr, int64_t num_rows, int32_t tile_rows, int32_t out_dim,
                                   const float* scratch, float* out, int64_t ldo, gd_stream_t stream) {
    if (num_rows == 0) return GD_OK;
    GD_CHECK_ARG(tile_item_ptr && scratch && out && tile_rows > 0 && out_dim % 4 == 0 && ldo % 4 == 0 && ldo >= out_dim, "bad argument");
    const int64_t total = num_rows * (out_dim >> 2);
    const int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(total, 256), (int64_t)kNumSMs * 16);
    rgcn_edge_reduce_kernel<<<blocks, 256, 0, as_stream(stream)>>>(tile_item_ptr, num_rows, tile_rows, out_dim, scratch, out, ldo);
    GD_LAUNCH_CHECK();
    return GD_OK;
}
