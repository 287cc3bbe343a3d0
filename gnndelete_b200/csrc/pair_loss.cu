// Dot-product / DistMult edge decoder fused with the Deleted-Edge-Consistency
// ("randomness") and edge-form Neighbourhood-Influence MSE losses and d loss/d logit.
// Gather-reduce, L2/HBM bound: each pair reads two (DEC items: four) 256 B rows of z.
// The gradient w.r.t. z is NOT scattered with atomics: d loss / d logit is written to the
// two incidence slots of each pair and dz is then one deterministic CSR gather (gd_spmm).
#include "common.cuh"

namespace gd {

struct EdgeLossArgs {
    const float* z; int64_t ldz; int dim;
    const int32_t* pu; const int32_t* pv;
    int64_t n_df, n_ni;
    const float* target;
    const int32_t* pos_u; const int32_t* pos_v;
    float* logits; float* inc_val; float* partial;   // partial[gridDim.x][2]
    float c_r, c_l;                                    // alpha*2/n_df, (1-alpha)*2/n_ni
};

template <int LANES>
__device__ __forceinline__ float pair_dot(const float* z, int64_t ldz, int u, int v, int sl, unsigned mask) {
    float4 a = ldg4(z + (int64_t)u * ldz + sl * 4);
    float4 b = ldg4(z + (int64_t)v * ldz + sl * 4);
    float s = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) s += __shfl_xor_sync(mask, s, o, LANES);
    return s;
}

// LANES == 0: generic width, one warp per item with scalar strips
template <int LANES>
__global__ void __launch_bounds__(256) edge_loss_fwd_kernel(const EdgeLossArgs a) {
    constexpr int L = LANES == 0 ? 32 : LANES;
    constexpr int PER_WARP = 32 / L;
    const int lane = threadIdx.x & 31, warp_in_block = threadIdx.x >> 5;
    const int sub = lane / L, sl = lane % L;
    const unsigned mask = (L == 32) ? 0xffffffffu : (((1u << L) - 1u) << (sub * L));
    const int64_t items = a.n_df + a.n_ni;
    const int64_t stride = (int64_t)gridDim.x * 8 * PER_WARP;
    float sum_r = 0.f, sum_l = 0.f;

    auto dot = [&](int u, int v) -> float {
        if (LANES != 0) return pair_dot<L>(a.z, a.ldz, u, v, sl, mask);
        float s = 0.f;
        for (int f = lane; f < a.dim; f += 32) s = fmaf(a.z[(int64_t)u * a.ldz + f], a.z[(int64_t)v * a.ldz + f], s);
        return warp_sum(s);
    };

    for (int64_t item = ((int64_t)blockIdx.x * 8 + warp_in_block) * PER_WARP + sub; item < items; item += stride) {
        if (item < a.n_df) {
            const int64_t p = item, q = a.n_df + item;
            const float lp = dot(a.pu[p], a.pv[p]);
            const float ln = dot(a.pu[q], a.pv[q]);
            if (sl == 0) {
                const float r = lp - ln;
                const float c = a.c_r * r;
                a.logits[p] = lp; a.logits[q] = ln;
                a.inc_val[a.pos_u[p]] = c;  a.inc_val[a.pos_v[p]] = c;
                a.inc_val[a.pos_u[q]] = -c; a.inc_val[a.pos_v[q]] = -c;
                sum_r += r * r;
            }
        } else {
            const int64_t j = item - a.n_df, p = 2 * a.n_df + j;
            const float l = dot(a.pu[p], a.pv[p]);
            if (sl == 0) {
                const float r = l - a.target[j];
                const float c = a.c_l * r;
                a.logits[p] = l;
                a.inc_val[a.pos_u[p]] = c; a.inc_val[a.pos_v[p]] = c;
                sum_l += r * r;
            }
        }
    }
    // deterministic block reduction: lanes -> warp -> block (fixed order)
    sum_r = warp_sum(sum_r);
    sum_l = warp_sum(sum_l);
    __shared__ float red[8][2];
    if (lane == 0) { red[warp_in_block][0] = sum_r; red[warp_in_block][1] = sum_l; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float r = 0.f, l = 0.f;
        for (int w = 0; w < 8; ++w) { r += red[w][0]; l += red[w][1]; }
        a.partial[2 * blockIdx.x] = r;
        a.partial[2 * blockIdx.x + 1] = l;
    }
}

__global__ void __launch_bounds__(1024) edge_loss_finalize_kernel(const float* __restrict__ partial, int nparts,
                                                                 float inv_ndf, float inv_nni, float alpha,
                                                                 float* __restrict__ losses) {
    __shared__ float sr[32], sl_[32];
    float r = 0.f, l = 0.f;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) { r += partial[2 * i]; l += partial[2 * i + 1]; }
    r = warp_sum(r); l = warp_sum(l);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sr[w] = r; sl_[w] = l; }
    __syncthreads();
    if (w == 0) {
        r = lane < (blockDim.x >> 5) ? sr[lane] : 0.f;
        l = lane < (blockDim.x >> 5) ? sl_[lane] : 0.f;
        r = warp_sum(r); l = warp_sum(l);
        if (lane == 0) {
            const float loss_r = r * inv_ndf, loss_l = l * inv_nni;
            losses[0] = alpha * loss_r + (1.0f - alpha) * loss_l;
            losses[1] = loss_r;
            losses[2] = loss_l;
        }
    }
}

__global__ void __launch_bounds__(256) pair_decode_kernel(const float* __restrict__ z, int64_t ldz, int dim,
                                                          const int32_t* __restrict__ pu, const int32_t* __restrict__ pv,
                                                          int64_t P, const float* __restrict__ relw,
                                                          const int32_t* __restrict__ prel, float* __restrict__ logits) {
    const int lane = threadIdx.x & 31;
    for (int64_t p = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; p < P;
         p += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const float* zu = z + (int64_t)pu[p] * ldz;
        const float* zv = z + (int64_t)pv[p] * ldz;
        const float* w = relw ? relw + (int64_t)prel[p] * dim : nullptr;
        float s = 0.f;
        for (int f = lane; f < dim; f += 32) {
            float t = __ldg(zu + f) * __ldg(zv + f);
            s = w ? fmaf(t, __ldg(w + f), s) : s + t;
        }
        s = warp_sum(s);
        if (lane == 0) logits[p] = s;
    }
}

static int edge_loss_grid(int64_t items, int per_warp) {
    int64_t blocks = ceil_div<int64_t>(items, 8 * per_warp);
    return (int)std::max<int64_t>(1, std::min<int64_t>(blocks, kNumSMs * 8));
}

}  // namespace gd

using namespace gd;

extern "C" size_t gd_edge_loss_workspace_bytes(int64_t num_pairs) {
    (void)num_pairs;
    return (size_t)kNumSMs * 8 * 2 * sizeof(float);
}

extern "C" int gd_edge_loss_fwd(const float* z, int64_t ldz, int32_t dim, const int32_t* pair_u,
                                const int32_t* pair_v, int64_t n_df, int64_t n_ni, const float* target,
                                float alpha, const int32_t* pos_u, const int32_t* pos_v, float* logits,
                                float* inc_val, float* losses, void* workspace, size_t workspace_bytes,
                                gd_stream_t stream_) {
    cudaStream_t stream = as_stream(stream_);
    GD_CHECK_ARG(n_df >= 0 && n_ni >= 0 && dim > 0, "bad shape");
    GD_CHECK_ARG(losses != nullptr, "null losses");
    GD_CHECK_ARG(2 * n_df + n_ni < INT32_MAX, "too many pairs");
    if (workspace_bytes < gd_edge_loss_workspace_bytes(2 * n_df + n_ni) || !workspace)
        return fail(GD_ERR_WORKSPACE, "gd_edge_loss_fwd: workspace too small");
    const int64_t items = n_df + n_ni;
    if (items == 0) { GD_CUDA(cudaMemsetAsync(losses, 0, 3 * sizeof(float), stream)); return GD_OK; }
    GD_CHECK_ARG(z && pair_u && pair_v && pos_u && pos_v && logits && inc_val, "null pointer");
    GD_CHECK_ARG(n_ni == 0 || target, "null target");
    EdgeLossArgs a;
    a.z = z; a.ldz = ldz; a.dim = dim; a.pu = pair_u; a.pv = pair_v; a.n_df = n_df; a.n_ni = n_ni;
    a.target = target; a.pos_u = pos_u; a.pos_v = pos_v; a.logits = logits; a.inc_val = inc_val;
    a.partial = static_cast<float*>(workspace);
    // mean over an empty set is taken as 0 (the reference substitutes torch.tensor(0), gnndelete.py:244-246)
    const float inv_ndf = n_df > 0 ? 1.0f / (float)n_df : 0.f;
    const float inv_nni = n_ni > 0 ? 1.0f / (float)n_ni : 0.f;
    a.c_r = alpha * 2.0f * inv_ndf;
    a.c_l = (1.0f - alpha) * 2.0f * inv_nni;
    const bool vec = (ldz % 4 == 0) && ((uintptr_t)z % 16 == 0);
    int grid;
    if (vec && dim == 64) { grid = edge_loss_grid(items, 2); edge_loss_fwd_kernel<16><<<grid, 256, 0, stream>>>(a); }
    else if (vec && dim == 128) { grid = edge_loss_grid(items, 1); edge_loss_fwd_kernel<32><<<grid, 256, 0, stream>>>(a); }
    else if (vec && dim == 32) { grid = edge_loss_grid(items, 4); edge_loss_fwd_kernel<8><<<grid, 256, 0, stream>>>(a); }
    else { grid = edge_loss_grid(items, 1); edge_loss_fwd_kernel<0><<<grid, 256, 0, stream>>>(a); }
    GD_LAUNCH_CHECK();
    edge_loss_finalize_kernel<<<1, 1024, 0, stream>>>(a.partial, grid, inv_ndf, inv_nni, alpha, losses);
    GD_LAUNCH_CHECK();
    return GD_OK;
}

extern "C" int gd_pair_decode(const float* z, int64_t ldz, int32_t dim, const int32_t* pair_u,
                              const int32_t* pair_v, int64_t num_pairs, const float* rel_weight,
                              const int32_t* pair_rel, float* logits, gd_stream_t stream) {
    if (num_pairs == 0) return GD_OK;
    GD_CHECK_ARG(z && pair_u && pair_v && logits && dim > 0, "bad argument");
    GD_CHECK_ARG(!rel_weight || pair_rel, "rel_weight without pair_rel");
    int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(num_pairs, 8), kNumSMs * 16);
    pair_decode_kernel<<<blocks, 256, 0, as_stream(stream)>>>(z, ldz, dim, pair_u, pair_v, num_pairs, rel_weight, pair_rel, logits);
    GD_LAUNCH_CHECK();
    return GD_OK;
}
