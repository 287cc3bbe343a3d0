// Dot-product / DistMult edge decoder fused with the Deleted-Edge-Consistency
// ("randomness") and edge-form Neighbourhood-Influence MSE losses and d loss/d logit.
// Gather-reduce, L2/HBM bound: each pair reads two (DEC items: four) 256 B rows of z.
// The gradient w.r.t. z is NOT scattered with atomics: d loss / d logit is written to the
// two incidence slots of each pair and dz is then one deterministic CSR gather (gd_spmm).
#include <climits>

#include "common.cuh"

namespace gd {

struct EdgeLossArgs {
    const float* z; int64_t ldz; int dim;
    const int32_t* pu; const int32_t* pv;
    int64_t n_df, n_ni;
    const float* target;
    const int32_t* pos_u; const int32_t* pos_v;
    float* logits; float* inc_val; float* partial;   // partial[gridDim.x][2]
    float c_r, c_l;                                    // alpha*2/norm_df, (1-alpha)*2/norm_ni
    int64_t own_df, own_ni;                            // leading items whose squared residual this caller counts
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

// Vector path (dim = 4 * LANES): a sub-warp of LANES lanes handles TWO decoder items per step, i.e.
// the Df pair i and its negative (4 rows) twice, or four NI pairs (8 rows): 8 independent 128-bit
// row gathers in flight per lane.  The pair endpoints / incidence slots of the NEXT step are
// fetched while the rows of the current step are in flight (software pipeline across steps), so a
// step costs one memory latency instead of two.
template <int LANES>
__global__ void __launch_bounds__(256, 4) edge_loss_fwd_kernel(const EdgeLossArgs a) {
    pdl_wait();
    pdl_trigger();
    constexpr int PER_WARP = 32 / LANES;
    const int lane = threadIdx.x & 31, warp_in_block = threadIdx.x >> 5;
    const int sub = lane / LANES, sl = lane % LANES;
    const unsigned mask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << (sub * LANES));
    // steps: [0, S_dec) handle Df items 2s, 2s+1 (pairs i and n_df + i); [S_dec, S_dec + S_ni) handle 4 NI pairs each.
    // 32-bit step / pair arithmetic (the host checks P < 2^31): the kernel is instruction bound (ncu: issue slots 68 % busy).
    const int n_df = (int)a.n_df, n_ni = (int)a.n_ni, own_df = (int)a.own_df, own_ni = (int)a.own_ni;
    const int S_dec = (n_df + 1) / 2, S_ni = (n_ni + 3) / 4, S = S_dec + S_ni;
    const int G = (int)gridDim.x * 8 * PER_WARP;
    int st = ((int)blockIdx.x * 8 + warp_in_block) * PER_WARP + sub;
    float sum_r = 0.f, sum_l = 0.f;

    // Pair slot q of a step is OWNED by one lane of the sub-warp (it loads the slot's indices and writes its
    // outputs): lane q, or lane 4q when LANES = 16 - where the four dot products are reduced by a transposing
    // butterfly (5 shuffles for all four instead of 4 x 4) that leaves the total of slot q on lanes 4q .. 4q+3.
    // Slots (0,1) = Df items (2s, 2s+1), slots (2,3) = their negatives; for NI steps four consecutive NI pairs.
    constexpr bool T16 = LANES == 16;
    const bool owner = T16 ? (sl & 3) == 0 : sl < 4;
    const int my_q = T16 ? sl >> 2 : sl;
    // pair index of this lane's slot at step s (-1: none)
    auto my_pair = [&](int s) -> int {
        if (!owner || s >= S) return -1;
        if (s < S_dec) { const int i = 2 * s + (my_q & 1); return i < n_df ? ((my_q & 2) ? n_df + i : i) : -1; }
        const int j = 4 * (s - S_dec) + my_q;
        return j < n_ni ? 2 * n_df + j : -1;
    };
    auto fetch = [&](int p, bool ni, int& u, int& v, int& pu_, int& pv_, float& tgt) {
        u = -1; v = 0; pu_ = 0; pv_ = 0; tgt = 0.f;
        if (p >= 0) {
            u = __ldg(a.pu + p); v = __ldg(a.pv + p);
            pu_ = __ldg(a.pos_u + p); pv_ = __ldg(a.pos_v + p);
            if (ni) tgt = __ldg(a.target + (p - 2 * n_df));
        }
    };
    unsigned long long zl = reinterpret_cast<unsigned long long>(a.z) + sl * 16;       // this lane's 16 bytes of every row
    asm volatile("" : "+l"(zl));
    const unsigned pitch = (unsigned)(a.ldz * 4);
    auto row = [&](int r) -> float4 {          // u = -1 marks an empty slot: zero row (the load is predicated off)
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r >= 0) x = __ldg(reinterpret_cast<const float4*>(zl + (unsigned long long)(unsigned)r * pitch));
        return x;
    };
    int p0 = my_pair(st);
    int u0, v0, pu0, pv0; float t0;
    fetch(p0, st >= S_dec, u0, v0, pu0, pv0, t0);
    while (st < S) {
        const int p1 = my_pair(st + G);
        int u1, v1, pu1, pv1; float t1;
        fetch(p1, st + G >= S_dec, u1, v1, pu1, pv1, t1);         // prefetch the next step's indices
        float4 ra[4], rb[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int src = T16 ? 4 * q : q;
            const int uu = __shfl_sync(mask, u0, src, LANES), vv = __shfl_sync(mask, v0, src, LANES);
            ra[q] = row(uu);
            rb[q] = row(uu >= 0 ? vv : -1);
        }
        float d[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) d[q] = ra[q].x * rb[q].x + ra[q].y * rb[q].y + ra[q].z * rb[q].z + ra[q].w * rb[q].w;
        float mine, other;
        if (T16) {
            const bool b3 = sl & 8, b2 = sl & 4;
            const float r0 = (b3 ? d[2] : d[0]) + __shfl_xor_sync(mask, b3 ? d[0] : d[2], 8, LANES);
            const float r1 = (b3 ? d[3] : d[1]) + __shfl_xor_sync(mask, b3 ? d[1] : d[3], 8, LANES);
            float t = (b2 ? r1 : r0) + __shfl_xor_sync(mask, b2 ? r0 : r1, 4, LANES);
            t += __shfl_xor_sync(mask, t, 2, LANES);
            t += __shfl_xor_sync(mask, t, 1, LANES);
            mine = t;                                              // total of slot sl >> 2
            other = __shfl_xor_sync(mask, t, 8, LANES);            // slot q ^ 2: the negative of a Df item / vice versa
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
                for (int o = LANES / 2; o > 0; o >>= 1) d[q] += __shfl_xor_sync(mask, d[q], o, LANES);
            }
            mine = sl == 0 ? d[0] : (sl == 1 ? d[1] : (sl == 2 ? d[2] : d[3]));
            other = sl == 0 ? d[2] : (sl == 1 ? d[3] : (sl == 2 ? d[0] : d[1]));
        }
        if (p0 >= 0) {
            float c;
            if (st < S_dec) {
                const float r = (my_q < 2) ? mine - other : other - mine;      // pos - neg
                c = (my_q < 2) ? a.c_r * r : -a.c_r * r;
                if (my_q < 2 && p0 < own_df) sum_r += r * r;
            } else {
                const float r = mine - t0;
                c = a.c_l * r;
                if (p0 - 2 * n_df < own_ni) sum_l += r * r;
            }
            a.logits[p0] = mine;
            a.inc_val[pu0] = c;
            a.inc_val[pv0] = c;
        }
        st += G; p0 = p1; u0 = u1; v0 = v1; pu0 = pu1; pv0 = pv1; t0 = t1;
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

// any width: one warp per item, scalar strips
__global__ void __launch_bounds__(256) edge_loss_fwd_generic_kernel(const EdgeLossArgs a) {
    const int lane = threadIdx.x & 31, warp_in_block = threadIdx.x >> 5;
    const int64_t items = a.n_df + a.n_ni;
    const int64_t stride = (int64_t)gridDim.x * 8;
    float sum_r = 0.f, sum_l = 0.f;
    auto dot = [&](int u, int v) -> float {
        float s = 0.f;
        for (int f = lane; f < a.dim; f += 32) s = fmaf(a.z[(int64_t)u * a.ldz + f], a.z[(int64_t)v * a.ldz + f], s);
        return warp_sum(s);
    };
    for (int64_t item = (int64_t)blockIdx.x * 8 + warp_in_block; item < items; item += stride) {
        if (item < a.n_df) {
            const int64_t p = item, q = a.n_df + item;
            const float lp = dot(a.pu[p], a.pv[p]);
            const float ln = dot(a.pu[q], a.pv[q]);
            if (lane == 0) {
                const float r = lp - ln;
                const float c = a.c_r * r;
                a.logits[p] = lp; a.logits[q] = ln;
                a.inc_val[a.pos_u[p]] = c;  a.inc_val[a.pos_v[p]] = c;
                a.inc_val[a.pos_u[q]] = -c; a.inc_val[a.pos_v[q]] = -c;
                if (item < a.own_df) sum_r += r * r;
            }
        } else {
            const int64_t j = item - a.n_df, p = 2 * a.n_df + j;
            const float l = dot(a.pu[p], a.pv[p]);
            if (lane == 0) {
                const float r = l - a.target[j];
                const float c = a.c_l * r;
                a.logits[p] = l;
                a.inc_val[a.pos_u[p]] = c; a.inc_val[a.pos_v[p]] = c;
                if (j < a.own_ni) sum_l += r * r;
            }
        }
    }
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
    pdl_wait();
    pdl_trigger();
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
    return (int)std::max<int64_t>(1, std::min<int64_t>(blocks, kNumSMs * 4));     // 4 resident CTAs per SM (64 registers): one wave
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
    return gd_edge_loss_fwd_part(z, ldz, dim, pair_u, pair_v, n_df, n_ni, target, alpha, pos_u, pos_v, logits, inc_val,
                                 losses, n_df, n_ni, n_df, n_ni, workspace, workspace_bytes, stream_);
}

extern "C" int gd_edge_loss_fwd_part(const float* z, int64_t ldz, int32_t dim, const int32_t* pair_u,
                                     const int32_t* pair_v, int64_t n_df, int64_t n_ni, const float* target,
                                     float alpha, const int32_t* pos_u, const int32_t* pos_v, float* logits,
                                     float* inc_val, float* losses, int64_t own_df, int64_t own_ni,
                                     int64_t norm_df, int64_t norm_ni, void* workspace, size_t workspace_bytes,
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
    const float inv_ndf = norm_df > 0 ? 1.0f / (float)norm_df : 0.f;
    const float inv_nni = norm_ni > 0 ? 1.0f / (float)norm_ni : 0.f;
    a.own_df = own_df; a.own_ni = own_ni;
    a.c_r = alpha * 2.0f * inv_ndf;
    a.c_l = (1.0f - alpha) * 2.0f * inv_nni;
    const bool vec = (ldz % 4 == 0) && ((uintptr_t)z % 16 == 0);
    int grid;
    GD_CHECK_ARG(2 * n_df + n_ni < (int64_t)INT32_MAX - 8, "more than 2^31 pairs");
    const int64_t steps = (n_df + 1) / 2 + (n_ni + 3) / 4;
    if (vec && dim == 64) { grid = edge_loss_grid(steps, 2); GD_CUDA(launch_pdl(edge_loss_fwd_kernel<16>, grid, 256, 0, stream, a)); }
    else if (vec && dim == 128) { grid = edge_loss_grid(steps, 1); GD_CUDA(launch_pdl(edge_loss_fwd_kernel<32>, grid, 256, 0, stream, a)); }
    else if (vec && dim == 32) { grid = edge_loss_grid(steps, 4); GD_CUDA(launch_pdl(edge_loss_fwd_kernel<8>, grid, 256, 0, stream, a)); }
    else { grid = edge_loss_grid(items, 1); edge_loss_fwd_generic_kernel<<<grid, 256, 0, stream>>>(a); }
    GD_LAUNCH_CHECK();
    GD_CUDA(launch_pdl(edge_loss_finalize_kernel, 1, 1024, 0, stream, (const float*)a.partial, (int)grid, inv_ndf, inv_nni, alpha, losses));
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
