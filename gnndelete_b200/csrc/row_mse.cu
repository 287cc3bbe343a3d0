// Node-embedding MSE terms of the layer-wise Del losses, forward and gradient in one pass:
//   Deleted-Edge-Consistency   loss_r = MSE(cat(z[h], z[t]), cat(z_ori[h'], z_ori[t']))
//   Neighbourhood-Influence    loss_l = MSE(z[S \ Df], z_ori[S \ Df])
// (framework/trainer/gnndelete_nodeemb.py:196-212 full-batch, :770-781 KG step).  The reference materialises
// four gathered [M, F] matrices per layer and scatters their gradients back with index_put atomics.  Here both
// terms are ONE destination-major incidence (row r of z -> the z_ori rows it is compared with, term in the sign
// of the entry): a warp reads z[r] once, streams its partners, accumulates the two squared-error sums and writes
// dz[r] - no atomics, fixed summation order, bitwise reproducible.  HBM/L2 gather bound: 4F bytes per entry.
#include <algorithm>

#include "common.cuh"

namespace gd {

struct RowMseArgs {
    const float* z; int64_t ldz;
    const float* ref; int64_t ldref;
    int dim; int64_t n;
    const int32_t* rowptr; const int32_t* code;   // code >= 0: term 0, partner = code; code < 0: term 1, partner = -1 - code
    float g0, g1;                                  // d loss / d (squared error) * 2 of the two terms
    float* dz; int64_t lddz;
    float* partial;                                // [num_warps][2] squared-error sums
};

template <int VEC> struct Frag;
template <> struct Frag<4> {
    float4 v;
    __device__ __forceinline__ void load(const float* p) { v = ldg4(p); }
    __device__ __forceinline__ void zero() { v = make_float4(0.f, 0.f, 0.f, 0.f); }
    __device__ __forceinline__ void store(float* p) const { stg4(p, v); }
    // d = a - b; returns |d|^2 and accumulates g += c * d
    __device__ __forceinline__ float diff_acc(const Frag& a, const Frag& b, float c) {
        const float dx = a.v.x - b.v.x, dy = a.v.y - b.v.y, dz = a.v.z - b.v.z, dw = a.v.w - b.v.w;
        v.x = fmaf(c, dx, v.x); v.y = fmaf(c, dy, v.y); v.z = fmaf(c, dz, v.z); v.w = fmaf(c, dw, v.w);
        return dx * dx + dy * dy + dz * dz + dw * dw;
    }
};
template <> struct Frag<1> {
    float v;
    __device__ __forceinline__ void load(const float* p) { v = __ldg(p); }
    __device__ __forceinline__ void zero() { v = 0.f; }
    __device__ __forceinline__ void store(float* p) const { *p = v; }
    __device__ __forceinline__ float diff_acc(const Frag& a, const Frag& b, float c) {
        const float d = a.v - b.v;
        v = fmaf(c, d, v);
        return d * d;
    }
};

// One warp per destination row (grid-stride over rows, so every warp's summation order is fixed by the
// launch shape); lanes own VEC-wide column fragments.  The partner loop is unrolled by 4 so that four
// independent row gathers are in flight per lane.
template <int VEC>
__global__ void __launch_bounds__(256) row_mse_kernel(const RowMseArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
    const int chunks = a.dim / VEC;
    float l0 = 0.f, l1 = 0.f;
    for (int64_t r = warp; r < a.n; r += nwarps) {
        const int beg = __ldg(a.rowptr + r), end = __ldg(a.rowptr + r + 1);
        for (int c = lane; c < chunks; c += 32) {
            Frag<VEC> zr, g;
            g.zero();
            if (beg < end) zr.load(a.z + r * a.ldz + c * VEC);
            int e = beg;
            for (; e + 4 <= end; e += 4) {
                int cd[4];
                Frag<VEC> t[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) cd[k] = __ldg(a.code + e + k);
#pragma unroll
                for (int k = 0; k < 4; ++k) t[k].load(a.ref + (int64_t)(cd[k] < 0 ? -1 - cd[k] : cd[k]) * a.ldref + c * VEC);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool s1 = cd[k] < 0;
                    const float q = g.diff_acc(zr, t[k], s1 ? a.g1 : a.g0);
                    if (s1) l1 += q; else l0 += q;
                }
            }
            for (; e < end; ++e) {
                const int cd = __ldg(a.code + e);
                const bool s1 = cd < 0;
                Frag<VEC> t;
                t.load(a.ref + (int64_t)(s1 ? -1 - cd : cd) * a.ldref + c * VEC);
                const float q = g.diff_acc(zr, t, s1 ? a.g1 : a.g0);
                if (s1) l1 += q; else l0 += q;
            }
            if (a.dz) g.store(a.dz + r * a.lddz + c * VEC);
        }
    }
    l0 = warp_sum(l0);
    l1 = warp_sum(l1);
    if (lane == 0) { a.partial[2 * warp] = l0; a.partial[2 * warp + 1] = l1; }
}

// losses = [a0 * w0 * S0 + a1 * w1 * S1, w0 * S0, w1 * S1]; fixed-order tree over the per-warp sums
__global__ void __launch_bounds__(256) row_mse_finalize_kernel(const float* partial, int64_t nwarps, float w0, float w1,
                                                               float a0, float a1, float* losses) {
    __shared__ double s0[256], s1[256];
    double x0 = 0.0, x1 = 0.0;
    for (int64_t i = threadIdx.x; i < nwarps; i += 256) { x0 += partial[2 * i]; x1 += partial[2 * i + 1]; }
    s0[threadIdx.x] = x0; s1[threadIdx.x] = x1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { s0[threadIdx.x] += s0[threadIdx.x + o]; s1[threadIdx.x] += s1[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float t0 = w0 * (float)s0[0], t1 = w1 * (float)s1[0];
        losses[1] = t0; losses[2] = t1;
        losses[0] = a0 * t0 + a1 * t1;
    }
}

static int row_mse_blocks(int64_t n) {
    return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div<int64_t>(n, 8), kNumSMs * 8));
}

}  // namespace gd

using namespace gd;

extern "C" size_t gd_row_mse_workspace_bytes(int64_t n) {
    return align_up((size_t)row_mse_blocks(n) * 8 * 2 * sizeof(float));
}

extern "C" int gd_row_mse_fwd_bwd(const float* z, int64_t ldz, const float* z_ref, int64_t ldref, int32_t dim,
                                  int64_t n, const int32_t* rowptr, const int32_t* code, float w0, float w1,
                                  float a0, float a1, float* dz, int64_t lddz, float* losses, void* workspace,
                                  size_t workspace_bytes, gd_stream_t stream) {
    GD_CHECK_ARG(z && z_ref && rowptr && losses && dim > 0 && n >= 0, "bad argument");
    GD_CHECK_ARG(ldz >= dim && ldref >= dim && (!dz || lddz >= dim), "leading dimension smaller than dim");
    if (!workspace || workspace_bytes < gd_row_mse_workspace_bytes(n))
        return fail(GD_ERR_WORKSPACE, "gd_row_mse_fwd_bwd: workspace too small");
    RowMseArgs a;
    a.z = z; a.ldz = ldz; a.ref = z_ref; a.ldref = ldref; a.dim = dim; a.n = n;
    a.rowptr = rowptr; a.code = code;
    a.g0 = 2.f * a0 * w0; a.g1 = 2.f * a1 * w1;
    a.dz = dz; a.lddz = lddz; a.partial = static_cast<float*>(workspace);
    const int blocks = row_mse_blocks(n);
    const bool vec = dim % 4 == 0 && ldz % 4 == 0 && ldref % 4 == 0 && (!dz || lddz % 4 == 0) &&
                     reinterpret_cast<uintptr_t>(z) % 16 == 0 && reinterpret_cast<uintptr_t>(z_ref) % 16 == 0 &&
                     reinterpret_cast<uintptr_t>(dz) % 16 == 0;
    if (vec) row_mse_kernel<4><<<blocks, 256, 0, as_stream(stream)>>>(a);
    else row_mse_kernel<1><<<blocks, 256, 0, as_stream(stream)>>>(a);
    GD_LAUNCH_CHECK();
    row_mse_finalize_kernel<<<1, 256, 0, as_stream(stream)>>>(a.partial, (int64_t)blocks * 8, w0, w1, a0, a1, losses);
    GD_LAUNCH_CHECK();
    return GD_OK;
}
