// fp32 CUDA-core contractions over (gathered) rows: exact-fp32 path used for parity at
// 1e-5 and as the numerical reference of the tcgen05 path (gemm_tc.cu).
//   gd_gemm_rows     out[r(i),:] = epi(pro(a[r(i),:]) . B)        nn.Linear / deletion_weight
//   gd_gemm_tn_rows  c = sum_i a[r(i),:]^T (x) g[r(i),:]          weight gradients
#include "common.cuh"

namespace gd {

constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4;   // 256 threads: 16 x 16 thread grid

struct GemmArgs {
    const float* a; int64_t lda;
    const int32_t* rows; int64_t m; int k;
    const float* b; int b_is_nk; int n;
    const float* bias; const float* out_scale; const float* gate; int64_t ldgate;
    int relu_in, relu_out;
    float* out; int64_t ldo;
};

template <bool VEC>
__global__ void __launch_bounds__(256) gemm_rows_kernel(const GemmArgs g) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    __shared__ int32_t row_id[BM];
    const int t = threadIdx.x;
    const int tx = t & 15, ty = t >> 4;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    if (t < BM) {
        int64_t i = m0 + t;
        row_id[t] = i < g.m ? (g.rows ? g.rows[i] : (int32_t)i) : -1;
    }
    __syncthreads();
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int a_row = t >> 1, a_k = (t & 1) * 8;          // A tile: 128 rows x 16 k, 8 floats / thread
    const int32_t a_src = row_id[a_row];
    for (int k0 = 0; k0 < g.k; k0 += BK) {
        // ---- A tile (transposed into As[k][row])
        float av[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) av[i] = 0.f;
        if (a_src >= 0) {
            const float* ap = g.a + (int64_t)a_src * g.lda + k0 + a_k;
            if (VEC && k0 + a_k + 8 <= g.k) {
                float4 v0 = __ldg(reinterpret_cast<const float4*>(ap));
                float4 v1 = __ldg(reinterpret_cast<const float4*>(ap + 4));
                av[0] = v0.x; av[1] = v0.y; av[2] = v0.z; av[3] = v0.w;
                av[4] = v1.x; av[5] = v1.y; av[6] = v1.z; av[7] = v1.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) if (k0 + a_k + i < g.k) av[i] = __ldg(ap + i);
            }
            if (g.relu_in) {
#pragma unroll
                for (int i = 0; i < 8; ++i) av[i] = fmaxf(av[i], 0.f);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) As[a_k + i][a_row] = av[i];
        // ---- B tile into Bs[k][n]
        if (g.b_is_nk) {
            const int bn = t >> 2, bk = (t & 3) * 4;
            float bv[4] = {0.f, 0.f, 0.f, 0.f};
            if (n0 + bn < g.n) {
                const float* bp = g.b + (int64_t)(n0 + bn) * g.k + k0 + bk;
                if (VEC && k0 + bk + 4 <= g.k) {
                    float4 v = __ldg(reinterpret_cast<const float4*>(bp));
                    bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) if (k0 + bk + i < g.k) bv[i] = __ldg(bp + i);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) Bs[bk + i][bn] = bv[i];
        } else {
            const int bk = t >> 4, bn = (t & 15) * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float v = 0.f;
                if (k0 + bk < g.k && n0 + bn + i < g.n) v = __ldg(g.b + (int64_t)(k0 + bk) * g.n + n0 + bn + i);
                Bs[bk][bn + i] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * TM]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * TM + 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN]);
            const float ar[TM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float br[TN] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
        __syncthreads();
    }
    // ---- epilogue
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int32_t r = row_id[ty * TM + i];
        if (r < 0) continue;
        const float s = g.out_scale ? __ldg(g.out_scale + r) : 1.0f;
        float* op = g.out + (int64_t)r * g.ldo + n0 + tx * TN;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int c = n0 + tx * TN + j;
            if (c >= g.n) continue;
            float v = acc[i][j];
            if (g.bias) v += __ldg(g.bias + c);
            v *= s;
            if (g.relu_out) v = fmaxf(v, 0.f);
            if (g.gate && !(__ldg(g.gate + (int64_t)r * g.ldgate + c) > 0.f)) v = 0.f;
            op[j] = v;
        }
    }
}

// ---- weight gradient: C[k1,n2] tile 64x64 per CTA, rows split over blockIdx.x ------------
constexpr int WT = 64, WR = 16;   // output tile, rows per smem stage

__global__ void __launch_bounds__(256) gemm_tn_partial_kernel(
    const float* __restrict__ a, int64_t lda, const float* __restrict__ gmat, int64_t ldg,
    const int32_t* __restrict__ rows, int64_t m, int k1, int n2, int relu_a,
    const float* __restrict__ a_scale, int64_t rows_per_cta, float* __restrict__ partial) {
    __shared__ __align__(16) float As[WR][WT + 4];
    __shared__ __align__(16) float Gs[WR][WT + 4];
    const int t = threadIdx.x;
    const int tx = t & 15, ty = t >> 4;                   // 16x16 threads, 4x4 outputs each
    const int i0 = blockIdx.y * WT, j0 = blockIdx.z * WT; // tile origin in C
    const int64_t r_beg = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r_end = min(m, r_beg + rows_per_cta);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int lr = t >> 4, lc = (t & 15) * 4;             // loader: 16 rows x 64 cols, 4 floats / thread
    for (int64_t rb = r_beg; rb < r_end; rb += WR) {
        const int64_t i = rb + lr;
        float av[4] = {0.f, 0.f, 0.f, 0.f}, gv[4] = {0.f, 0.f, 0.f, 0.f};
        if (i < r_end) {
            const int64_t r = rows ? rows[i] : i;
            const float sc = a_scale ? __ldg(a_scale + r) : 1.0f;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (i0 + lc + u < k1) { float v = __ldg(a + r * lda + i0 + lc + u); av[u] = (relu_a ? fmaxf(v, 0.f) : v) * sc; }
                if (j0 + lc + u < n2) gv[u] = __ldg(gmat + r * ldg + j0 + lc + u);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { As[lr][lc + u] = av[u]; Gs[lr][lc + u] = gv[u]; }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < WR; ++kk) {
            float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            float4 g4 = *reinterpret_cast<const float4*>(&Gs[kk][tx * 4]);
            const float ar[4] = {a4.x, a4.y, a4.z, a4.w};
            const float gr[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
            for (int i2 = 0; i2 < 4; ++i2)
#pragma unroll
                for (int j2 = 0; j2 < 4; ++j2) acc[i2][j2] = fmaf(ar[i2], gr[j2], acc[i2][j2]);
        }
        __syncthreads();
    }
    float* p = partial + (int64_t)blockIdx.x * k1 * n2;
#pragma unroll
    for (int i2 = 0; i2 < 4; ++i2)
#pragma unroll
        for (int j2 = 0; j2 < 4; ++j2) {
            int ci = i0 + ty * 4 + i2, cj = j0 + tx * 4 + j2;
            if (ci < k1 && cj < n2) p[(int64_t)ci * n2 + cj] = acc[i2][j2];
        }
}

__global__ void reduce_partials_kernel(const float* __restrict__ partial, int nparts, int64_t count,
                                       float* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int p = 0; p < nparts; ++p) s += partial[(int64_t)p * count + i];
        out[i] = s;
    }
}

// dst[r, :] = scale[r] * src[r, :] for r in rows (scale optional).  VEC: 128-bit accesses, feat / 4 lanes per row
// (a warp moves 128 / feat ... rows per step); the scalar variant covers unaligned operands.
template <bool VEC>
__global__ void copy_rows_kernel(const float* __restrict__ src, int64_t lds, const int32_t* __restrict__ rows,
                                 int64_t m, int feat, const float* __restrict__ scale, float* __restrict__ dst, int64_t ldd) {
    pdl_wait();
    pdl_trigger();
    if (VEC) {
        const int f4 = feat >> 2;
        const int64_t total = m * f4, step = (int64_t)gridDim.x * blockDim.x;
        for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += step) {
            const int64_t i = t / f4;
            const int c = (int)(t - i * f4);
            const int64_t r = rows ? rows[i] : i;
            float4 v = __ldg(reinterpret_cast<const float4*>(src + r * lds) + c);
            if (scale) { const float sc = __ldg(scale + r); v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc; }
            reinterpret_cast<float4*>(dst + r * ldd)[c] = v;
        }
    } else {
        const int lane = threadIdx.x & 31;                  // one warp per row
        for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; i < m;
             i += ((int64_t)gridDim.x * blockDim.x) >> 5) {
            const int64_t r = rows ? rows[i] : i;
            const float sc = scale ? scale[r] : 1.0f;
            for (int f = lane; f < feat; f += 32) dst[r * ldd + f] = scale ? sc * src[r * lds + f] : src[r * lds + f];
        }
    }
}

__global__ void relu_bwd_kernel(const float* __restrict__ grad, const float* __restrict__ pre, int64_t count,
                                float* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x)
        out[i] = pre[i] > 0.f ? grad[i] : 0.f;
}

__global__ void relu_fwd_kernel(const float* __restrict__ x, int64_t count, float* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x)
        out[i] = fmaxf(x[i], 0.f);
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, float* __restrict__ step, int64_t count, float lr,
                            float beta1, float beta2, float eps) {
    pdl_wait();
    pdl_trigger();
    // torch.optim.Adam (single-tensor path): bias corrections from the incremented step
    const float t = *step + 1.0f;
    const float bc1 = 1.0f - powf(beta1, t);
    const float bc2 = 1.0f - powf(beta2, t);
    const float step_size = lr / bc1;
    const float bc2_sqrt = sqrtf(bc2);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float gi = g[i];
        const float mi = m[i] + (gi - m[i]) * (1.0f - beta1);          // exp_avg.lerp_(grad, 1-beta1)
        const float vi = v[i] * beta2 + (1.0f - beta2) * gi * gi;      // mul_(beta2).addcmul_
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] -= step_size * (mi / denom);
    }
}

__global__ void adam_bump_kernel(float* step) {
    pdl_wait();
    pdl_trigger();
    *step += 1.0f;
}

}  // namespace gd

using namespace gd;

extern "C" int gd_gemm_rows(const float* a, int64_t lda, const int32_t* rows, int64_t m, int32_t k,
                            const float* b, int32_t b_is_nk, int32_t n, const float* bias,
                            const float* out_scale, const float* gate, int64_t ldgate, int32_t relu_in,
                            int32_t relu_out, float* out, int64_t ldo, gd_stream_t stream) {
    GD_CHECK_ARG(m >= 0 && k > 0 && n > 0, "bad shape");
    if (m == 0) return GD_OK;
    GD_CHECK_ARG(a && b && out, "null pointer");
    GD_CHECK_ARG(lda >= k && ldo >= n, "leading dimension too small");
    GemmArgs g{a, lda, rows, m, k, b, b_is_nk, n, bias, out_scale, gate, ldgate, relu_in, relu_out, out, ldo};
    dim3 grid((unsigned)ceil_div<int64_t>(m, BM), (unsigned)ceil_div(n, BN));
    const bool vec = (lda % 4 == 0) && (k % 4 == 0) && (((uintptr_t)a | (uintptr_t)b) % 16 == 0);
    if (vec) gemm_rows_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(g);
    else gemm_rows_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(g);
    GD_LAUNCH_CHECK();
    return GD_OK;
}

static int tn_parts(int64_t m) { return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div<int64_t>(m, 256), 2 * kNumSMs)); }

extern "C" size_t gd_gemm_tn_workspace_bytes(int64_t m, int32_t k1, int32_t n2) {
    return (size_t)tn_parts(m) * k1 * n2 * sizeof(float);
}

extern "C" int gd_gemm_tn_rows(const float* a, int64_t lda, const float* g, int64_t ldg, const int32_t* rows,
                               int64_t m, int32_t k1, int32_t n2, int32_t relu_a, const float* a_scale, float* c,
                               void* workspace, size_t workspace_bytes, gd_stream_t stream_) {
    cudaStream_t stream = as_stream(stream_);
    GD_CHECK_ARG(m >= 0 && k1 > 0 && n2 > 0, "bad shape");
    GD_CHECK_ARG(c != nullptr, "null output");
    if (m == 0) { GD_CUDA(cudaMemsetAsync(c, 0, (size_t)k1 * n2 * sizeof(float), stream)); return GD_OK; }
    GD_CHECK_ARG(a && g, "null pointer");
    if (workspace_bytes < gd_gemm_tn_workspace_bytes(m, k1, n2) || !workspace)
        return fail(GD_ERR_WORKSPACE, "gd_gemm_tn_rows: workspace too small");
    const int parts = tn_parts(m);
    int64_t rows_per_cta = ceil_div<int64_t>(ceil_div<int64_t>(m, parts), WR) * WR;
    dim3 grid((unsigned)ceil_div<int64_t>(m, rows_per_cta), (unsigned)ceil_div(k1, WT), (unsigned)ceil_div(n2, WT));
    float* partial = static_cast<float*>(workspace);
    gemm_tn_partial_kernel<<<grid, 256, 0, stream>>>(a, lda, g, ldg, rows, m, k1, n2, relu_a, a_scale, rows_per_cta, partial);
    GD_LAUNCH_CHECK();
    const int64_t count = (int64_t)k1 * n2;
    reduce_partials_kernel<<<(unsigned)ceil_div<int64_t>(count, 256), 256, 0, stream>>>(partial, (int)grid.x, count, c);
    GD_LAUNCH_CHECK();
    return GD_OK;
}

extern "C" int gd_copy_rows_scaled(const float* src, int64_t lds, const int32_t* rows, int64_t m, int32_t feat,
                                   const float* row_scale, float* dst, int64_t ldd, gd_stream_t stream) {
    if (m == 0) return GD_OK;
    GD_CHECK_ARG(src && dst && feat > 0, "bad argument");
    const bool vec = feat % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0 && (((uintptr_t)src | (uintptr_t)dst) % 16) == 0;
    if (vec) {
        const int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(m * (feat / 4), 256), kNumSMs * 16);
        GD_CUDA(launch_pdl(copy_rows_kernel<true>, blocks, 256, 0, as_stream(stream), src, lds, rows, m, feat, row_scale, dst, ldd));
    } else {
        const int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(m, 8), kNumSMs * 32);
        GD_CUDA(launch_pdl(copy_rows_kernel<false>, blocks, 256, 0, as_stream(stream), src, lds, rows, m, feat, row_scale, dst, ldd));
    }
    GD_LAUNCH_CHECK();
    return GD_OK;
}

extern "C" int gd_copy_rows(const float* src, int64_t lds, const int32_t* rows, int64_t m, int32_t feat,
                            float* dst, int64_t ldd, gd_stream_t stream) {
    return gd_copy_rows_scaled(src, lds, rows, m, feat, nullptr, dst, ldd, stream);
}

extern "C" int gd_relu_bwd(const float* grad, const float* pre, int64_t count, float* out, gd_stream_t stream) {
    if (count == 0) return GD_OK;
    GD_CHECK_ARG(grad && pre && out, "null pointer");
    int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(count, 256), kNumSMs * 32);
    relu_bwd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(grad, pre, count, out);
    GD_LAUNCH_CHECK();
    return GD_OK;
}

extern "C" int gd_relu_fwd(const float* x, int64_t count, float* out, gd_stream_t stream) {
    if (count == 0) return GD_OK;
    GD_CHECK_ARG(x && out, "null pointer");
    int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(count, 256), kNumSMs * 32);
    relu_fwd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x, count, out);
    GD_LAUNCH_CHECK();
    return GD_OK;
}

extern "C" int gd_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* step,
                            int64_t count, float lr, float beta1, float beta2, float eps, gd_stream_t stream) {
    GD_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && step, "null pointer");
    if (count > 0) {
        int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(count, 256), kNumSMs * 8);
        GD_CUDA(launch_pdl(adam_kernel, blocks, 256, 0, as_stream(stream), param, grad, exp_avg, exp_avg_sq, step, count, lr, beta1, beta2, eps));
        GD_LAUNCH_CHECK();
    }
    GD_CUDA(launch_pdl(adam_bump_kernel, 1, 1, 0, as_stream(stream), step));
    GD_LAUNCH_CHECK();
    return GD_OK;
}
