// Dense-block Neighbourhood-Influence loss of train_fullbatch (gnndelete.py:163-193, 239-241):
//   loss_l = mean_{(i,j) in M} ( sigmoid(<z_i, z_j>) - sigmoid(logits_ori[i, j]) )^2
// M = strictly-lower-triangular node pairs inside the 2-hop node set S, minus the Df pairs.
// The reference materialises three N x N boolean masks and z z^T (1.6 GB fp32 at the Cora shape)
// every epoch; here only the S x S block is touched, tile by tile, and nothing is materialised:
// a CTA owns 64 rows of z_S, sweeps all 64-row column blocks, forms the 64 x 64 logit tile with
// one small GEMM, turns it into the residual / coefficient tile in registers, and immediately
// contracts the coefficients with z_J (second GEMM) into its private 64 x 64 gradient tile.
// Every unordered pair is visited from both sides, so each CTA writes only its own rows of dz —
// deterministic, no atomics.  fp32 CUDA-core math (exact parity); a tcgen05 variant of the two
// tile GEMMs is the natural next step, the contraction is genuinely dense.
#include "common.cuh"

namespace gd {

constexpr int NT = 64;                 // tile edge (rows of z_S per CTA / per column block)
constexpr int ND = 64;                 // embedding width handled by this kernel

struct DenseNiArgs {
    const float* zs; int64_t ldz; int64_t n_s;
    const float* tgt_sig; int64_t ldt;            // sigmoid(logits_ori[S][:, S])
    const uint32_t* excl;                         // n_s * n_s bits, 1 = pair excluded (Df pairs, both orders)
    float scale;                                  // weight / |M|
    float* dzs; int64_t lddz;
    float* partial;                               // [gridDim.x] sum of squared residuals over i > j
};

__global__ void __launch_bounds__(256) dense_ni_kernel(const DenseNiArgs a) {
    extern __shared__ __align__(16) float sm[];
    float (*zIt)[NT + 4] = reinterpret_cast<float (*)[NT + 4]>(sm);                          // [ND][NT+4]  z_I^T
    float (*zJt)[NT + 4] = reinterpret_cast<float (*)[NT + 4]>(sm + ND * (NT + 4));          // [ND][NT+4]  z_J^T
    float (*zJ)[ND + 4] = reinterpret_cast<float (*)[ND + 4]>(sm + 2 * ND * (NT + 4));       // [NT][ND+4]  z_J
    float (*C)[NT + 4] = reinterpret_cast<float (*)[NT + 4]>(sm + 2 * ND * (NT + 4) + NT * (ND + 4));  // [NT][NT+4]
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const int64_t I0 = (int64_t)blockIdx.x * NT;
    // z_I^T
    for (int i = t; i < NT * (ND / 4); i += 256) {
        const int r = i / (ND / 4), c4 = i % (ND / 4);
        const int64_t gi = I0 + r;
        const float4 v = gi < a.n_s ? ldg4(a.zs + gi * a.ldz + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        zIt[c4 * 4 + 0][r] = v.x; zIt[c4 * 4 + 1][r] = v.y; zIt[c4 * 4 + 2][r] = v.z; zIt[c4 * 4 + 3][r] = v.w;
    }
    float dz[4][4];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) dz[p][q] = 0.f;
    float loss = 0.f;

    for (int64_t J0 = 0; J0 < a.n_s; J0 += NT) {
        __syncthreads();                                   // previous tile fully consumed (also covers the z_I^T fill)
        for (int i = t; i < NT * (ND / 4); i += 256) {
            const int r = i / (ND / 4), c4 = i % (ND / 4);
            const int64_t gj = J0 + r;
            const float4 v = gj < a.n_s ? ldg4(a.zs + gj * a.ldz + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            zJt[c4 * 4 + 0][r] = v.x; zJt[c4 * 4 + 1][r] = v.y; zJt[c4 * 4 + 2][r] = v.z; zJt[c4 * 4 + 3][r] = v.w;
            *reinterpret_cast<float4*>(&zJ[r][c4 * 4]) = v;
        }
        __syncthreads();
        // ---- logits tile G = z_I z_J^T : thread (ty, tx) -> rows ty*4.., cols tx*4..
        float g[4][4];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int q = 0; q < 4; ++q) g[p][q] = 0.f;
#pragma unroll 8
        for (int k = 0; k < ND; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(&zIt[k][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&zJt[k][tx * 4]);
            const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) g[p][q] = fmaf(ar[p], br[q], g[p][q]);
        }
        // ---- residual / coefficient tile
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int64_t i = I0 + ty * 4 + p;
            float cv[4] = {0.f, 0.f, 0.f, 0.f};
            if (i < a.n_s) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int64_t j = J0 + tx * 4 + q;
                    if (j < a.n_s && j != i) {
                        const int64_t bit = i * a.n_s + j;
                        if (!((__ldg(a.excl + (bit >> 5)) >> (bit & 31)) & 1u)) {
                            const float s = 1.0f / (1.0f + expf(-g[p][q]));
                            const float r = s - __ldg(a.tgt_sig + i * a.ldt + j);
                            cv[q] = a.scale * 2.0f * r * s * (1.0f - s);
                            if (i > j) loss += r * r;
                        }
                    }
                }
            }
            *reinterpret_cast<float4*>(&C[ty * 4 + p][tx * 4]) = make_float4(cv[0], cv[1], cv[2], cv[3]);
        }
        __syncthreads();
        // ---- dz_I += C z_J : thread (ty, tx) -> rows ty*4.., embedding columns tx*4..
#pragma unroll 8
        for (int j = 0; j < NT; ++j) {
            const float4 zv = *reinterpret_cast<const float4*>(&zJ[j][tx * 4]);
            const float zr[4] = {zv.x, zv.y, zv.z, zv.w};
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const float c = C[ty * 4 + p][j];
#pragma unroll
                for (int q = 0; q < 4; ++q) dz[p][q] = fmaf(c, zr[q], dz[p][q]);
            }
        }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int64_t i = I0 + ty * 4 + p;
        if (i < a.n_s) stg4(a.dzs + i * a.lddz + tx * 4, make_float4(dz[p][0], dz[p][1], dz[p][2], dz[p][3]));
    }
    // deterministic block reduction of the loss
    loss = warp_sum(loss);
    __shared__ float red[8];
    if ((t & 31) == 0) red[t >> 5] = loss;
    __syncthreads();
    if (t == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[w];
        a.partial[blockIdx.x] = s;
    }
}

__global__ void dense_ni_finalize_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < n; ++i) s += partial[i];
        *out = s;
    }
}

__global__ void add_rows_kernel(const float* __restrict__ src, int64_t lds, const int32_t* __restrict__ rows, int64_t m,
                                int feat, float* __restrict__ dst, int64_t ldd) {
    const int lane = threadIdx.x & 31;
    for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; i < m; i += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const int64_t r = rows[i];
        for (int f = lane; f < feat; f += 32) dst[r * ldd + f] += src[i * lds + f];
    }
}

}  // namespace gd

using namespace gd;

extern "C" size_t gd_dense_ni_workspace_bytes(int64_t n_s) { return (size_t)(ceil_div<int64_t>(n_s, NT) + 1) * sizeof(float); }

extern "C" int gd_dense_ni_fwd_bwd(const float* zs, int64_t ldz, int32_t dim, int64_t n_s, const float* tgt_sig, int64_t ldt,
                                   const uint32_t* excl_bits, float coef_scale, float* dzs, int64_t lddz, float* loss_sum,
                                   void* workspace, size_t workspace_bytes, gd_stream_t stream_) {
    cudaStream_t stream = as_stream(stream_);
    GD_CHECK_ARG(loss_sum != nullptr, "null loss_sum");
    GD_CHECK_ARG(dim == ND, "the dense NI kernel handles 64-wide embeddings (out_dim of the reference's models)");
    if (n_s == 0) { GD_CUDA(cudaMemsetAsync(loss_sum, 0, sizeof(float), stream)); return GD_OK; }
    GD_CHECK_ARG(zs && tgt_sig && excl_bits && dzs, "null pointer");
    GD_CHECK_ARG(ldz % 4 == 0 && lddz % 4 == 0, "leading dimensions must be multiples of 4");
    if (!workspace || workspace_bytes < gd_dense_ni_workspace_bytes(n_s))
        return fail(GD_ERR_WORKSPACE, "gd_dense_ni_fwd_bwd: workspace too small");
    DenseNiArgs a{zs, ldz, n_s, tgt_sig, ldt, excl_bits, coef_scale, dzs, lddz, static_cast<float*>(workspace)};
    const int grid = (int)ceil_div<int64_t>(n_s, NT);
    const size_t smem = (size_t)(2 * ND * (NT + 4) + NT * (ND + 4) + NT * (NT + 4)) * sizeof(float);
    GD_CUDA(cudaFuncSetAttribute(dense_ni_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dense_ni_kernel<<<grid, 256, smem, stream>>>(a);
    GD_LAUNCH_CHECK();
    dense_ni_finalize_kernel<<<1, 32, 0, stream>>>(a.partial, grid, loss_sum);
    GD_LAUNCH_CHECK();
    return GD_OK;
}

extern "C" int gd_add_rows(const float* src, int64_t lds, const int32_t* rows, int64_t m, int32_t feat, float* dst,
                           int64_t ldd, gd_stream_t stream) {
    if (m == 0) return GD_OK;
    GD_CHECK_ARG(src && rows && dst && feat > 0, "bad argument");
    const int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(m, 8), kNumSMs * 32);
    add_rows_kernel<<<blocks, 256, 0, as_stream(stream)>>>(src, lds, rows, m, feat, dst, ldd);
    GD_LAUNCH_CHECK();
    return GD_OK;
}
