// RGCNConv(in, out, R, num_blocks=B|None), aggr='mean' (rgcn.py:17-22; PyG semantics in
// SURVEY.md §9.4):   out_i = sum_r mean_{k in N_r(i)} x_k . W_r + x_i . root + bias.
//
// PyG runs a Python loop over the R relations, each iteration masking the whole edge list.
// Here the CSR is sorted by (dst, relation, src) once and ONE kernel walks it: a CTA owns
// a tile of 64 destination rows and loops over the relations; for every relation it
//   (1) gathers the per-row relation mean into a shared-memory tile M_r[64, in]
//       (per-entry weight 1 / |N_r(i)| precomputed by gd_rgcn_norm),
//   (2) multiplies by the (block-diagonal) W_r staged in shared memory and accumulates into
//       register accumulators that live across all relations,
// then adds x_tile . root + bias and writes the tile once.  Relations with no entry in the
// tile are skipped block-uniformly.  The backward w.r.t. x is the same kernel on the
// transposed CSR with W_r^T / root^T and the forward's weights permuted to transposed order.
#include "common.cuh"

namespace gd {

constexpr int RT_M = 64;          // destination rows per CTA
constexpr int RT_THREADS = 256;

// per-entry weight 1 / count(dst, rel); entries of a row are sorted by relation
__global__ void rgcn_norm_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ rel, int64_t n,
                                 float* __restrict__ w) {
    for (int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; row < n; row += (int64_t)gridDim.x * blockDim.x) {
        int k = rowptr[row];
        const int end = rowptr[row + 1];
        while (k < end) {
            const int r = rel[k];
            int e = k + 1;
            while (e < end && rel[e] == r) ++e;
            const float inv = 1.0f / (float)(e - k);
            for (int j = k; j < e; ++j) w[j] = inv;
            k = e;
        }
    }
}

__global__ void permute_kernel(const float* __restrict__ src, const int32_t* __restrict__ perm, int64_t n,
                               float* __restrict__ dst) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[perm[i]] = src[i];
}

__global__ void gather_rows_kernel(const float* __restrict__ src, int64_t lds, const int64_t* __restrict__ idx,
                                   int64_t m, int64_t src_rows, int feat, float* __restrict__ dst, int64_t ldd,
                                   int32_t* __restrict__ bad) {
    const int lane = threadIdx.x & 31;
    for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; i < m;
         i += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const int64_t r = idx[i];
        if (r < 0 || r >= src_rows) { if (lane == 0) atomicAdd(bad, 1); continue; }
        for (int f = lane; f < feat; f += 32) dst[i * ldd + f] = src[r * lds + f];
    }
}

struct RgcnArgs {
    const int32_t* rowptr; const int32_t* col; const int32_t* rel; const float* w;
    const float* x; int64_t ldx;
    const float* weight;      // [R, B, ib, ob]   (ib = in/B, ob = out/B of the FORWARD layer)
    const float* root;        // forward [in, out]
    const float* bias;        // [fout] or null
    float* out; int64_t ldo;
    int64_t n;
    int num_rel, blocks, ib, ob;
};

// FIN / FOUT are the widths of THIS pass (transposed pass: FIN = out, FOUT = in of the layer).
template <int FIN, int FOUT, bool TRANSPOSED>
__global__ void __launch_bounds__(RT_THREADS) rgcn_tile_kernel(const RgcnArgs a) {
    extern __shared__ __align__(16) float smem[];
    constexpr int MP = FIN + 4;                        // padded row of the mean tile
    float* Msm = smem;                                 // [RT_M][MP]
    float* Wsm = smem + RT_M * MP;                     // [FIN][wcols]  wcols = FOUT / blocks
    __shared__ int cur[RT_M];
    __shared__ int endp[RT_M];

    constexpr int CG = FOUT / 8;                       // column groups (8 columns per thread)
    constexpr int RG = RT_THREADS / CG;                // row groups
    constexpr int RPT = RT_M / RG;                     // rows per thread
    static_assert(RPT >= 1, "tile too small for this width");
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int cg = t % CG, rg = t / CG;
    const int c0 = cg * 8;                             // first output column of this thread
    const int64_t row0 = (int64_t)blockIdx.x * RT_M;
    // widths per block for THIS pass
    const int fin_b = FIN / a.blocks, fout_b = FOUT / a.blocks;
    const int blk = c0 / fout_b;                       // block-diagonal block of this thread's columns
    const int cb = c0 - blk * fout_b;                  // column inside the block

    if (t < RT_M) {
        const int64_t r = row0 + t;
        cur[t] = r < a.n ? a.rowptr[r] : 0;
        endp[t] = r < a.n ? a.rowptr[r + 1] : 0;
    }
    float acc[RPT][8];
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    __syncthreads();

    for (int r = 0; r < a.num_rel; ++r) {
        // ---- (1) relation-r weighted sums of the tile rows into Msm; 8 rows per warp
        int any = 0;
        for (int lr = warp; lr < RT_M; lr += RT_THREADS / 32) {
            int k = cur[lr];
            const int end = endp[lr];
            float4 s[(FIN + 127) / 128];
#pragma unroll
            for (int q = 0; q < (FIN + 127) / 128; ++q) s[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            while (k < end && __ldg(a.rel + k) == r) {
                const int c = __ldg(a.col + k);
                const float wk = __ldg(a.w + k);
#pragma unroll
                for (int q = 0; q < (FIN + 127) / 128; ++q) {
                    const int f = q * 128 + lane * 4;
                    if (f < FIN) fma4(s[q], wk, ldg4(a.x + (int64_t)c * a.ldx + f));
                }
                ++k; any = 1;
            }
            __syncwarp();
            if (lane == 0) cur[lr] = k;
#pragma unroll
            for (int q = 0; q < (FIN + 127) / 128; ++q) {
                const int f = q * 128 + lane * 4;
                if (f < FIN) *reinterpret_cast<float4*>(&Msm[lr * MP + f]) = s[q];
            }
        }
        const int go = __syncthreads_or(any);
        if (!go) continue;
        // ---- (2) stage W_r: Wsm[k][c] for k in [0, FIN), c in [0, fout_b): the block of row k
        {
            const float* wr = a.weight + (int64_t)r * a.blocks * a.ib * a.ob;
            for (int i = t; i < FIN * fout_b; i += RT_THREADS) {
                const int k = i / fout_b, c = i - k * fout_b;
                const int b = k / fin_b, kk = k - b * fin_b;
                // forward: W[b][kk][c]; transposed: W[b][c][kk]^T i.e. element (c, kk) of the forward block
                Wsm[i] = TRANSPOSED ? __ldg(wr + ((int64_t)b * a.ib + c) * a.ob + kk)
                                    : __ldg(wr + ((int64_t)b * a.ib + kk) * a.ob + c);
            }
        }
        __syncthreads();
        // ---- (3) acc += M_r[:, block] . W_r[block]
        for (int kk = 0; kk < fin_b; ++kk) {
            const int k = blk * fin_b + kk;
            const float4 w0 = *reinterpret_cast<const float4*>(&Wsm[k * fout_b + cb]);
            const float4 w1 = *reinterpret_cast<const float4*>(&Wsm[k * fout_b + cb + 4]);
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const float m = Msm[(rg * RPT + i) * MP + k];
                acc[i][0] = fmaf(m, w0.x, acc[i][0]); acc[i][1] = fmaf(m, w0.y, acc[i][1]);
                acc[i][2] = fmaf(m, w0.z, acc[i][2]); acc[i][3] = fmaf(m, w0.w, acc[i][3]);
                acc[i][4] = fmaf(m, w1.x, acc[i][4]); acc[i][5] = fmaf(m, w1.y, acc[i][5]);
                acc[i][6] = fmaf(m, w1.z, acc[i][6]); acc[i][7] = fmaf(m, w1.w, acc[i][7]);
            }
        }
        __syncthreads();
    }

    // ---- root term: acc += x_tile . root (dense FIN x FOUT), staged in FIN/4 row chunks of root
    for (int i = t; i < RT_M * (FIN / 4); i += RT_THREADS) {
        const int lr = i / (FIN / 4), f = (i - lr * (FIN / 4)) * 4;
        const int64_t r = row0 + lr;
        const float4 v = r < a.n ? ldg4(a.x + r * a.ldx + f) : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(&Msm[lr * MP + f]) = v;
    }
    constexpr int KC = 32;                             // root rows per stage: KC x FOUT floats <= FIN*fout_b?  (checked on host)
    for (int k0 = 0; k0 < FIN; k0 += KC) {
        __syncthreads();
        for (int i = t; i < KC * FOUT; i += RT_THREADS) {
            const int k = i / FOUT, c = i - k * FOUT;
            // forward: root[k0+k][c]; transposed: root[c][k0+k]
            Wsm[i] = TRANSPOSED ? __ldg(a.root + (int64_t)c * FIN + k0 + k) : __ldg(a.root + (int64_t)(k0 + k) * FOUT + c);
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < KC; ++k) {
            const float4 w0 = *reinterpret_cast<const float4*>(&Wsm[k * FOUT + c0]);
            const float4 w1 = *reinterpret_cast<const float4*>(&Wsm[k * FOUT + c0 + 4]);
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const float m = Msm[(rg * RPT + i) * MP + k0 + k];
                acc[i][0] = fmaf(m, w0.x, acc[i][0]); acc[i][1] = fmaf(m, w0.y, acc[i][1]);
                acc[i][2] = fmaf(m, w0.z, acc[i][2]); acc[i][3] = fmaf(m, w0.w, acc[i][3]);
                acc[i][4] = fmaf(m, w1.x, acc[i][4]); acc[i][5] = fmaf(m, w1.y, acc[i][5]);
                acc[i][6] = fmaf(m, w1.z, acc[i][6]); acc[i][7] = fmaf(m, w1.w, acc[i][7]);
            }
        }
    }
    // ---- epilogue
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int64_t r = row0 + rg * RPT + i;
        if (r >= a.n) continue;
        float4 o0 = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        float4 o1 = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
        if (a.bias) { add4(o0, ldg4(a.bias + c0)); add4(o1, ldg4(a.bias + c0 + 4)); }
        stg4(a.out + r * a.ldo + c0, o0);
        stg4(a.out + r * a.ldo + c0 + 4, o1);
    }
}

template <int FIN, int FOUT, bool TR>
static int launch_rgcn(const RgcnArgs& a, cudaStream_t stream) {
    const int fout_b = FOUT / a.blocks;
    const size_t wfloats = std::max<size_t>((size_t)FIN * fout_b, (size_t)32 * FOUT);
    const size_t smem = ((size_t)RT_M * (FIN + 4) + wfloats) * sizeof(float);
    auto kern = rgcn_tile_kernel<FIN, FOUT, TR>;
    GD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)ceil_div<int64_t>(a.n, RT_M);
    kern<<<grid, RT_THREADS, smem, stream>>>(a);
    GD_LAUNCH_CHECK();
    return GD_OK;
}

}  // namespace gd

using namespace gd;

extern "C" int gd_rgcn_norm(const int32_t* rowptr, const int32_t* rel, int64_t n, float* w, gd_stream_t stream) {
    if (n == 0) return GD_OK;
    GD_CHECK_ARG(rowptr && rel && w, "null pointer");
    const int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(n, 128), kNumSMs * 16);
    rgcn_norm_kernel<<<blocks, 128, 0, as_stream(stream)>>>(rowptr, rel, n, w);
    GD_LAUNCH_CHECK();
    return GD_OK;
}

extern "C" int gd_permute_f32(const float* src, const int32_t* perm, int64_t n, float* dst, gd_stream_t stream) {
    if (n == 0) return GD_OK;
    GD_CHECK_ARG(src && perm && dst, "null pointer");
    const int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(n, 256), kNumSMs * 16);
    permute_kernel<<<blocks, 256, 0, as_stream(stream)>>>(src, perm, n, dst);
    GD_LAUNCH_CHECK();
    return GD_OK;
}

extern "C" int gd_gather_rows(const float* src, int64_t lds, int64_t src_rows, const int64_t* idx, int64_t m,
                              int32_t feat, float* dst, int64_t ldd, int32_t* status, gd_stream_t stream) {
    GD_CHECK_ARG(status != nullptr, "null status");
    GD_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), as_stream(stream)));
    if (m == 0) return GD_OK;
    GD_CHECK_ARG(src && idx && dst && feat > 0, "bad argument");
    const int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(m, 8), kNumSMs * 32);
    gather_rows_kernel<<<blocks, 256, 0, as_stream(stream)>>>(src, lds, idx, m, src_rows, feat, dst, ldd, status);
    GD_LAUNCH_CHECK();
    return GD_OK;
}

extern "C" int gd_rgcn_conv(const gd_csr_t* csr, const int32_t* rel, const float* entry_weight, const float* x,
                            int64_t ldx, const float* weight, const float* root, const float* bias,
                            int32_t num_rel, int32_t num_blocks, int32_t in_dim, int32_t out_dim,
                            int32_t transposed, float* out, int64_t ldo, gd_stream_t stream_) {
    cudaStream_t stream = as_stream(stream_);
    GD_CHECK_ARG(csr != nullptr, "null csr");
    if (csr->num_rows == 0) return GD_OK;
    GD_CHECK_ARG(csr->rowptr && x && weight && root && out, "null pointer");
    GD_CHECK_ARG(csr->nnz == 0 || (csr->col && rel && entry_weight), "null edge arrays");
    GD_CHECK_ARG(num_rel >= 1 && num_blocks >= 1, "bad relation / block count");
    GD_CHECK_ARG(in_dim % num_blocks == 0 && out_dim % num_blocks == 0, "dims must be divisible by num_blocks");
    GD_CHECK_ARG((out_dim / num_blocks) % 8 == 0 && (in_dim / num_blocks) % 8 == 0, "block width must be a multiple of 8");
    GD_CHECK_ARG(ldx % 4 == 0 && ldo % 4 == 0, "leading dimensions must be multiples of 4");
    RgcnArgs a{csr->rowptr, csr->col, rel, entry_weight, x, ldx, weight, root, transposed ? nullptr : bias, out, ldo,
               csr->num_rows, num_rel, num_blocks, in_dim / num_blocks, out_dim / num_blocks};
    const int fin = transposed ? out_dim : in_dim, fout = transposed ? in_dim : out_dim;
#define RGCN_CASE(FI, FO)                                                          \
    if (fin == FI && fout == FO)                                                   \
        return transposed ? launch_rgcn<FI, FO, true>(a, stream) : launch_rgcn<FI, FO, false>(a, stream);
    RGCN_CASE(128, 128) RGCN_CASE(128, 64) RGCN_CASE(64, 128) RGCN_CASE(64, 64)
    RGCN_CASE(128, 32) RGCN_CASE(32, 128) RGCN_CASE(64, 32) RGCN_CASE(32, 64) RGCN_CASE(32, 32)
#undef RGCN_CASE
    return fail(GD_ERR_INVALID, "gd_rgcn_conv: in/out dims must be in {32, 64, 128}");
}
