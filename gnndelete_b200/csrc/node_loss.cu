// Fused decoder + edge-form Neighbourhood-Influence loss + d loss / d z in ONE pass over the node -> incident-pair
// lists (gcn.py:26-35 decode; gnndelete.py:379-386 NI; the index_add backward of `z[u] * z[v]`).
//
// For node w with incident pairs {(w, x)}:   dz_w = sum_x g_wx z_x,   g_wx = d loss / d logit_wx.
//   * NI pair (residual local to the pair):  logit = <z_w, z_x> is recomputed from w's side (z_w stays in registers for
//     the whole row, z_x is the one gathered row), g = c_l (logit - target).  Every NI pair is visited from both of its
//     endpoints, so the dot product is computed twice - and in exchange there is no logits array, no coefficient
//     array, no second gather pass: one 256-byte row gather per incidence entry instead of two (forward) + one (backward).
//   * DEC pair (Df pair i / its negative, residual couples two pairs): g comes from gd_edge_loss_fwd (n_ni = 0), which
//     wrote it to the entry's slot of `valp`.
// The incidence uses the batch plan of the aggregation kernel (spmm_batched.cu): rows cut into batches of 8 padded
// slots, equal contiguous batch ranges per resident sub-warp, split rows reduced through scratch in piece order - no
// float atomics, bitwise reproducible.  `bmeta[b]` = row of batch b (low 24 bits) | NI flags of its 8 slots (high 8).
// A sub-warp is LANES lanes; a lane holds NLD 16-byte fragments of a row: fragment k = bytes [k*LANES*16 + sl*16, +16)
// (every warp-wide load instruction covers whole 128-byte lines).  The 8 partial dot products of a batch are reduced
// by a transposing butterfly (8 shuffles for all eight totals with LANES = 8) and broadcast back (8 shuffles).
#include <cstdlib>
#include <cstring>

#include "spmm_batched.cuh"

namespace gd {

struct NLArgs {
    const int32_t* desc;
    const int4* colp;
    const float* valp;               // per slot: NI target logit | DEC coefficient
    const int32_t* bmeta;            // [num_batches + 2]
    const void* z;                   // partner rows (fp32, or bf16 in the BF16 instantiations)
    const void* zself;               // own rows, same type (row r of the plan -> zself row row_slot[r], or r)
    const int32_t* row_slot;
    float* out;                      // dz, one row per plan row
    float* scratch;
    const int32_t* piece_split;
    const int32_t* split_row;
    const int32_t* split_piece_beg;
    const int32_t* split_npiece;
    int32_t* split_ticket;
    const int32_t* tail_rowptr;      // optional plain CSR of given-coefficient entries (this step's negative pairs)
    const int32_t* tail_col;
    const float* tail_val;
    float* partial;                  // [gridDim.x] sum of squared NI residuals
    int64_t ldz, ldself, ldo;
    int32_t num_batches, per_worker, feat;
    float c_l;
};

template <int NLD>
struct Frag { f4p f[NLD]; };

__device__ __forceinline__ unsigned long long pack2(unsigned lo, unsigned hi) {
    unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r;
}
// 8 bf16 (element 2k in the low half of word k) -> two fragments of 4 fp32 (shift / mask, no CVT)
__device__ __forceinline__ void unpack_bf16x8(const f4p& raw, f4p& a, f4p& b) {
    unsigned w0, w1, w2, w3;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(w0), "=r"(w1) : "l"(raw.lo));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(w2), "=r"(w3) : "l"(raw.hi));
    a.lo = pack2(w0 << 16, w0 & 0xffff0000u); a.hi = pack2(w1 << 16, w1 & 0xffff0000u);
    b.lo = pack2(w2 << 16, w2 & 0xffff0000u); b.hi = pack2(w3 << 16, w3 & 0xffff0000u);
}
// fp32 rows: fragment k = bytes [k * LANES * 16 + sl * 16, +16) of the row.  bf16 rows (NLD must be 2): ONE 16-byte load
// of 8 bf16 at byte sl * 16, unpacked into two fragments (features 8 sl .. 8 sl + 7).
// Plain (unhinted, unpredicated) 16-byte gathers: ncu on the first version showed 36 CS2R (zeroing for the predicated
// loads) and 35 R2UR (the L2 policy operand of every hinted load) per 16 loads.  Padding slots therefore point at a
// valid row (the plan's colp is rewritten, losses.py) and carry coefficient 0; the partner matrix of the loss is one
// [N, 64] block that stays L2 resident without hints.
__device__ __forceinline__ f4p ldg_plain(const char* p) {
    f4p r;
    asm("ld.global.nc.v2.b64 {%0, %1}, [%2];" : "=l"(r.lo), "=l"(r.hi) : "l"(p));
    return r;
}
template <int LANES, int NLD, bool BF16>
__device__ __forceinline__ Frag<NLD> load_frag(const char* p) {
    Frag<NLD> r;
    if (BF16) {
        unpack_bf16x8(ldg_plain(p), r.f[0], r.f[NLD - 1]);
    } else {
#pragma unroll
        for (int k = 0; k < NLD; ++k) r.f[k] = ldg_plain(p + k * LANES * 16);
    }
    return r;
}
__device__ __forceinline__ unsigned long long mul_p(unsigned long long a, unsigned long long b) {
    unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ unsigned long long fmad_p(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
template <int NLD>
__device__ __forceinline__ float dot_frag(const Frag<NLD>& a, const Frag<NLD>& b) {
    unsigned long long p = mul_p(a.f[0].lo, b.f[0].lo);
    p = fmad_p(a.f[0].hi, b.f[0].hi, p);
#pragma unroll
    for (int k = 1; k < NLD; ++k) { p = fmad_p(a.f[k].lo, b.f[k].lo, p); p = fmad_p(a.f[k].hi, b.f[k].hi, p); }
    float x, y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p));
    return x + y;
}
template <int NLD>
__device__ __forceinline__ void axpy_frag(Frag<NLD>& acc, float w, const Frag<NLD>& v) {
    unsigned long long ww;
    asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w));
#pragma unroll
    for (int k = 0; k < NLD; ++k) { acc.f[k].lo = fmad_p(ww, v.f[k].lo, acc.f[k].lo); acc.f[k].hi = fmad_p(ww, v.f[k].hi, acc.f[k].hi); }
}

// Eight per-lane partial sums d[0..7] -> the total of slot (sl * 8 / LANES) over the LANES lanes of the sub-warp, on
// every lane of that slot's group of LANES / 8 lanes.
template <int LANES>
__device__ __forceinline__ float transpose_reduce8(float (&d)[8], int sl, unsigned mask) {
    {
        const bool up = sl & (LANES / 2);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float send = up ? d[i] : d[i + 4], keep = up ? d[i + 4] : d[i];
            d[i] = keep + __shfl_xor_sync(mask, send, LANES / 2, LANES);
        }
    }
    {
        const bool up = sl & (LANES / 4);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float send = up ? d[i] : d[i + 2], keep = up ? d[i + 2] : d[i];
            d[i] = keep + __shfl_xor_sync(mask, send, LANES / 4, LANES);
        }
    }
    {
        const bool up = sl & (LANES / 8);
        const float send = up ? d[0] : d[1], keep = up ? d[1] : d[0];
        d[0] = keep + __shfl_xor_sync(mask, send, LANES / 8, LANES);
    }
#pragma unroll
    for (int o = LANES / 16; o > 0; o >>= 1) d[0] += __shfl_xor_sync(mask, d[0], o, LANES);
    return d[0];
}

template <int LANES, int NLD, bool TAIL, bool BF16, int MINB>
__global__ void __launch_bounds__(256, MINB) node_loss_kernel(const NLArgs a) {
    static_assert(!BF16 || NLD == 2, "bf16 rows: one 16-byte load = two fragments");
    constexpr int ELT = BF16 ? 2 : 4;
    constexpr int PER_WARP = 32 / LANES;
    constexpr int GRP = LANES / 8;               // lanes holding the same slot total after the reduction
    const int lane = threadIdx.x & 31, warp_in_block = threadIdx.x >> 5;
    const int sub = lane / LANES, sl = lane % LANES;
    const unsigned mask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << (sub * LANES));
    const int64_t worker = ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5) * PER_WARP + sub;
    const int myslot = sl / GRP;
    float sum_l = 0.f;
    const int64_t b0 = worker * a.per_worker;
    const int nb = b0 < a.num_batches ? (int)min((int64_t)a.per_worker, (int64_t)a.num_batches - b0) : 0;
    if (nb > 0) {
        unsigned long long xl = reinterpret_cast<unsigned long long>(a.z) + sl * 16;
        unsigned long long sl_base = reinterpret_cast<unsigned long long>(a.zself) + sl * 16;
        asm volatile("" : "+l"(xl), "+l"(sl_base));
        const unsigned pitch = (unsigned)(a.ldz * ELT), spitch = (unsigned)(a.ldself * ELT), opitch = (unsigned)(a.ldo * 4);
        // byte offset of fragment k inside an fp32 OUTPUT / scratch row
        auto out_off = [&](int k) -> int { return BF16 ? sl * 32 + k * 16 : k * LANES * 16 + sl * 16; };
        auto row_ptr = [&](int c) -> const char* { return reinterpret_cast<const char*>(xl + (unsigned long long)(unsigned)c * pitch); };
        auto self_of = [&](int meta) -> Frag<NLD> {
            int r = meta & 0xffffff;
            if (a.row_slot) r = __ldg(a.row_slot + r);
            return load_frag<LANES, NLD, BF16>(reinterpret_cast<const char*>(sl_base + (unsigned long long)(unsigned)r * spitch));
        };
        // one running batch index; the plan streams are addressed from their (uniform) base pointers
        int64_t b = b0;
        const float* wbase = a.valp + myslot;
        int4 c0 = __ldg(a.colp + 2 * b), c1 = __ldg(a.colp + 2 * b + 1);
        int d_cur = __ldg(a.desc + b), d_nxt = __ldg(a.desc + b + 1);
        int m_cur = __ldg(a.bmeta + b), m_nxt = __ldg(a.bmeta + b + 1);
        Frag<NLD> zs;
        int m_prev = ~m_cur;                     // forces the first batch to load its row
        Frag<NLD> acc;
#pragma unroll
        for (int k = 0; k < NLD; ++k) acc.f[k] = f4p_zero();

        for (int it = 0; it < nb; ++it, ++b) {
            // ---- next batch's column ids / descriptor / meta FIRST: they are consumed (rotated into the loop-carried
            //      registers) at the end of this iteration and must not arrive after the gathers issued below
            const int4 c0n = __ldg(a.colp + 2 * b + 2), c1n = __ldg(a.colp + 2 * b + 3);
            const int d_n2 = __ldg(a.desc + b + 2), m_n2 = __ldg(a.bmeta + b + 2);
            const float my_val = __ldg(wbase + 8 * b);
            // this node's own row travels with the gathers of the first batch of its row (same latency window)
            if (((m_cur ^ m_prev) & 0xffffff) != 0) zs = self_of(m_cur);
            Frag<NLD> v[8];
            v[0] = load_frag<LANES, NLD, BF16>(row_ptr(c0.x)); v[1] = load_frag<LANES, NLD, BF16>(row_ptr(c0.y));
            v[2] = load_frag<LANES, NLD, BF16>(row_ptr(c0.z)); v[3] = load_frag<LANES, NLD, BF16>(row_ptr(c0.w));
            v[4] = load_frag<LANES, NLD, BF16>(row_ptr(c1.x)); v[5] = load_frag<LANES, NLD, BF16>(row_ptr(c1.y));
            v[6] = load_frag<LANES, NLD, BF16>(row_ptr(c1.z)); v[7] = load_frag<LANES, NLD, BF16>(row_ptr(c1.w));
            // ---- logits of the 8 slots from this node's side
            float d[8];
#pragma unroll
            for (int s = 0; s < 8; ++s) d[s] = dot_frag<NLD>(zs, v[s]);
            const float tot = transpose_reduce8<LANES>(d, sl, mask);
            const bool is_ni = (m_cur >> (24 + myslot)) & 1;
            const float r = tot - my_val;
            const float g = is_ni ? a.c_l * r : my_val;          // padding slots: flag 0, value 0 (a valid row is gathered)
            if (is_ni && (sl % GRP) == 0) sum_l = fmaf(r, r, sum_l);
#pragma unroll
            for (int s = 0; s < 8; ++s) axpy_frag<NLD>(acc, __shfl_sync(mask, g, s * GRP, LANES), v[s]);
            // ---- end of a row (or of this worker's piece of it)
            if (d_cur < 0) {
                int row = d_cur & kDescId;
                bool write = true;
                if (d_cur & kDescPiece) {
                    const int piece = row;
#pragma unroll
                    for (int k = 0; k < NLD; ++k)
                        stg4(a.scratch + (int64_t)piece * a.feat + (out_off(k) >> 2), to_f4(acc.f[k]));
                    const int h = __ldg(a.piece_split + piece);
                    const int np = __ldg(a.split_npiece + h);
                    __threadfence();
                    int ticket = 0;
                    if (sl == 0) ticket = atomicAdd(a.split_ticket + h, 1);
                    ticket = __shfl_sync(mask, ticket, 0, LANES);
                    write = ticket == np - 1;
                    if (write) {                     // last piece to arrive: add the partial sums in piece order
                        __threadfence();
                        if (sl == 0) a.split_ticket[h] = 0;
                        const int p0 = __ldg(a.split_piece_beg + h);
#pragma unroll
                        for (int k = 0; k < NLD; ++k) {
                            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                            for (int p = 0; p < np; ++p)
                                add4(o, __ldcg(reinterpret_cast<const float4*>(a.scratch + (int64_t)(p0 + p) * a.feat + (out_off(k) >> 2))));
                            asm("mov.b64 %0, {%1, %2};" : "=l"(acc.f[k].lo) : "f"(o.x), "f"(o.y));
                            asm("mov.b64 %0, {%1, %2};" : "=l"(acc.f[k].hi) : "f"(o.z), "f"(o.w));
                        }
                        row = __ldg(a.split_row + h);
                    }
                }
                if (write) {
                    if (TAIL) {                      // given-coefficient entries of the second CSR (this step's negative pairs)
                        const int k0 = __ldg(a.tail_rowptr + row), k1 = __ldg(a.tail_rowptr + row + 1);
                        for (int k = k0; k < k1; ++k) {
                            const Frag<NLD> t = load_frag<LANES, NLD, BF16>(row_ptr(__ldg(a.tail_col + k)));
                            axpy_frag<NLD>(acc, __ldg(a.tail_val + k), t);
                        }
                    }
                    const unsigned long long ob = reinterpret_cast<unsigned long long>(a.out) + (unsigned long long)(unsigned)row * opitch;
#pragma unroll
                    for (int k = 0; k < NLD; ++k) *reinterpret_cast<float4*>(ob + out_off(k)) = to_f4(acc.f[k]);
                }
#pragma unroll
                for (int k = 0; k < NLD; ++k) acc.f[k] = f4p_zero();
            }
            d_cur = d_nxt; d_nxt = d_n2; m_prev = m_cur; m_cur = m_nxt; m_nxt = m_n2; c0 = c0n; c1 = c1n;
        }
    }
    // deterministic block reduction of the squared NI residuals: lanes -> warp -> block (fixed order)
    sum_l = warp_sum(sum_l);
    __shared__ float red[8];
    if (lane == 0) red[warp_in_block] = sum_l;
    __syncthreads();
    if (threadIdx.x == 0) {
        float l = 0.f;
        for (int w = 0; w < 8; ++w) l += red[w];
        a.partial[blockIdx.x] = l;
    }
}

// losses = (alpha loss_r + (1 - alpha) loss_l, loss_r, loss_l): loss_r comes from the DEC pass (`dec_losses[1]`), loss_l
// from the per-block sums of squared NI residuals; every NI pair was visited from both endpoints (factor 1/2).
__global__ void __launch_bounds__(1024) node_loss_finalize_kernel(const float* __restrict__ partial, int nparts,
                                                                 const float* __restrict__ dec_losses, float half_inv_nni,
                                                                 float alpha, float* __restrict__ losses,
                                                                 float* __restrict__ sums) {
    __shared__ float sh[32];
    float l = 0.f;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) l += partial[i];
    l = warp_sum(l);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = l;
    __syncthreads();
    if (w == 0) {
        l = lane < (blockDim.x >> 5) ? sh[lane] : 0.f;
        l = warp_sum(l);
        if (lane == 0) {
            const float loss_r = dec_losses ? dec_losses[1] : 0.f;
            const float loss_l = l * half_inv_nni;
            if (sums) sums[0] = l;                  // raw sum (row-partitioned callers all-reduce it)
            losses[0] = alpha * loss_r + (1.0f - alpha) * loss_l;
            losses[1] = loss_r;
            losses[2] = loss_l;
        }
    }
}

// DEC residuals of a list of Df items for the row-partitioned epoch: item i = Df pair (pu[i], pv[i]) and its negative
// (pu[n + i], pv[n + i]);  r_i = <z_u, z_v> - <z_nu, z_nv>;  coef_pos[i] = c_r r_i, coef_neg[i] = -c_r r_i (the
// d loss / d logit of the two pairs, consumed by gd_node_loss_fwd_bwd after an all-gather);  partial sums of r_i^2.
// 8 lanes per item, 16-byte loads, fp32 or bf16 rows.
template <bool BF16>
__global__ void __launch_bounds__(256) dec_items_kernel(const void* __restrict__ z, int64_t ldz, int feat,
                                                        const int32_t* __restrict__ pu, const int32_t* __restrict__ pv,
                                                        int64_t n, float c_r, float* __restrict__ coef_pos,
                                                        float* __restrict__ coef_neg, float* __restrict__ logits,
                                                        float* __restrict__ partial) {
    constexpr int VEC = BF16 ? 8 : 4, ELT = BF16 ? 2 : 4;
    const int lane = threadIdx.x & 31, sl = lane & 7, warp_in_block = threadIdx.x >> 5;
    const unsigned mask = 0xffu << (lane & 24);
    const int chunks = feat / VEC;
    const char* zb = static_cast<const char*>(z);
    const int64_t pitch = ldz * ELT;
    auto dot = [&](int u, int v) -> float {
        float s = 0.f;
        for (int k = sl; k < chunks; k += 8) {
            const f4p ra = ldg_plain(zb + (int64_t)u * pitch + k * 16), rb = ldg_plain(zb + (int64_t)v * pitch + k * 16);
            if (BF16) {
                f4p a0, a1, b0, b1;
                unpack_bf16x8(ra, a0, a1); unpack_bf16x8(rb, b0, b1);
                const float4 x0 = to_f4(a0), x1 = to_f4(a1), y0 = to_f4(b0), y1 = to_f4(b1);
                s += x0.x * y0.x + x0.y * y0.y + x0.z * y0.z + x0.w * y0.w + x1.x * y1.x + x1.y * y1.y + x1.z * y1.z + x1.w * y1.w;
            } else {
                const float4 x = to_f4(ra), y = to_f4(rb);
                s += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(mask, s, o, 8);
        return s;
    };
    float sum = 0.f;
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) >> 3;
    for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3; i < n; i += groups) {
        const float lp = dot(__ldg(pu + i), __ldg(pv + i));
        const float ln = dot(__ldg(pu + n + i), __ldg(pv + n + i));
        if (sl == 0) {
            const float r = lp - ln;
            coef_pos[i] = c_r * r; coef_neg[i] = -c_r * r;
            if (logits) { logits[i] = lp; logits[n + i] = ln; }
            sum = fmaf(r, r, sum);
        }
    }
    sum = warp_sum(sum);
    __shared__ float red[8];
    if (lane == 0) red[warp_in_block] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        partial[blockIdx.x] = t;
    }
}

// dz[pu[i], :] += coef[i] z[pv[i], :]  and  dz[pv[i], :] += coef[i] z[pu[i], :]  for a list of given-coefficient pairs
// (this step's negative pairs: resampled every epoch, gnndelete.py:221-225, so they have no slot in the fixed
// incidence).  Vector float reductions (red.global.v4.f32.add): no sort and no per-step incidence rebuild; the
// summation order of the few contributions per row is not fixed (the reference's index_add scatter is the same).
__global__ void __launch_bounds__(256) pair_scatter_add_kernel(const float* __restrict__ z, int64_t ldz, int feat,
                                                               const int32_t* __restrict__ pu, const int32_t* __restrict__ pv,
                                                               const float* __restrict__ coef, int64_t n,
                                                               float* __restrict__ dz, int64_t ldo) {
    const int sl = threadIdx.x & 7;
    const int chunks = feat >> 2;
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) >> 3;
    for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3; i < n; i += groups) {
        const int u = __ldg(pu + i), v = __ldg(pv + i);
        const float c = __ldg(coef + i);
        for (int k = sl; k < chunks; k += 8) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(z + (int64_t)u * ldz) + k);
            const float4 b = __ldg(reinterpret_cast<const float4*>(z + (int64_t)v * ldz) + k);
            asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(dz + (int64_t)u * ldo + 4 * k), "f"(c * b.x), "f"(c * b.y), "f"(c * b.z), "f"(c * b.w) : "memory");
            asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(dz + (int64_t)v * ldo + 4 * k), "f"(c * a.x), "f"(c * a.y), "f"(c * a.z), "f"(c * a.w) : "memory");
        }
    }
}

__global__ void __launch_bounds__(1024) sum_partials_kernel(const float* __restrict__ partial, int nparts, float scale,
                                                           float* __restrict__ out) {
    __shared__ float sh[32];
    float l = 0.f;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) l += partial[i];
    l = warp_sum(l);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = l;
    __syncthreads();
    if (w == 0) {
        l = lane < (blockDim.x >> 5) ? sh[lane] : 0.f;
        l = warp_sum(l);
        if (lane == 0) out[0] = l * scale;
    }
}

// Sub-warp shape / residency of the fp32 F = 64 instantiation: 8 lanes x 2 fragments (fewer shuffles per entry, 16
// gathers in flight per lane, 128 registers -> 2 CTAs / SM) or 16 lanes x 1 fragment (64-80 registers -> 3-4 CTAs / SM).
// GD_NL_CFG = "8x2" (default) | "16x1@4" | "16x1@3" selects it (A/B measurements, profiles/).
static int nl_cfg64() {
    static int cfg = -1;
    if (cfg < 0) {
        const char* e = getenv("GD_NL_CFG");
        cfg = 0;
        if (e && !strcmp(e, "16x1@4")) cfg = 1;
        else if (e && !strcmp(e, "16x1@3")) cfg = 2;
    }
    return cfg;
}

template <int LANES, int NLD, bool BF16, int MINB>
static int node_loss_workers_t() {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, node_loss_kernel<LANES, NLD, true, BF16, MINB>, 256, 0) != cudaSuccess) {
        cudaGetLastError();
        per_sm = 0;
    }
    if (per_sm <= 0) per_sm = MINB;
    return kNumSMs * per_sm * 8 * (32 / LANES);
}

template <int LANES, int NLD, bool BF16, int MINB>
static int launch_node_loss(const NLArgs& a, int64_t workers, cudaStream_t stream, int* grid_out) {
    const int per_cta = 8 * (32 / LANES);
    const unsigned blocks = (unsigned)ceil_div<int64_t>(workers, per_cta);
    *grid_out = (int)blocks;
    if (a.tail_rowptr) node_loss_kernel<LANES, NLD, true, BF16, MINB><<<blocks, 256, 0, stream>>>(a);
    else node_loss_kernel<LANES, NLD, false, BF16, MINB><<<blocks, 256, 0, stream>>>(a);
    GD_LAUNCH_CHECK();
    return GD_OK;
}

}  // namespace gd

using namespace gd;

extern "C" int32_t gd_node_loss_workers(int32_t feat, int32_t bf16) {
    if (bf16) {
        switch (feat) {
            case 128: return node_loss_workers_t<16, 2, true, 2>();
            case 64: return node_loss_workers_t<8, 2, true, 2>();
            default: return 0;
        }
    }
    switch (feat) {
        case 128: return node_loss_workers_t<16, 2, false, 2>();
        case 64: return nl_cfg64() == 1 ? node_loss_workers_t<16, 1, false, 4>()
                      : nl_cfg64() == 2 ? node_loss_workers_t<16, 1, false, 3>() : node_loss_workers_t<8, 2, false, 2>();
        case 32: return node_loss_workers_t<8, 1, false, 3>();
        default: return 0;
    }
}

extern "C" size_t gd_node_loss_workspace_bytes(int32_t num_workers) {
    return (size_t)(num_workers / 8 + 2) * sizeof(float);
}

extern "C" int gd_node_loss_fwd_bwd(const gd_spmm_bplan_t* plan, const int32_t* bmeta, const float* valp,
                                    const int32_t* tail_rowptr, const int32_t* tail_col, const float* tail_val,
                                    const void* z, int64_t ldz, const void* zself, int64_t ldself, int32_t bf16,
                                    const int32_t* row_slot, int32_t feat, int64_t norm_ni, float alpha,
                                    const float* dec_losses, float* dz, int64_t ldo, float* scratch, float* losses,
                                    float* ni_sq_sum, void* workspace, size_t workspace_bytes, gd_stream_t stream_) {
    cudaStream_t stream = as_stream(stream_);
    GD_CHECK_ARG(plan != nullptr && losses != nullptr, "null plan / losses");
    GD_CHECK_ARG(bf16 ? (feat == 64 || feat == 128) : (feat == 32 || feat == 64 || feat == 128),
                 "feat must be 32, 64 or 128 (bf16 rows: 64 or 128)");
    GD_CHECK_ARG(plan->num_rows < (1 << 24), "the batch meta word holds 24-bit row ids");
    GD_CHECK_ARG(plan->num_rows > 0 && plan->num_batches > 0, "empty plan");
    GD_CHECK_ARG(plan->desc && plan->colp && bmeta && valp && z && zself && dz, "null pointer");
    GD_CHECK_ARG(plan->num_workers > 0 && plan->batches_per_worker > 0 &&
                     (int64_t)plan->num_workers * plan->batches_per_worker >= plan->num_batches, "inconsistent worker partition");
    GD_CHECK_ARG(plan->num_piece == 0 || (scratch && plan->piece_split && plan->split_row && plan->split_piece_beg &&
                                          plan->split_npiece && plan->split_ticket), "split rows without scratch / ticket arrays");
    const int elt = bf16 ? 2 : 4, ldm = bf16 ? 8 : 4;
    GD_CHECK_ARG(ldz >= feat && ldself >= feat && ldo >= feat && ldz % ldm == 0 && ldself % ldm == 0 && ldo % 4 == 0,
                 "leading dimensions must be multiples of 16 bytes and >= feat");
    GD_CHECK_ARG(ldz * elt < (int64_t)1 << 32 && ldself * elt < (int64_t)1 << 32 && ldo * 4 < (int64_t)1 << 32, "row pitch out of range");
    GD_CHECK_ARG((((uintptr_t)z | (uintptr_t)zself | (uintptr_t)dz | (uintptr_t)scratch | (uintptr_t)plan->colp) % 16) == 0,
                 "operands must be 16-byte aligned");
    GD_CHECK_ARG(!tail_rowptr || (tail_col && tail_val), "tail CSR without columns / values");
    GD_CHECK_ARG(workspace && workspace_bytes >= gd_node_loss_workspace_bytes(plan->num_workers), "workspace too small");
    NLArgs a;
    a.desc = plan->desc; a.colp = reinterpret_cast<const int4*>(plan->colp); a.valp = valp; a.bmeta = bmeta;
    a.z = z; a.zself = zself; a.row_slot = row_slot; a.out = dz; a.scratch = scratch;
    a.piece_split = plan->piece_split; a.split_row = plan->split_row; a.split_piece_beg = plan->split_piece_beg;
    a.split_npiece = plan->split_npiece; a.split_ticket = plan->split_ticket;
    a.tail_rowptr = tail_rowptr; a.tail_col = tail_col; a.tail_val = tail_val;
    a.partial = static_cast<float*>(workspace);
    a.ldz = ldz; a.ldself = ldself; a.ldo = ldo;
    a.num_batches = (int32_t)plan->num_batches; a.per_worker = plan->batches_per_worker; a.feat = feat;
    const float inv_nni = norm_ni > 0 ? 1.0f / (float)norm_ni : 0.f;
    a.c_l = (1.0f - alpha) * 2.0f * inv_nni;
    int grid = 0, rc;
    if (bf16) {
        if (feat == 128) rc = launch_node_loss<16, 2, true, 2>(a, plan->num_workers, stream, &grid);
        else rc = launch_node_loss<8, 2, true, 2>(a, plan->num_workers, stream, &grid);
    } else if (feat == 128) rc = launch_node_loss<16, 2, false, 2>(a, plan->num_workers, stream, &grid);
    else if (feat == 64) {
        if (nl_cfg64() == 1) rc = launch_node_loss<16, 1, false, 4>(a, plan->num_workers, stream, &grid);
        else if (nl_cfg64() == 2) rc = launch_node_loss<16, 1, false, 3>(a, plan->num_workers, stream, &grid);
        else rc = launch_node_loss<8, 2, false, 2>(a, plan->num_workers, stream, &grid);
    } else rc = launch_node_loss<8, 1, false, 3>(a, plan->num_workers, stream, &grid);
    if (rc != GD_OK) return rc;
    GD_CHECK_ARG((size_t)grid * sizeof(float) <= workspace_bytes, "workspace too small for the launched grid");
    node_loss_finalize_kernel<<<1, 1024, 0, stream>>>(a.partial, grid, dec_losses, 0.5f * inv_nni, alpha, losses, ni_sq_sum);
    GD_LAUNCH_CHECK();
    return GD_OK;
}

extern "C" size_t gd_dec_items_workspace_bytes(void) { return (size_t)kNumSMs * 8 * sizeof(float); }

extern "C" int gd_dec_items_fwd(const void* z, int64_t ldz, int32_t bf16, int32_t feat, const int32_t* pair_u,
                                const int32_t* pair_v, int64_t n_items, int64_t norm_df, float alpha, float* coef_pos,
                                float* coef_neg, float* logits, float* loss_r_part, void* workspace,
                                size_t workspace_bytes, gd_stream_t stream_) {
    cudaStream_t stream = as_stream(stream_);
    GD_CHECK_ARG(loss_r_part != nullptr, "null output");
    if (n_items == 0) { GD_CUDA(cudaMemsetAsync(loss_r_part, 0, sizeof(float), stream)); return GD_OK; }
    const int vec = bf16 ? 8 : 4;
    GD_CHECK_ARG(z && pair_u && pair_v && coef_pos && coef_neg && n_items > 0, "null pointer");
    GD_CHECK_ARG(feat > 0 && feat % vec == 0 && ldz >= feat && ldz % vec == 0 && (uintptr_t)z % 16 == 0, "rows must be 16-byte aligned multiples of 16 bytes");
    GD_CHECK_ARG(workspace && workspace_bytes >= gd_dec_items_workspace_bytes(), "workspace too small");
    const float inv = norm_df > 0 ? 1.0f / (float)norm_df : 0.f;
    const int grid = (int)std::min<int64_t>(ceil_div<int64_t>(n_items, 32), kNumSMs * 8);
    float* partial = static_cast<float*>(workspace);
    if (bf16) dec_items_kernel<true><<<grid, 256, 0, stream>>>(z, ldz, feat, pair_u, pair_v, n_items, alpha * 2.0f * inv, coef_pos, coef_neg, logits, partial);
    else dec_items_kernel<false><<<grid, 256, 0, stream>>>(z, ldz, feat, pair_u, pair_v, n_items, alpha * 2.0f * inv, coef_pos, coef_neg, logits, partial);
    GD_LAUNCH_CHECK();
    sum_partials_kernel<<<1, 1024, 0, stream>>>(partial, grid, inv, loss_r_part);     // this caller's share of loss_r
    GD_LAUNCH_CHECK();
    return GD_OK;
}

extern "C" int gd_pair_scatter_add(const float* z, int64_t ldz, int32_t feat, const int32_t* pair_u, const int32_t* pair_v,
                                   const float* coef, int64_t num_pairs, float* dz, int64_t ldo, gd_stream_t stream) {
    if (num_pairs == 0) return GD_OK;
    GD_CHECK_ARG(z && pair_u && pair_v && coef && dz && feat > 0 && feat % 4 == 0, "bad argument");
    GD_CHECK_ARG(ldz % 4 == 0 && ldo % 4 == 0 && ldz >= feat && ldo >= feat && (((uintptr_t)z | (uintptr_t)dz) % 16) == 0,
                 "rows must be 16-byte aligned");
    const int grid = (int)std::min<int64_t>(ceil_div<int64_t>(num_pairs, 32), kNumSMs * 16);
    pair_scatter_add_kernel<<<grid, 256, 0, as_stream(stream)>>>(z, ldz, feat, pair_u, pair_v, coef, num_pairs, dz, ldo);
    GD_LAUNCH_CHECK();
    return GD_OK;
}
