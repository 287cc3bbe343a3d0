// Batched CSR aggregation with a bf16 SOURCE matrix (fp32 accumulation, fp32 output): the "bf16-gather" mode.
//
// Same contract and batch plan as gd_spmm_batched (spmm_batched.cu); only the gathered operand is stored in bf16, which
// halves the bytes every non-zero pulls through L2 / HBM.  It is the wire format of the row-partitioned epoch (the
// halo blocks travel over NVLink in bf16 and are aggregated as they arrive, dist.py) and an opt-in single-GPU mode;
// results differ from the fp32 path by the bf16 rounding of the source rows (stated tolerance 2e-2, north_star).
// A sub-warp is feat / 8 lanes; a lane loads 16 bytes = 8 bf16 of every gathered row and keeps 8 fp32 partial sums
// (four packed f32x2 registers).  bf16 -> fp32 is a shift / mask on the integer pipe (no CVT).
#include "spmm_batched.cuh"

namespace gd {

struct BArgs16 {
    const int32_t* desc;
    const int4* colp;
    const float4* valp;
    const float* row_scale;
    const void* x;                   // bf16 [*, ldx]
    const float* bias;
    float* out;
    float* scratch;
    const int32_t* piece_split;
    const int32_t* split_row;
    const int32_t* split_piece_beg;
    const int32_t* split_npiece;
    int32_t* split_ticket;
    const int32_t* tail_rowptr;
    const int32_t* tail_col;
    const float* tail_val;
    int64_t ldx, ldo;
    int32_t num_batches, per_worker, feat, accumulate;
    float self_coef;
};

struct f8p { f4p a, b; };           // 8 consecutive features: a = elements 0-3, b = elements 4-7

__device__ __forceinline__ unsigned long long pack2(unsigned lo, unsigned hi) {
    unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r;
}
// 8 bf16 (little endian: element 2k in the low half of word k) -> 8 fp32 as packed pairs
__device__ __forceinline__ f8p unpack_bf16x8(const f4p& raw) {
    unsigned w0, w1, w2, w3;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(w0), "=r"(w1) : "l"(raw.lo));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(w2), "=r"(w3) : "l"(raw.hi));
    f8p r;
    r.a.lo = pack2(w0 << 16, w0 & 0xffff0000u); r.a.hi = pack2(w1 << 16, w1 & 0xffff0000u);
    r.b.lo = pack2(w2 << 16, w2 & 0xffff0000u); r.b.hi = pack2(w3 << 16, w3 & 0xffff0000u);
    return r;
}
__device__ __forceinline__ void add8(f8p& acc, const f4p& raw) {
    const f8p v = unpack_bf16x8(raw);
    add_p(acc.a, v.a); add_p(acc.b, v.b);
}
__device__ __forceinline__ void fma8(f8p& acc, float w, const f4p& raw) {
    const f8p v = unpack_bf16x8(raw);
    fma_p(acc.a, w, v.a); fma_p(acc.b, w, v.b);
}

template <int LANES, bool WEIGHTED>
__global__ void __launch_bounds__(256, 3) spmm_batched_bf16_kernel(const BArgs16 a) {
    constexpr int PER_WARP = 32 / LANES;
    const int lane = threadIdx.x & 31;
    const int sub = lane / LANES, sl = lane % LANES;
    const unsigned mask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << (sub * LANES));
    const int64_t worker = ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5) * PER_WARP + sub;
    const int64_t b0 = worker * a.per_worker;
    if (b0 >= a.num_batches) return;
    const int nb = (int)min((int64_t)a.per_worker, (int64_t)a.num_batches - b0);
    unsigned long long xl = reinterpret_cast<unsigned long long>(a.x) + sl * 16;   // this lane's 8 bf16 of every source row
    unsigned long long ol = reinterpret_cast<unsigned long long>(a.out) + sl * 32; // ... and its 8 floats of every output row
    asm volatile("" : "+l"(xl), "+l"(ol));
    const unsigned pitch = (unsigned)(a.ldx * 2), opitch = (unsigned)(a.ldo * 4);
    auto row_ptr = [&](int c) -> const char* { return reinterpret_cast<const char*>(xl + (unsigned long long)(unsigned)c * pitch); };
    const unsigned long long keep = policy_evict_last(), once = policy_evict_first();
    const bool has_scale = a.row_scale != nullptr, has_bias = a.bias != nullptr, has_self = a.self_coef != 0.f;
    const bool has_acc = a.accumulate != 0, has_tail = a.tail_rowptr != nullptr;
    auto scale_of = [&](int d) -> float {
        return (has_scale && d < 0 && !(d & kDescPiece)) ? __ldg(a.row_scale + (d & kDescId)) : 1.0f;
    };
    const int4* cp = a.colp + 2 * b0;
    const float4* wp = WEIGHTED ? a.valp + 2 * b0 : nullptr;
    const int32_t* dp = a.desc + b0;
    int4 c0 = __ldg(cp), c1 = __ldg(cp + 1);
    int d_cur = __ldg(dp);
    int d_nxt = __ldg(dp + 1);
    float rs_cur = scale_of(d_cur);
    f8p acc{f4p_zero(), f4p_zero()};

    for (int it = 0; it < nb; ++it) {
        f4p v[8];
        v[0] = ldg_p_if(row_ptr(c0.x), c0.x, keep); v[1] = ldg_p_if(row_ptr(c0.y), c0.y, keep);
        v[2] = ldg_p_if(row_ptr(c0.z), c0.z, keep); v[3] = ldg_p_if(row_ptr(c0.w), c0.w, keep);
        v[4] = ldg_p_if(row_ptr(c1.x), c1.x, keep); v[5] = ldg_p_if(row_ptr(c1.y), c1.y, keep);
        v[6] = ldg_p_if(row_ptr(c1.z), c1.z, keep); v[7] = ldg_p_if(row_ptr(c1.w), c1.w, keep);
        float4 wc0 = make_float4(0.f, 0.f, 0.f, 0.f), wc1 = wc0;
        if (WEIGHTED) { wc0 = __ldg(wp); wc1 = __ldg(wp + 1); wp += 2; }
        cp += 2; dp += 1;
        c0 = __ldg(cp); c1 = __ldg(cp + 1);
        const int d_n2 = __ldg(dp + 1);
        const float rs_nxt = scale_of(d_nxt);
        if (WEIGHTED) {
            fma8(acc, wc0.x, v[0]); fma8(acc, wc0.y, v[1]); fma8(acc, wc0.z, v[2]); fma8(acc, wc0.w, v[3]);
            fma8(acc, wc1.x, v[4]); fma8(acc, wc1.y, v[5]); fma8(acc, wc1.z, v[6]); fma8(acc, wc1.w, v[7]);
        } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) add8(acc, v[u]);
        }
        if (d_cur < 0) {
            int row = d_cur & kDescId;
            float rs = rs_cur;
            bool write = true;
            float4 o0 = to_f4(acc.a), o1 = to_f4(acc.b);
            if (d_cur & kDescPiece) {
                const int piece = row;
                float* sp = a.scratch + (int64_t)piece * a.feat + sl * 8;
                stg4(sp, o0); stg4(sp + 4, o1);
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
                    o0 = make_float4(0.f, 0.f, 0.f, 0.f); o1 = o0;
                    for (int p = 0; p < np; ++p) {
                        const float4* q = reinterpret_cast<const float4*>(a.scratch + (int64_t)(p0 + p) * a.feat + sl * 8);
                        add4(o0, __ldcg(q)); add4(o1, __ldcg(q + 1));
                    }
                    row = __ldg(a.split_row + h);
                    rs = has_scale ? __ldg(a.row_scale + row) : 1.0f;
                }
            }
            if (write) {
                if (has_tail) {
                    const int k0 = __ldg(a.tail_rowptr + row), k1 = __ldg(a.tail_rowptr + row + 1);
                    for (int k = k0; k < k1; ++k) {
                        const f8p t = unpack_bf16x8(ldg_p(row_ptr(__ldg(a.tail_col + k)), keep));
                        const float w = __ldg(a.tail_val + k);
                        fma4(o0, w, to_f4(t.a)); fma4(o1, w, to_f4(t.b));
                    }
                }
                if (has_scale) { o0.x *= rs; o0.y *= rs; o0.z *= rs; o0.w *= rs; o1.x *= rs; o1.y *= rs; o1.z *= rs; o1.w *= rs; }
                if (has_self) {
                    const f8p t = unpack_bf16x8(ldg_p(row_ptr(row), keep));
                    fma4(o0, a.self_coef, to_f4(t.a)); fma4(o1, a.self_coef, to_f4(t.b));
                }
                if (has_bias) {
                    add4(o0, __ldg(reinterpret_cast<const float4*>(a.bias) + sl * 2));
                    add4(o1, __ldg(reinterpret_cast<const float4*>(a.bias) + sl * 2 + 1));
                }
                float4* op = reinterpret_cast<float4*>(ol + (unsigned long long)(unsigned)row * opitch);
                if (has_acc) { add4(o0, op[0]); add4(o1, op[1]); }
                stg4_hint(op, o0, once); stg4_hint(op + 1, o1, once);
            }
            acc.a = f4p_zero(); acc.b = f4p_zero();
        }
        d_cur = d_nxt; d_nxt = d_n2; rs_cur = rs_nxt;
    }
}

template <int LANES, bool WEIGHTED>
static int resident_workers16() {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spmm_batched_bf16_kernel<LANES, WEIGHTED>, 256, 0) != cudaSuccess) {
        cudaGetLastError();
        per_sm = 0;
    }
    if (per_sm <= 0) per_sm = 3;
    return kNumSMs * per_sm * 8 * (32 / LANES);
}

template <int LANES>
static int launch_batched16(const BArgs16& a, int64_t workers, bool weighted, cudaStream_t stream) {
    const int per_cta = 8 * (32 / LANES);
    const unsigned blocks = (unsigned)ceil_div<int64_t>(workers, per_cta);
    if (blocks == 0) return GD_OK;
    if (weighted) spmm_batched_bf16_kernel<LANES, true><<<blocks, 256, 0, stream>>>(a);
    else spmm_batched_bf16_kernel<LANES, false><<<blocks, 256, 0, stream>>>(a);
    GD_LAUNCH_CHECK();
    return GD_OK;
}

// fp32 -> bf16 (round to nearest even) of a row block, optionally scaled per row; 8 elements per thread
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int32_t feat,
                                                        const float* __restrict__ row_scale,
                                                        unsigned short* __restrict__ out, int64_t ldo) {
    const int per_row = feat >> 3;
    const int64_t total = rows * per_row;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / per_row;
        const int c = (int)(i - r * per_row) << 3;
        float4 a = __ldg(reinterpret_cast<const float4*>(x + r * ldx + c));
        float4 b = __ldg(reinterpret_cast<const float4*>(x + r * ldx + c + 4));
        if (row_scale) {
            const float s = __ldg(row_scale + r);
            a.x *= s; a.y *= s; a.z *= s; a.w *= s; b.x *= s; b.y *= s; b.z *= s; b.w *= s;
        }
        unsigned w0, w1, w2, w3;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w0) : "f"(a.y), "f"(a.x));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w1) : "f"(a.w), "f"(a.z));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w2) : "f"(b.y), "f"(b.x));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w3) : "f"(b.w), "f"(b.z));
        *reinterpret_cast<uint4*>(out + r * ldo + c) = make_uint4(w0, w1, w2, w3);
    }
}

__global__ void __launch_bounds__(256) move_f32_kernel(const float* __restrict__ src, const int32_t* __restrict__ src_idx,
                                                       float* __restrict__ dst, const int32_t* __restrict__ dst_idx, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[dst_idx ? dst_idx[i] : i] = src[src_idx ? src_idx[i] : i];
}

}  // namespace gd

using namespace gd;

extern "C" int32_t gd_spmm_batched_bf16_workers(int32_t feat, int32_t weighted) {
    switch (feat) {
        case 128: return weighted ? resident_workers16<16, true>() : resident_workers16<16, false>();
        case 64: return weighted ? resident_workers16<8, true>() : resident_workers16<8, false>();
        default: return 0;
    }
}

extern "C" int gd_spmm_batched_bf16(const gd_spmm_bplan_t* plan, const float* valp, const int32_t* tail_rowptr,
                                    const int32_t* tail_col, const float* tail_val, const float* row_scale,
                                    const void* x_bf16, int64_t ldx, int32_t feat, float self_coef, const float* bias,
                                    float* out, int64_t ldo, float* scratch, int32_t accumulate, gd_stream_t stream_) {
    cudaStream_t stream = as_stream(stream_);
    GD_CHECK_ARG(plan != nullptr, "null plan");
    GD_CHECK_ARG(feat == 64 || feat == 128, "feat must be 64 or 128");
    if (plan->num_rows == 0 || plan->num_batches == 0) return GD_OK;
    GD_CHECK_ARG(plan->desc && plan->colp && x_bf16 && out, "null pointer");
    GD_CHECK_ARG(plan->num_workers > 0 && plan->batches_per_worker > 0 &&
                     (int64_t)plan->num_workers * plan->batches_per_worker >= plan->num_batches, "inconsistent worker partition");
    GD_CHECK_ARG(plan->num_piece == 0 || (scratch && plan->piece_split && plan->split_row && plan->split_piece_beg &&
                                          plan->split_npiece && plan->split_ticket), "split rows without scratch / ticket arrays");
    GD_CHECK_ARG(ldx >= feat && ldo >= feat && ldx % 8 == 0 && ldo % 4 == 0, "leading dimensions: ldx multiple of 8, ldo of 4, >= feat");
    GD_CHECK_ARG((((uintptr_t)x_bf16 | (uintptr_t)out | (uintptr_t)scratch | (uintptr_t)bias | (uintptr_t)valp | (uintptr_t)plan->colp) % 16) == 0,
                 "operands must be 16-byte aligned");
    GD_CHECK_ARG(ldx * 2 < (int64_t)1 << 32 && ldo * 4 < (int64_t)1 << 32 && plan->num_rows < kDescId, "row pitch / row count out of range");
    GD_CHECK_ARG(!tail_rowptr || (tail_col && tail_val), "tail CSR without columns / values");
    BArgs16 a;
    a.desc = plan->desc; a.colp = reinterpret_cast<const int4*>(plan->colp); a.valp = reinterpret_cast<const float4*>(valp);
    a.row_scale = row_scale; a.x = x_bf16; a.bias = bias; a.out = out; a.scratch = scratch;
    a.piece_split = plan->piece_split; a.split_row = plan->split_row; a.split_piece_beg = plan->split_piece_beg;
    a.split_npiece = plan->split_npiece; a.split_ticket = plan->split_ticket;
    a.tail_rowptr = tail_rowptr; a.tail_col = tail_col; a.tail_val = tail_val;
    a.ldx = ldx; a.ldo = ldo;
    a.num_batches = (int32_t)plan->num_batches; a.per_worker = plan->batches_per_worker; a.feat = feat; a.accumulate = accumulate;
    a.self_coef = self_coef;
    const bool weighted = valp != nullptr;
    if (feat == 128) return launch_batched16<16>(a, plan->num_workers, weighted, stream);
    return launch_batched16<8>(a, plan->num_workers, weighted, stream);
}

extern "C" int gd_cast_bf16(const float* x, int64_t ldx, int64_t rows, int32_t feat, const float* row_scale,
                            void* out_bf16, int64_t ldo, gd_stream_t stream) {
    if (rows == 0) return GD_OK;
    GD_CHECK_ARG(x && out_bf16 && feat > 0 && feat % 8 == 0 && ldx % 4 == 0 && ldo % 8 == 0 && ldx >= feat && ldo >= feat, "bad shape");
    GD_CHECK_ARG((((uintptr_t)x | (uintptr_t)out_bf16) % 16) == 0, "operands must be 16-byte aligned");
    const int64_t total = rows * (feat >> 3);
    const int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(total, 256), kNumSMs * 16);
    cast_bf16_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x, ldx, rows, feat, row_scale, static_cast<unsigned short*>(out_bf16), ldo);
    GD_LAUNCH_CHECK();
    return GD_OK;
}

extern "C" int gd_move_f32(const float* src, const int32_t* src_idx, float* dst, const int32_t* dst_idx, int64_t n,
                           gd_stream_t stream) {
    if (n == 0) return GD_OK;
    GD_CHECK_ARG(src && dst && n > 0, "bad argument");
    const int blocks = (int)std::min<int64_t>(ceil_div<int64_t>(n, 256), kNumSMs * 16);
    move_f32_kernel<<<blocks, 256, 0, as_stream(stream)>>>(src, src_idx, dst, dst_idx, n);
    GD_LAUNCH_CHECK();
    return GD_OK;
}
