// Shared pieces of the batched aggregation kernels (spmm_batched.cu, node_loss.cu): plan descriptor bits, the kernel
// argument block, packed-fp32 helpers and the L2-hinted gather loads.
#pragma once
#include "common.cuh"

namespace gd {

constexpr int kDescFlush = (int)0x80000000u;
constexpr int kDescPiece = 0x40000000;
constexpr int kDescId = 0x3fffffff;

struct BArgs {
    const int32_t* desc;
    const int4* colp;
    const float4* valp;
    const float* row_scale;
    const float* x;
    const float* bias;
    float* out;
    float* scratch;
    const int32_t* piece_split;
    const int32_t* split_row;
    const int32_t* split_piece_beg;
    const int32_t* split_npiece;
    int32_t* split_ticket;
    const int32_t* tail_rowptr;      // optional second (plain CSR) operand added to every row at its flush
    const int32_t* tail_col;
    const float* tail_val;
    int64_t ldx, ldo;
    int32_t num_batches, per_worker, feat, accumulate;
    float self_coef;
};

// packed fp32 pairs (sm_100 FADD2 / FFMA2): a feature fragment of 4 floats is two 64-bit registers
struct f4p { unsigned long long lo, hi; };
__device__ __forceinline__ f4p f4p_zero() { return f4p{0ull, 0ull}; }
__device__ __forceinline__ void add_p(f4p& acc, const f4p& v) {
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc.lo) : "l"(v.lo));
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc.hi) : "l"(v.hi));
}
__device__ __forceinline__ void fma_p(f4p& acc, float w, const f4p& v) {
    unsigned long long ww;
    asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc.lo) : "l"(ww), "l"(v.lo));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc.hi) : "l"(ww), "l"(v.hi));
}
// L2 eviction hints: gathered source rows are re-read by other rows' neighbourhoods (evict_last), the output rows
// and the plan streams are touched once (evict_first) - at F = 128 the source matrix (120 MB) only just fits the
// 126 MB L2 and the streaming traffic was evicting it (ncu: 430 MB of DRAM reads for 120 MB of compulsory bytes).
__device__ __forceinline__ unsigned long long policy_evict_last() {
    unsigned long long p; asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ unsigned long long policy_evict_first() {
    unsigned long long p; asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ f4p ldg_p(const char* p, unsigned long long pol) {
    f4p r;
    asm("ld.global.nc.L2::cache_hint.v2.b64 {%0, %1}, [%2], %3;" : "=l"(r.lo), "=l"(r.hi) : "l"(p), "l"(pol));
    return r;
}
// zero when the slot is padding (c < 0); the address is computed unconditionally (one IMAD.WIDE) and only the
// load is predicated - written in PTX because the compiler otherwise predicates (and re-derives) the whole
// 64-bit address computation per slot
__device__ __forceinline__ f4p ldg_p_if(const char* p, int c, unsigned long long pol) {
    f4p r;
    asm("{\n.reg .pred q;\nsetp.ge.s32 q, %3, 0;\nmov.b64 %0, 0;\nmov.b64 %1, 0;\n@q ld.global.nc.L2::cache_hint.v2.b64 {%0, %1}, [%2], %4;\n}"
        : "=&l"(r.lo), "=&l"(r.hi) : "l"(p), "r"(c), "l"(pol));
    return r;
}
__device__ __forceinline__ void stg4_hint(float4* p, const float4& v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ float4 to_f4(const f4p& v) {
    float4 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v.lo));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.z), "=f"(r.w) : "l"(v.hi));
    return r;
}

}  // namespace gd
