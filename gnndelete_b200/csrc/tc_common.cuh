// tcgen05 / TMEM / mbarrier building blocks shared by the tensor-core kernels (gemm_tc.cu, dense_ni_tc.cu): PTX wrappers,
// shared-memory matrix descriptors, the 3xTF32 operand split and the SWIZZLE_128B tile addressing.  sm_100a only.
#pragma once
#include "common.cuh"

namespace gd {
namespace tc {

constexpr int BM = 128;            // rows per tile (UMMA M)
constexpr int KC = 32;             // k per stage: 32 tf32 = one 128-byte swizzle row
constexpr int TILE_BYTES = BM * 128;            // one [128 rows x 128 B] operand tile

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
// explicit 128-bit shared load (the compiler split the float4 dereference of the transpose tile into two LDS.64,
// which breaks the quarter-warp conflict-free pattern the tile is laid out for)
__device__ __forceinline__ float4 lds128(const float* p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
    return v;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100): rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);            // start address            bits [0,14)
    d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset       bits [32,46)
    d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                             // layout: SWIZZLE_128B
    return d;
}
// MN-major descriptor for 32-bit operands.  tf32 MN-major has exactly one legal shared-memory layout,
// SWIZZLE_128B_BASE32B (layout type 1): atoms of 4 k-rows x 128 B (32 tf32 contiguous along M / N), the 32-byte
// chunk c of k-row i stored at chunk c ^ (i & 3) (Swizzle<2,5,2> on the byte address).  lbo = byte stride between
// atoms along M / N, sbo = between atoms along K; one K = 8 instruction reads two k-atoms.  (The 16-byte-granular
// SWIZZLE_128B layout is accepted by the assembler for MN-major tf32 but the MMA then produces zeros.)
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}
// The descriptors of one kernel differ only in their start-address field (bits [0,14) = shared address >> 4; every
// operand tile lies below 256 KB, so adding (byte offset >> 4) to the low word never carries into another field):
// the MMA thread builds one descriptor per operand and advances it with integer additions.  Re-deriving every
// descriptor from its address (shift / mask / or on the uniform datapath, 4 per k-step) made the single issuing
// thread the slowest stage of the pipeline (profiles/r1_gemm_phase_knobs.md: ~0.45 us per stage).
__device__ __forceinline__ uint64_t desc_advance(uint64_t d, uint32_t byte_off) { return d + (uint64_t)(byte_off >> 4); }
// One lane of a converged warp (elect.sync): unlike `lane == 0`, the compiler knows a single thread is active inside
// and issues the uniform-datapath tcgen05 instructions without an election loop around each of them.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128
__device__ __forceinline__ uint32_t make_idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
// same with A and B MN-major (instruction descriptor bits 15 / 16)
__device__ __forceinline__ uint32_t make_idesc_mn(int n) { return make_idesc(n) | (1u << 15) | (1u << 16); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// same with the A operand in tensor memory (M along the 128 lanes, K along 32-bit columns: k-step s starts at column 8 s)
__device__ __forceinline__ void umma_tf32_tmem_a(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// hi = v rounded to tf32 (10-bit mantissa), lo = the remainder rounded to tf32.
// Round-to-nearest (ties away from zero, what cvt.rna.tf32.f32 does) on the integer pipe: add half an
// ulp of the 10-bit mantissa to the magnitude bits and clear the 13 low bits.  cvt.rna.tf32.f32 itself
// issues on the quarter-rate conversion pipe: with 2 conversions per element the producers spent
// ~1500 cycles per 128 x 32 stage on it and bounded the whole kernel (tools/gemm_sweep.py, GD_TC_DEBUG=30).
__device__ __forceinline__ float round_tf32(float v) {
    return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
    hi = round_tf32(v);
    lo = round_tf32(v - hi);
}
__device__ __forceinline__ void split4(const float4& v, float4& hi, float4& lo) {
    split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y);
    split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
}
// byte offset of 16-byte chunk j of row r inside a [rows x 128 B] SWIZZLE_128B tile
__device__ __forceinline__ uint32_t swz(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
// registers -> TMEM: thread = TMEM lane (32 (warp % 4) + lane), 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


// arguments of the gathered-row GEMMs (gemm_tc.cu: A ring + resident B in shared memory; gemm_tc_wt.cu: weights in TMEM)
struct Args {
    const float* a; int64_t lda;
    const int32_t* rows; int64_t m; int k;
    const float* b; int b_is_nk; int n;
    const float* bias; const float* out_scale; const float* gate; int64_t ldgate;
    int relu_in, relu_out;
    float* out; int64_t ldo;
    uint32_t* relu_mask_out; const uint32_t* gate_bits;      // [row][n/32] bit c%32 of word c/32 <=> value > 0
    int num_tiles;
    int stages;
    int debug;          // GD_TC_DEBUG bit mask (measurement only): 1 = no epilogue work, 2 = producers do not load, 4 = no MMAs, 8 = epilogue without global stores, 16 = epilogue without TMEM loads
};

// arguments of the weight-gradient contractions  c[k1, n2] = sum_i a_scale[r(i)] pro(a[r(i), :k1])^T (x) g[r(i), :n2]
struct TnArgs {
    const float* a; int64_t lda;
    const float* g; int64_t ldg;
    const int32_t* rows; int64_t m;
    int k1, n2, relu_a;
    const float* a_scale;
    float* partial;                 // [gridDim.x][k1][n2]
    int64_t rows_per_cta;
    int stages;
};

bool tn_wt_supported(const TnArgs& t);
int launch_tn_wt(const TnArgs& t, cudaStream_t stream, int* nparts);     // fills t.partial [nparts][k1][n2]

// out = sum of the partials in CTA order; tr_k1 > 0: partials are [tr_n][tr_k1], out is [tr_k1][tr_n]
int launch_tn_reduce(const float* partial, int nparts, int64_t count, float* out, cudaStream_t stream, int tr_k1 = 0, int tr_n = 0);

bool rows_wt_supported(const Args& g);
int launch_rows_wt(const Args& g, cudaStream_t stream);

}  // namespace tc
}  // namespace gd
