// Ceiling probe 2: random-row gathers through the async-copy engines instead of LDG.
//   variant 0: LDG.128 baseline (sub-warp per row, 8 loads in flight per lane)
//   variant 1: cp.async.bulk (1-D bulk copy, one row per lane, mbarrier complete_tx) -> smem ring -> LDS.128 reduce
//   variant 2: cp.async.bulk.tensor.2d tile::gather4 (4 rows per instruction, tensor map) -> smem ring -> LDS.128 reduce
// Every variant sums all gathered rows per lane slice and the host checks the grand total, so a variant that
// moves the wrong bytes is reported instead of timed.
// usage: gather_bench2 <variant> <F> <rows> <powerlaw 0|1> [stages] [warps]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(cnt)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* tm, int c0, int r0, int r1, int r2, int r3, uint64_t* b) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(b)) : "memory");
}

// ------------------------------------------------------------------ variant 0
template <int LANES, int UNROLL>
__global__ void gather_ldg(const float4* __restrict__ x, const int* __restrict__ idx, long n_idx, int row_f4, float* __restrict__ out) {
    long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
    long grp = t / LANES; int sl = t % LANES;
    long ngrp = (long)gridDim.x * blockDim.x / LANES;
    float4 acc = make_float4(0, 0, 0, 0);
    for (long i = grp * UNROLL; i + UNROLL <= n_idx; i += ngrp * UNROLL) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = __ldg(x + (long)idx[i + u] * row_f4 + sl);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    out[t] = acc.x + acc.y + acc.z + acc.w;
}

// ------------------------------------------------------------------ variants 1 / 2
// Each warp is its own producer and consumer: ROWS rows per stage, STAGES stages per warp.
template <int F, int ROWS, bool TENSOR>
__global__ void gather_async(const float* __restrict__ x, const __grid_constant__ CUtensorMap tm, const int* __restrict__ idx,
                             long n_idx, int stages, float* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int LPR = F / 4;                   // lanes per row in the consume phase
    constexpr int RPI = 32 / LPR;                // rows per LDS.128 instruction
    constexpr uint32_t STAGE_BYTES = ROWS * F * 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem) + warp * 8;             // up to 8 stages
    unsigned char* ring = smem + nwarp * 64 + (size_t)warp * stages * STAGE_BYTES;
    ring = (unsigned char*)(((uintptr_t)ring + 127) & ~(uintptr_t)127);
    if (lane == 0) for (int s = 0; s < stages; ++s) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const long gw = (long)blockIdx.x * nwarp + warp, nw = (long)gridDim.x * nwarp;
    const long nchunk = n_idx / ROWS;

    auto load_idx = [&](long chunk) -> int4 {
        int4 r = make_int4(0, 0, 0, 0);
        if (chunk >= nchunk) return r;
        if (!TENSOR) { if (lane < ROWS) r.x = __ldg(idx + chunk * ROWS + lane); }
        else if (lane < ROWS / 4) r = __ldg(reinterpret_cast<const int4*>(idx + chunk * ROWS) + lane);
        return r;
    };
    auto issue = [&](int4 r, int s) {
        if (lane == 0) mbar_expect_tx(bars + s, STAGE_BYTES);
        __syncwarp();
        unsigned char* dst = ring + (size_t)s * STAGE_BYTES;
        if (!TENSOR) {
            if (lane < ROWS) bulk_g2s(dst + lane * F * 4, x + (long)r.x * F, F * 4, bars + s);
        } else {
            if (lane < ROWS / 4) tma_gather4(dst + lane * 4 * F * 4, &tm, 0, r.x, r.y, r.z, r.w, bars + s);
        }
    };

    long c = gw;
    for (int s = 0; s < stages; ++s) if (c + (long)s * nw < nchunk) issue(load_idx(c + (long)s * nw), s);
    int4 nidx = load_idx(c + (long)stages * nw);
    float4 acc = make_float4(0, 0, 0, 0);
    int s = 0; uint32_t phase = 0;
    for (; c < nchunk; c += nw) {
        mbar_wait(bars + s, phase);
        const float4* p = reinterpret_cast<const float4*>(ring + (size_t)s * STAGE_BYTES) + lane;
#pragma unroll
        for (int j = 0; j < ROWS / RPI; ++j) {
            const float4 v = p[j * 32];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        __syncwarp();
        const long nxt = c + (long)stages * nw;
        if (nxt < nchunk) issue(nidx, s);
        nidx = load_idx(nxt + nw);
        if (++s == stages) { s = 0; phase ^= 1; }
    }
    atomicAdd(out, acc.x + acc.y + acc.z + acc.w);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int f = argc > 2 ? atoi(argv[2]) : 64;
    const long rows = argc > 3 ? atol(argv[3]) : 235368;
    const int powerlaw = argc > 4 ? atoi(argv[4]) : 0;
    const int stages = argc > 5 ? atoi(argv[5]) : 4;
    const int warps = argc > 6 ? atoi(argv[6]) : 4;
    const int boxrows = argc > 7 ? atoi(argv[7]) : 1;
    const long n_idx = 2800000L * 4 / 128 * 128;
    std::vector<int> h(n_idx);
    std::vector<float> hx((size_t)rows * f);
    srand(1);
    for (size_t i = 0; i < hx.size(); ++i) hx[i] = (float)((i * 2654435761u >> 20) & 7) * 0.125f;
    std::vector<double> rowsum(rows, 0.0);
    for (long r = 0; r < rows; ++r) { double s = 0; for (int k = 0; k < f; ++k) s += hx[(size_t)r * f + k]; rowsum[r] = s; }
    double expect = 0;
    for (auto& v : h) {
        double u = (rand() + 0.5) / ((double)RAND_MAX + 1.0);
        long r = powerlaw ? (long)(rows * pow(u, 2.5)) : (long)(((long)rand() * 32768 + rand()) % rows);
        if (r >= rows) r = rows - 1;
        v = (int)r; expect += rowsum[r];
    }
    int* idx; float *x, *out;
    CK(cudaMalloc(&idx, n_idx * 4)); CK(cudaMalloc(&x, (size_t)rows * f * 4)); CK(cudaMalloc(&out, 4 << 20)); CK(cudaMemset(out, 0, 4 << 20));
    CK(cudaMemcpy(x, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);

    CUtensorMap tm; memset(&tm, 0, sizeof(tm));
    if (variant == 2) {
        void* fn = nullptr; cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
        cuuint64_t gdim[2] = {(cuuint64_t)f, (cuuint64_t)rows};
        cuuint64_t gstr[1] = {(cuuint64_t)f * 4};
        cuuint32_t box[2] = {(cuuint32_t)f, (cuuint32_t)boxrows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
    }
    auto report = [&](const char* name, int blocks, float ms) {
        float got = 0;
        if (variant == 0) { std::vector<float> ho(1 << 20); CK(cudaMemcpy(ho.data(), out, 4 << 20, cudaMemcpyDeviceToHost)); double t = 0; for (float v : ho) t += v; got = (float)t; }
        else CK(cudaMemcpy(&got, out, 4, cudaMemcpyDeviceToHost));
        const double rel = fabs(got - expect) / expect;
        printf("v%d rows=%ld F=%d %s %s blocks=%d: %.1f us  %.2f TB/s gathered  check %s (rel %.2e)\n", variant, rows, f,
               powerlaw ? "powerlaw" : "uniform", name, blocks, ms * 1e3, n_idx * (double)f * 4 / ms / 1e9, rel < 1e-3 ? "ok" : "WRONG", rel);
    };
    if (variant == 0) {
        for (int blocks : {148 * 4, 148 * 8, 148 * 16}) {
            auto launch = [&] {
                if (f == 64) gather_ldg<16, 8><<<blocks, 256>>>((const float4*)x, idx, n_idx, f / 4, out);
                else gather_ldg<32, 8><<<blocks, 256>>>((const float4*)x, idx, n_idx, f / 4, out);
            };
            launch(); CK(cudaDeviceSynchronize());
            cudaEventRecord(a);
            for (int r = 0; r < 5; ++r) launch();
            cudaEventRecord(b); CK(cudaEventSynchronize(b));
            float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
            CK(cudaMemset(out, 0, 4 << 20)); launch(); CK(cudaDeviceSynchronize());
            report("ldg x8", blocks, ms);
        }
        return 0;
    }
    const int rper = argc > 8 ? atoi(argv[8]) : 32;
    const size_t smem = (size_t)warps * 64 + (size_t)warps * stages * rper * f * 4 + 128;
    for (int per_sm : {1, 2, 3, 4}) {
        if (smem * per_sm > 220 * 1024) continue;
        const int blocks = 148 * per_sm;
        auto go = [&](auto kern) {
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<blocks, warps * 32, smem>>>(x, tm, idx, n_idx, stages, out);
        };
        auto launch = [&] {
            if (rper == 32) {
                if (f == 64 && variant == 1) go(gather_async<64, 32, false>);
                else if (f == 64) go(gather_async<64, 32, true>);
                else if (variant == 1) go(gather_async<128, 32, false>);
                else go(gather_async<128, 32, true>);
            } else {
                if (f == 64 && variant == 1) go(gather_async<64, 16, false>);
                else if (f == 64) go(gather_async<64, 16, true>);
                else if (variant == 1) go(gather_async<128, 16, false>);
                else go(gather_async<128, 16, true>);
            }
        };
        launch(); CK(cudaDeviceSynchronize());
        cudaEventRecord(a);
        for (int r = 0; r < 5; ++r) launch();
        cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
        CK(cudaMemset(out, 0, 4)); launch(); CK(cudaDeviceSynchronize());
        char name[64]; snprintf(name, sizeof name, "%s stages=%d warps=%d rows/stage=%d", variant == 1 ? "bulk1d" : "gather4", stages, warps, rper);
        report(name, blocks, ms);
    }
    return 0;
}
