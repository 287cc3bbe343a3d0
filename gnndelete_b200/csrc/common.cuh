// Shared helpers for libgnndelete_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <atomic>
#include <string>
#include <utility>

#include "../../include/gnndelete_b200.h"

namespace gd {

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs; grids are sized in multiples of this

extern std::atomic<long long> g_launches;   // kernels this library has launched (capi.cu)
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define GD_CHECK_ARG(cond, msg)                                              \
    do {                                                                     \
        if (!(cond)) return ::gd::fail(GD_ERR_INVALID, std::string(__func__) + ": " + (msg)); \
    } while (0)

#define GD_CUDA(expr)                                                        \
    do {                                                                     \
        cudaError_t e__ = (expr);                                            \
        if (e__ != cudaSuccess)                                              \
            return ::gd::fail(GD_ERR_CUDA, std::string(__func__) + ": " #expr ": " + cudaGetErrorString(e__)); \
    } while (0)

#define GD_LAUNCH_CHECK()                                                    \
    do {                                                                     \
        ::gd::g_launches.fetch_add(1, std::memory_order_relaxed);            \
        GD_CUDA(cudaGetLastError());                                         \
    } while (0)

inline cudaStream_t as_stream(gd_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// ---- kernel launch helper ------------------------------------------------------------
// Programmatic dependent launch across the epoch's kernels was measured in round 1 (917 instead of 1090 epochs/s on the
// Collab epoch: early-resident dependents take SM slots from the primary's last CTAs) and removed; launch_pdl is a plain
// stream-ordered launch, pdl_wait / pdl_trigger compile to nothing.
__device__ __forceinline__ void pdl_wait() {}
__device__ __forceinline__ void pdl_trigger() {}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cfg.attrs = nullptr;
    cfg.numAttrs = 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}

// ---- device helpers ---------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// sum within aligned groups of G lanes (G power of two <= 32)
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 128-bit read-only gather of a feature-row fragment.  Gathered rows are re-read by
// other rows' neighbourhoods, so they are allowed to allocate in L1 (default policy).
__device__ __forceinline__ float4 ldg4(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}
// streaming 128-bit store (outputs are written once and consumed by a later kernel)
__device__ __forceinline__ void stg4(float* p, float4 v) {
    *reinterpret_cast<float4*>(p) = v;
}
__device__ __forceinline__ void fma4(float4& acc, float a, const float4& v) {
    acc.x = fmaf(a, v.x, acc.x); acc.y = fmaf(a, v.y, acc.y);
    acc.z = fmaf(a, v.z, acc.z); acc.w = fmaf(a, v.w, acc.w);
}
__device__ __forceinline__ void add4(float4& acc, const float4& v) {
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
}

}  // namespace gd
