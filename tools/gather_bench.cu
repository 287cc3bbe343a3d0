// Ceiling probe: how fast can B200 gather random 256 B / 512 B rows from an L2-resident / HBM-resident matrix
// when nothing else limits (indices precomputed, 8 independent 128-bit loads in flight per lane)?
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
template <int LANES, int UNROLL>
__global__ void gather(const float4* __restrict__ x, const int* __restrict__ idx, long n_idx, int row_f4, float4* __restrict__ out) {
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
    if (acc.x == 12345.f) out[t] = acc;
}
int main() {
    for (int pass = 0; pass < 3; ++pass) {
        long rows = pass == 2 ? 8000000 : 235368; int f = pass == 1 ? 128 : 64;
        long n_idx = 2800000 * 4;
        std::vector<int> h(n_idx);
        srand(1);
        for (auto& v : h) v = (int)(((long)rand() * 32768 + rand()) % rows);
        int* idx; float4 *x, *out;
        cudaMalloc(&idx, n_idx * 4); cudaMalloc(&x, rows * f * 4); cudaMalloc(&out, 1 << 20);
        cudaMemset(x, 0, rows * f * 4);
        cudaMemcpy(idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        auto run = [&](auto kern, int lanes, const char* name) {
            for (int blocks : {148 * 4, 148 * 8, 148 * 16}) {
                kern<<<blocks, 256>>>(x, idx, n_idx, f / 4, out);
                cudaEventRecord(a);
                for (int r = 0; r < 5; ++r) kern<<<blocks, 256>>>(x, idx, n_idx, f / 4, out);
                cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
                printf("rows=%ld F=%d %s blocks=%d: %.1f us  %.2f TB/s gathered\n", rows, f, name, blocks, ms * 1e3, n_idx * (double)f * 4 / ms / 1e9);
            }
        };
        if (f == 64) { run(gather<16, 8>, 16, "16 lanes x8"); run(gather<16, 16>, 16, "16 lanes x16"); }
        else { run(gather<32, 8>, 32, "32 lanes x8"); run(gather<32, 16>, 32, "32 lanes x16"); }
        cudaFree(idx); cudaFree(x); cudaFree(out);
    }
    return 0;
}
