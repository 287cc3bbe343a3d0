// Input-gradient GEMM and weight-gradient contraction of a DeletionLayer in ONE kernel, chained through tensor memory:
//
//   dX[r, :]  = gate[r, :] (.) ( (s[r] x[r, :k]) . B )                      (k <= 64 -> n <= 128; r over the gathered rows)
//   c[k1, n]  = sum_r a[r, :k1]^T (x) dX[r, :n]                              (dW_del1 = a1[S1]^T . dX1[S1])
//
// Separately (gemm_rows_wt_kernel + gemm_tn_wt_kernel) dX is written to HBM and read back (2 x 4 m n bytes) and `a` is the
// only other large operand; here dX never leaves the SM.  The transposed accumulator of the weights-in-TMEM row GEMM,
// D[n features (lanes), rows (columns)], is exactly the A-operand layout [M x K = rows] of the weight gradient in its
// transposed form c^T[n, k1] = dX^T . a.  Per tile of 32 rows:
//   gather 16-byte cp.async copies of the rows of x and of a into a raw ring in shared memory, FW_RS tiles ahead: the bytes
//          in flight live in shared memory, not in registers (5 tiles = 120 KB per SM at k = 64, k1 = 128).  Every thread
//          reads back exactly the chunks it copied, so cp.async.wait_group is the only synchronisation of the ring
//   convert raw rows -> hi / lo operand tiles (x: K-major [x_hi | x_lo], row scale applied; a: MN-major [a_hi | a_lo])
//   MMA1   D = W (TMEM) x [x_hi | x_lo]^T                    main | corr, 32 columns each (as gemm_tc_wt.cu)
//   epilogue-1 (thread = feature, 32 rows): main + corr, gate bit, hi / lo split, tcgen05.st IN PLACE over D  ->  C_hi | C_lo
//   MMA2   acc[n, k1] += C (TMEM, K = 32 rows) x [a_hi | a_lo]
// D / C is double buffered (even / odd tiles, one group of four epilogue warps each): MMA1 of tile t + 2 is issued behind
// MMA2 of tile t in the in-order tensor pipe, so epilogue-1 of a tile overlaps the MMAs of its neighbours.
// The accumulators (main | corr, 2 k1 columns) are added to the CTA's partial [n][k1] every FW_FLUSH tiles with one vector
// reduction per 4 elements and thread (deterministic); tn_reduce_kernel adds the partials in CTA order and transposes.
// TMEM: W 2k <= 128 columns | D / C 2 x 64 | acc 256.
//   warps 0-3   x gather + conversion        warps 4-11  a gather + conversion
//   warps 12-19 epilogue-1 (warp % 4 = TMEM lane quarter, (warp - 12) / 4 = D / C buffer = tile parity; they stage W first)
//   warps 20-23 accumulator drain            warp  24    MMA issue
#include "tc_common.cuh"

namespace gd {
namespace tc {

constexpr int FW_ROWS = 32;
constexpr int FW_X_WARPS = 4, FW_A_WARPS = 8, FW_E_WARPS = 8, FW_D_WARPS = 4;
constexpr int FW_E_WARP0 = FW_X_WARPS + FW_A_WARPS, FW_D_WARP0 = FW_E_WARP0 + FW_E_WARPS;
constexpr int FW_MMA_WARP = FW_D_WARP0 + FW_D_WARPS;
constexpr int FW_THREADS = (FW_MMA_WARP + 1) * 32;                  // 800
constexpr int FW_FLUSH = 16;                     // tiles per accumulator flush: 512 rows = 64 k-steps of MMA2
constexpr int FW_RS = 5;                         // raw ring stages (tiles of gathered rows in flight)
constexpr int FW_SC_BYTES = FW_X_WARPS * 32 * 2 * 4;    // per raw stage: the row scales, one private copy per x thread and row
constexpr int FW_OPS = 2;                        // operand stages (converted tiles) of x and of a
constexpr int FW_XATOM = 2 * FW_ROWS * 128;      // one k-atom of an x stage: [32 rows hi | 32 rows lo] x 128 B
constexpr int FW_ATOM_COL = 4096;                // one 32-feature atom column of a 32-row block of `a`
constexpr uint32_t FW_COL_D = 128, FW_COL_ACC = 256;
constexpr int FW_TMEM_COLS = 512;
static_assert(FW_E_WARP0 % 4 == 0 && FW_D_WARP0 % 4 == 0, "warp % 4 must be the TMEM lane quarter");

struct DxDwArgs {
    const float* x; int64_t ldx;                 // [*, k]
    const float* b; int b_is_nk; int k, n;       // dX = x . B
    const float* in_scale;                       // per row (optional)
    const uint32_t* gate_bits;                   // [row][n / 32] (optional)
    const float* a; int64_t lda; int k1;         // [*, k1]
    const int32_t* rows; int64_t m;
    float* partial;                              // [gridDim.x][n][k1]
    int64_t rows_per_cta;
};

// 16-byte (4-byte) asynchronous copy global -> shared; src_bytes = 0 fills the destination with zeros
__device__ __forceinline__ void cp_async16(void* dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int KCH>
__global__ void __launch_bounds__(FW_THREADS, 1) gemm_dxdw_wt_kernel(const DxDwArgs t) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // by offset: keeps the shared address space
    constexpr int x_stage_bytes = KCH * FW_XATOM;
    constexpr int x_row_bytes = KCH * KC * 4;
    const int a_tile = t.k1 * 128;                                  // bytes of 32 rows of a_hi (or a_lo)
    const int a_stage_bytes = 2 * a_tile;
    const int a_row_bytes = t.k1 * 4;
    const int raw_stage_bytes = FW_ROWS * (x_row_bytes + a_row_bytes) + FW_SC_BYTES;    // [32 rows of x][32 rows of a][scales]
    uint8_t* a_smem = smem + FW_OPS * x_stage_bytes;
    uint8_t* raw_smem = a_smem + FW_OPS * a_stage_bytes;
    __shared__ uint64_t xfull[FW_OPS], xempty[FW_OPS], afull[FW_OPS], aempty[FW_OPS];
    __shared__ uint64_t d_full[2], c_ready[2], acc_full, acc_free, w_bar;
    __shared__ uint32_t tmem_base_smem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < FW_OPS; ++s) {
            mbar_init(&xfull[s], FW_X_WARPS); mbar_init(&xempty[s], 1);
            mbar_init(&afull[s], FW_A_WARPS); mbar_init(&aempty[s], 1);
        }
        for (int b = 0; b < 2; ++b) { mbar_init(&d_full[b], 1); mbar_init(&c_ready[b], FW_E_WARPS / 2); }
        mbar_init(&acc_full, 1); mbar_init(&acc_free, FW_D_WARPS);
        mbar_init(&w_bar, FW_E_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == FW_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"(FW_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    const int64_t r_beg = (int64_t)blockIdx.x * t.rows_per_cta;
    const int64_t r_end = min(t.m, r_beg + t.rows_per_cta);
    const int ntiles = r_end > r_beg ? (int)((r_end - r_beg + FW_ROWS - 1) / FW_ROWS) : 0;
    auto row_id = [&](int64_t i) -> int32_t { return (i >= r_beg && i < r_end) ? (t.rows ? __ldg(t.rows + i) : (int32_t)i) : -1; };

    if (warp < FW_X_WARPS) {
        // ------------------------------ x: gather, raw rows -> K-major [x_hi | x_lo] ------------------------------
        // thread = (rows r0 and r0 + 16 of the tile, 16-byte chunk j of every k-atom): a quarter-warp copies / reads the 128
        // contiguous bytes of a row atom and stores one conflict-free swizzled row
        const int j = tid & 7, r0 = tid >> 3;
        int32_t rid_q[2];                                           // row ids of the next tile to gather
        auto fetch = [&](int tl) {
#pragma unroll
            for (int b = 0; b < 2; ++b) rid_q[b] = tl < ntiles ? row_id(r_beg + (int64_t)tl * FW_ROWS + r0 + 16 * b) : -1;
        };
        auto issue = [&](int tl) {                                  // copies of tile tl (rows past the end: zero fill), one group
            uint8_t* rw = raw_smem + (tl % FW_RS) * raw_stage_bytes;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int32_t r = rid_q[b];
                const float* src = t.x + (int64_t)(r >= 0 ? r : 0) * t.ldx + j * 4;
#pragma unroll
                for (int p = 0; p < KCH; ++p)
                    cp_async16(rw + (r0 + 16 * b) * x_row_bytes + p * 128 + j * 16, src + p * KC, r >= 0 ? 16u : 0u);
                if (t.in_scale)
                    cp_async4(rw + FW_ROWS * (x_row_bytes + a_row_bytes) + (tid * 2 + b) * 4, t.in_scale + (r >= 0 ? r : 0), r >= 0 ? 4u : 0u);
            }
            cp_async_commit();
        };
        {
            int32_t rid_p[FW_RS][2];
#pragma unroll
            for (int d = 0; d < FW_RS; ++d) { fetch(d); rid_p[d][0] = rid_q[0]; rid_p[d][1] = rid_q[1]; }
#pragma unroll
            for (int d = 0; d < FW_RS; ++d) { rid_q[0] = rid_p[d][0]; rid_q[1] = rid_p[d][1]; issue(d); }
        }
        fetch(FW_RS);
        uint32_t xs = 0, xph = 0;
        for (int tl = 0; tl < ntiles; ++tl) {
            const int32_t rid_u[2] = {rid_q[0], rid_q[1]};          // ids of tile tl + FW_RS (gathered at the end of this iteration),
            fetch(tl + FW_RS + 1);                                   // fetched an iteration ago: the id load is never waited for
            cp_async_wait<FW_RS - 1>();                              // this thread's copies of tile tl have landed
            const uint8_t* rw = raw_smem + (tl % FW_RS) * raw_stage_bytes;
            float4 v[2][KCH];
            float sc[2] = {1.0f, 1.0f};
#pragma unroll
            for (int b = 0; b < 2; ++b) {
#pragma unroll
                for (int p = 0; p < KCH; ++p) v[b][p] = *reinterpret_cast<const float4*>(rw + (r0 + 16 * b) * x_row_bytes + p * 128 + j * 16);
                if (t.in_scale) sc[b] = *reinterpret_cast<const float*>(rw + FW_ROWS * (x_row_bytes + a_row_bytes) + (tid * 2 + b) * 4);
            }
            mbar_wait(&xempty[xs], xph ^ 1);
            uint8_t* st = smem + xs * x_stage_bytes;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
#pragma unroll
                for (int p = 0; p < KCH; ++p) {
                    float4 y = v[b][p];
                    y.x *= sc[b]; y.y *= sc[b]; y.z *= sc[b]; y.w *= sc[b];
                    float4 hi, lo;
                    split4(y, hi, lo);
                    const uint32_t o = p * FW_XATOM + swz(r0 + 16 * b, j);
                    *reinterpret_cast<float4*>(st + o) = hi;
                    *reinterpret_cast<float4*>(st + o + FW_XATOM / 2) = lo;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&xfull[xs]);                  // (no proxy fence on this side: the MMA thread fences after acquiring)
            if (++xs == (uint32_t)FW_OPS) { xs = 0; xph ^= 1; }
            {                                                        // refill the slot just read
                const int32_t nx[2] = {rid_q[0], rid_q[1]};
                rid_q[0] = rid_u[0]; rid_q[1] = rid_u[1];
                issue(tl + FW_RS);
                rid_q[0] = nx[0]; rid_q[1] = nx[1];
            }
        }
        cp_async_wait<0>();
    } else if (warp < FW_E_WARP0) {
        // ------------------------------ a: gather, raw rows -> MN-major [a_hi | a_lo] ------------------------------
        // 8 lanes cover the 128 B of one row of one 32-feature atom, a warp 4 rows, the 8 warps the 32 rows of a tile
        const int aw = warp - FW_X_WARPS;
        const int rr = aw * 4 + (lane >> 3), cj = lane & 7;
        const int a4 = t.k1 >> 2;
        const uint32_t row_off = (uint32_t)((rr >> 2) * 512 + (rr & 3) * 128 + ((((cj >> 1) ^ (rr & 3)) << 5) | ((cj & 1) << 4)));
        const int raw_off = FW_ROWS * x_row_bytes + rr * a_row_bytes + cj * 16;
        int32_t rid_q;
        auto fetch = [&](int tl) { rid_q = tl < ntiles ? row_id(r_beg + (int64_t)tl * FW_ROWS + rr) : -1; };
        auto issue = [&](int tl) {
            uint8_t* rw = raw_smem + (tl % FW_RS) * raw_stage_bytes + raw_off;
            const int32_t r = rid_q;
            const float* src = t.a + (int64_t)(r >= 0 ? r : 0) * t.lda + cj * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (cj + 8 * i < a4) cp_async16(rw + i * 128, src + i * 32, r >= 0 ? 16u : 0u);
            cp_async_commit();
        };
        {
            int32_t rid_p[FW_RS];
#pragma unroll
            for (int d = 0; d < FW_RS; ++d) { fetch(d); rid_p[d] = rid_q; }
#pragma unroll
            for (int d = 0; d < FW_RS; ++d) { rid_q = rid_p[d]; issue(d); }
        }
        fetch(FW_RS);
        uint32_t as = 0, aph = 0;
        for (int tl = 0; tl < ntiles; ++tl) {
            const int32_t rid_u = rid_q;                             // id of tile tl + FW_RS, fetched an iteration ago
            fetch(tl + FW_RS + 1);
            cp_async_wait<FW_RS - 1>();
            const uint8_t* rw = raw_smem + (tl % FW_RS) * raw_stage_bytes + raw_off;
            float4 y[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                y[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (cj + 8 * i < a4) y[i] = *reinterpret_cast<const float4*>(rw + i * 128);
            }
            mbar_wait(&aempty[as], aph ^ 1);
            uint8_t* st = a_smem + as * a_stage_bytes + row_off;
#pragma unroll
            for (int i = 0; i < 4; ++i) {                            // atom i = features 32 i .. 32 i + 31, 4 KB apart
                if (cj + 8 * i < a4) {
                    float4 hi, lo;
                    split4(y[i], hi, lo);
                    *reinterpret_cast<float4*>(st + i * FW_ATOM_COL) = hi;
                    *reinterpret_cast<float4*>(st + a_tile + i * FW_ATOM_COL) = lo;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&afull[as]);
            if (++as == (uint32_t)FW_OPS) { as = 0; aph ^= 1; }
            {
                const int32_t nx = rid_q;
                rid_q = rid_u;
                issue(tl + FW_RS);
                rid_q = nx;
            }
        }
        cp_async_wait<0>();
    } else if (warp == FW_MMA_WARP) {
        // ------------------------------ MMA issue ------------------------------
        if (elect_one()) {
            const uint32_t id1_2r = make_idesc(2 * FW_ROWS), id1_r = make_idesc(FW_ROWS);                        // B K-major, A from TMEM
            const uint32_t id2_n = make_idesc_mn(t.k1) & ~(1u << 15), id2_2n = make_idesc_mn(2 * t.k1) & ~(1u << 15);   // B MN-major
            const uint64_t x0 = make_desc(smem_u32(smem));
            const uint64_t a0 = make_desc_mn(smem_u32(a_smem), FW_ATOM_COL, 512);
            const uint32_t w_hi = tmem_base, w_lo = tmem_base + t.k;
            const uint32_t acc = tmem_base + FW_COL_ACC;
            uint32_t xs = 0, xph = 0, as = 0, aph = 0;
            // D (buffer tl & 1) = W x [x_hi | x_lo]^T; the buffer is free: MMA2 of tile tl - 2 precedes this in the in-order pipe
            auto mma1 = [&](int tl) {
                mbar_wait(&xfull[xs], xph);
                fence_proxy_async();                                  // the converters' generic stores -> async proxy
                tc_fence_after();
                const uint32_t dD = tmem_base + FW_COL_D + (tl & 1) * 2 * FW_ROWS;
                const uint64_t xd = desc_advance(x0, xs * x_stage_bytes);
#pragma unroll
                for (int p = 0; p < KCH; ++p) {
#pragma unroll
                    for (int ks = 0; ks < KC / 8; ++ks) {
                        const uint64_t b = xd + (uint64_t)(p * (FW_XATOM >> 4) + 2 * ks);
                        const uint32_t ka = p * KC + ks * 8;
                        umma_tf32_tmem_a(dD, w_hi + ka, b, id1_2r, (p | ks) != 0);        // main | corr (+)= W_hi x [x_hi | x_lo]
                        umma_tf32_tmem_a(dD + FW_ROWS, w_lo + ka, b, id1_r, 1);           // corr += W_lo x x_hi
                    }
                }
                umma_commit(&xempty[xs]);
                umma_commit(&d_full[tl & 1]);
                if (++xs == (uint32_t)FW_OPS) { xs = 0; xph ^= 1; }
            };
            mbar_wait(&w_bar, 0);
            tc_fence_after();
            if (ntiles > 0) mma1(0);
            if (ntiles > 1) mma1(1);
            for (int tl = 0; tl < ntiles; ++tl) {
                // ---- MMA2: acc += C (TMEM, buffer tl & 1) x [a_hi | a_lo]
                const int grp = tl / FW_FLUSH;
                const bool first = (tl % FW_FLUSH) == 0;
                mbar_wait(&c_ready[tl & 1], (uint32_t)((tl >> 1) & 1));
                mbar_wait(&afull[as], aph);
                if (first && grp > 0) mbar_wait(&acc_free, (uint32_t)((grp - 1) & 1));
                fence_proxy_async();
                tc_fence_after();
                const uint32_t c_hi = tmem_base + FW_COL_D + (tl & 1) * 2 * FW_ROWS, c_lo = c_hi + FW_ROWS;
                const uint64_t ad = desc_advance(a0, as * a_stage_bytes);
#pragma unroll
                for (int ks = 0; ks < FW_ROWS / 8; ++ks) {           // 8 rows per k-step: two 4-row k-atoms (1 KB) of the block
                    const uint64_t b = desc_advance(ad, ks * 1024);
                    umma_tf32_tmem_a(acc, c_hi + 8 * ks, b, id2_2n, !(first && ks == 0));     // main | corr (+)= C_hi x [a_hi | a_lo]
                    umma_tf32_tmem_a(acc + t.k1, c_lo + 8 * ks, b, id2_n, 1);                 // corr += C_lo x a_hi
                }
                umma_commit(&aempty[as]);
                if ((tl % FW_FLUSH) == FW_FLUSH - 1 || tl == ntiles - 1) umma_commit(&acc_full);
                if (++as == (uint32_t)FW_OPS) { as = 0; aph ^= 1; }
                if (tl + 2 < ntiles) mma1(tl + 2);
            }
        }
        __syncwarp();
    } else if (warp < FW_D_WARP0) {
        // ------------------------------ epilogue-1: D -> gate -> C_hi | C_lo in place ------------------------------
        const int q = warp & 3, ebuf = (warp - FW_E_WARP0) >> 2;
        const int f = q * 32 + lane;                                // dX feature = TMEM lane
        const bool active = q * 32 < t.n;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        const int nw = t.n >> 5;
        // ---- W -> tensor memory, once (the two warps of a lane quarter split the k range)
        {
            for (int c16 = ebuf; c16 < t.k / 16; c16 += 2) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int kk = c16 * 16 + e;
                    float w = 0.f;
                    if (f < t.n) w = t.b_is_nk ? __ldg(t.b + (int64_t)f * t.k + kk) : __ldg(t.b + (int64_t)kk * t.n + f);
                    float h, l;
                    split_tf32(w, h, l);
                    hi[e] = __float_as_uint(h); lo[e] = __float_as_uint(l);
                }
                tmem_st16(t_lane + 16 * c16, hi);
                tmem_st16(t_lane + t.k + 16 * c16, lo);
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&w_bar);
        }
        const uint32_t tD = t_lane + FW_COL_D + ebuf * 2 * FW_ROWS;
        // gate word of row `lane` of the tile, feature group q: the row id is fetched two of this group's tiles ahead, the
        // word one tile ahead, so neither load is waited for
        const bool gated = t.gate_bits && active;
        int32_t rid_n = -1;
        uint32_t gw_n = 0xffffffffu;
        auto fetch_rid = [&](int tl) { rid_n = (gated && tl < ntiles) ? row_id(r_beg + (int64_t)tl * FW_ROWS + lane) : -1; };
        auto fetch_gate = [&]() { gw_n = gated ? (rid_n >= 0 ? __ldg(t.gate_bits + (int64_t)rid_n * nw + q) : 0u) : 0xffffffffu; };
        fetch_rid(ebuf); fetch_gate(); fetch_rid(ebuf + 2);
        for (int tl = ebuf; tl < ntiles; tl += 2) {
            const uint32_t gw = gw_n;
            fetch_gate();                                            // tile tl + 2
            fetch_rid(tl + 4);
            mbar_wait(&d_full[ebuf], (uint32_t)((tl >> 1) & 1));
            tc_fence_after();
            if (active) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t pm[16], pc[16];
                    tmem_ld16_nowait(tD + 16 * c, pm);
                    tmem_ld16_nowait(tD + FW_ROWS + 16 * c, pc);
                    tmem_wait_ld();
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        float v = __uint_as_float(pm[e]) + __uint_as_float(pc[e]);
                        const uint32_t wj = __shfl_sync(0xffffffffu, gw, 16 * c + e);   // gate word of row 16 c + e
                        if (!((wj >> lane) & 1u)) v = 0.f;
                        float h, l;
                        split_tf32(v, h, l);
                        pm[e] = __float_as_uint(h); pc[e] = __float_as_uint(l);
                    }
                    tmem_st16(tD + 16 * c, pm);                        // C_hi over D main, C_lo over D corr (in place)
                    tmem_st16(tD + FW_ROWS + 16 * c, pc);
                }
                tmem_wait_st();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&c_ready[ebuf]);
        }
    } else {
        // ------------------------------ drain: add a flush group to the CTA's partial ------------------------------
        const int q = warp & 3;
        const int f = q * 32 + lane;
        const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + FW_COL_ACC;
        float* part = t.partial + (int64_t)blockIdx.x * t.k1 * t.n + (int64_t)f * t.k1;       // row f of [n][k1]
        const int ngroups = (ntiles + FW_FLUSH - 1) / FW_FLUSH;
        for (int grp = 0; grp < ngroups; ++grp) {
            mbar_wait(&acc_full, (uint32_t)(grp & 1));
            tc_fence_after();
            if (f < t.n) {
                for (int c0 = 0; c0 < t.k1; c0 += 16) {
                    uint32_t vm[16], vc[16];
                    tmem_ld16_nowait(t_acc + c0, vm);
                    tmem_ld16_nowait(t_acc + t.k1 + c0, vc);
                    tmem_wait_ld();
#pragma unroll
                    for (int e = 0; e < 16; e += 4) {
                        const float4 o = make_float4(__uint_as_float(vm[e]) + __uint_as_float(vc[e]), __uint_as_float(vm[e + 1]) + __uint_as_float(vc[e + 1]),
                                                     __uint_as_float(vm[e + 2]) + __uint_as_float(vc[e + 2]), __uint_as_float(vm[e + 3]) + __uint_as_float(vc[e + 3]));
                        // every element of the CTA's partial is updated by this one thread, group after group: a vector reduction
                        // is deterministic here and, unlike a read-modify-write, does not wait for the old value
                        float* dst = part + c0 + e;
                        if (grp > 0) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
                        else *reinterpret_cast<float4*>(dst) = o;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_free);
        }
        if (ngroups == 0 && f < t.n)                                // CTA without rows: its partial is zero
            for (int c = 0; c < t.k1; ++c) part[c] = 0.f;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == FW_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(FW_TMEM_COLS));
    }
}

constexpr size_t FW_SMEM_LIMIT = 227 * 1024 - 2048;
static size_t fw_smem_bytes(int k, int k1) {
    return 1024 + (size_t)FW_OPS * ((size_t)(k / KC) * FW_XATOM + (size_t)k1 * 256) + (size_t)FW_RS * ((size_t)FW_ROWS * 4 * (k + k1) + FW_SC_BYTES);
}

}  // namespace tc
}  // namespace gd

using namespace gd;

extern "C" int gd_gemm_dxdw_tc_supported(int32_t k, int32_t n, int32_t k1, int64_t ldx, int64_t lda) {
    if (k <= 0 || k > 64 || k % tc::KC != 0 || n <= 0 || n > 128 || n % 32 != 0 || k1 <= 0 || k1 > 128 || k1 % 32 != 0) return 0;
    if (ldx % 4 != 0 || lda % 4 != 0) return 0;
    return tc::fw_smem_bytes(k, k1) <= tc::FW_SMEM_LIMIT ? 1 : 0;
}

extern "C" int gd_gemm_dxdw_tc(const float* x, int64_t ldx, const float* b, int32_t b_is_nk, int32_t k, int32_t n,
                               const float* in_scale, const uint32_t* gate_bits, const float* a, int64_t lda, int32_t k1,
                               const int32_t* rows, int64_t m, float* c, void* workspace, size_t workspace_bytes,
                               gd_stream_t stream_) {
    cudaStream_t stream = as_stream(stream_);
    GD_CHECK_ARG(m >= 0 && c != nullptr, "bad argument");
    if (m == 0) { GD_CUDA(cudaMemsetAsync(c, 0, (size_t)k1 * n * sizeof(float), stream)); return GD_OK; }
    GD_CHECK_ARG(x && b && a, "null pointer");
    GD_CHECK_ARG(gd_gemm_dxdw_tc_supported(k, n, k1, ldx, lda), "shape not supported by the fused dX / dW kernel");
    GD_CHECK_ARG(((uintptr_t)x | (uintptr_t)a) % 16 == 0, "operands must be 16-byte aligned");
    if (!workspace || workspace_bytes < gd_gemm_tn_tc_workspace_bytes(k1, n))
        return fail(GD_ERR_WORKSPACE, "gd_gemm_dxdw_tc: workspace too small (gd_gemm_tn_tc_workspace_bytes(k1, n))");
    tc::DxDwArgs t{x, ldx, b, b_is_nk, k, n, in_scale, gate_bits, a, lda, k1, rows, m, static_cast<float*>(workspace), 0};
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(kNumSMs, ceil_div<int64_t>(m, 256)));
    t.rows_per_cta = ceil_div<int64_t>(ceil_div<int64_t>(m, grid), 64) * 64;
    const size_t smem = tc::fw_smem_bytes(k, k1);
    auto launch = [&](auto kern) -> int {
        GD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GD_CUDA(launch_pdl(kern, grid, tc::FW_THREADS, smem, stream, t));
        GD_LAUNCH_CHECK();
        return GD_OK;
    };
    const int rc = (k / tc::KC == 1) ? launch(tc::gemm_dxdw_wt_kernel<1>) : launch(tc::gemm_dxdw_wt_kernel<2>);
    if (rc != GD_OK) return rc;
    return tc::launch_tn_reduce(t.partial, grid, (int64_t)k1 * n, c, stream, k1, n);      // partials are [n][k1]
}
