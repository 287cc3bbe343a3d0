// Batched CSR aggregation: the same contract as gd_spmm on a BATCH PLAN built once per edge set.
//
// Why: tools/gather_bench.cu shows that a loop of "8 independent 128-bit row gathers per lane,
// then add" sustains 16-18 TB/s of L2-resident row gathers on B200, while the row-pipelined
// kernel in spmm.cu reaches 5.2-5.7 TB/s: ncu (profiles/r1_spmm_pipe_ncu_full.md) shows it
// executing ~23 warp instructions per non-zero (row bookkeeping, per-element predicates, shuffles)
// with the issue slots 50-64 % busy, i.e. it is instruction bound, not bandwidth bound.
//
// The batch plan removes all row bookkeeping from the kernel:
//   * every row is cut into BATCHES of 8 column slots (the last one padded with -1; an empty row
//     is one all-padding batch), stored in a padded column array colp[num_batches][8] so that the
//     slots of batch b are two aligned 128-bit loads whose address depends on b only;
//   * desc[b] says whether the batch ends its row (FLUSH) and which row that is;
//   * the flat batch list is cut into `num_workers` contiguous ranges of EQUAL length, one per
//     resident sub-warp (F/4 lanes), so the load balance is exact by construction whatever the
//     degree distribution.  A row that straddles a range boundary becomes PIECES: partial sums go
//     to scratch and the sub-warp that completes the row's last piece (ticket counter) adds them
//     in piece order - deterministic, no float atomics, a 7000-neighbour hub is simply spread
//     over ~25 workers.
// Per batch a sub-warp issues 8 independent LDG.128 (the column ids, descriptor and row scale of
// the NEXT batch are already in flight), adds, and flushes when the descriptor says so.
#include "spmm_batched.cuh"

namespace gd {

// Instruction budget matters: ncu on the first version of this kernel showed 16 warp instructions per
// non-zero with the issue slots 55-63 % busy.  Hence: one IMAD.WIDE per gather address (byte offset =
// column x row pitch added to a 64-bit lane base), packed FADD2 / FFMA2 accumulation, and - when a warp
// holds more than one sub-warp - a single predicated gather path, because a full / partial branch that
// the sub-warps of a warp take differently executes both sides.
// FLUSH_* select, at compile time, what the end-of-row code does; FLUSH_ANY keeps every runtime test.
enum : int { FLUSH_SCALE = 1, FLUSH_BIAS = 2, FLUSH_SELF = 4, FLUSH_ACC = 8, FLUSH_ANY = 16, FLUSH_TAIL = 32 };

template <int LANES, bool WEIGHTED, int FL>
__global__ void __launch_bounds__(256, 4) spmm_batched_kernel(const BArgs a) {
    constexpr int PER_WARP = 32 / LANES;
    const int lane = threadIdx.x & 31;
    const int sub = lane / LANES, sl = lane % LANES;
    const unsigned mask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << (sub * LANES));
    const int64_t worker = ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5) * PER_WARP + sub;
    pdl_wait();
    pdl_trigger();
    const int64_t b0 = worker * a.per_worker;
    if (b0 >= a.num_batches) return;
    const int nb = (int)min((int64_t)a.per_worker, (int64_t)a.num_batches - b0);
    unsigned long long xl = reinterpret_cast<unsigned long long>(a.x) + sl * 16;   // this lane's 16 bytes of every source row
    unsigned long long ol = reinterpret_cast<unsigned long long>(a.out) + sl * 16; // ... and of every output row
    asm volatile("" : "+l"(xl), "+l"(ol));       // opaque: keeps base + lane offset in ONE register pair (the IMAD.WIDE addend)
    const unsigned pitch = (unsigned)(a.ldx * 4), opitch = (unsigned)(a.ldo * 4);
    auto row_ptr = [&](int c) -> const char* { return reinterpret_cast<const char*>(xl + (unsigned long long)(unsigned)c * pitch); };
    const unsigned long long keep = policy_evict_last(), once = policy_evict_first();
    const bool has_scale = (FL & FLUSH_ANY) ? a.row_scale != nullptr : (FL & FLUSH_SCALE) != 0;
    const bool has_bias = (FL & FLUSH_ANY) ? a.bias != nullptr : (FL & FLUSH_BIAS) != 0;
    const bool has_self = (FL & FLUSH_ANY) ? a.self_coef != 0.f : (FL & FLUSH_SELF) != 0;
    const bool has_acc = (FL & FLUSH_ANY) ? a.accumulate != 0 : (FL & FLUSH_ACC) != 0;
    const bool has_tail = (FL & FLUSH_ANY) ? a.tail_rowptr != nullptr : (FL & FLUSH_TAIL) != 0;

    auto scale_of = [&](int d) -> float {      // row scale of a batch that flushes a whole row
        return (has_scale && d < 0 && !(d & kDescPiece)) ? __ldg(a.row_scale + (d & kDescId)) : 1.0f;
    };

    // running pointers; the plan arrays carry two batches of slack, so "the next batch" can be loaded unconditionally
    const int4* cp = a.colp + 2 * b0;
    const float4* wp = WEIGHTED ? a.valp + 2 * b0 : nullptr;
    const int32_t* dp = a.desc + b0;
    int4 c0 = __ldg(cp), c1 = __ldg(cp + 1);
    int d_cur = __ldg(dp);
    int d_nxt = __ldg(dp + 1);
    float rs_cur = scale_of(d_cur);
    f4p acc = f4p_zero();

    for (int it = 0; it < nb; ++it) {
        // ---- 8 independent row gathers (padding, -1, is at the end of a row's last batch)
        f4p v[8];
        if (LANES == 32 && c1.w >= 0) {         // whole warp on one batch: the full-batch test is uniform
            v[0] = ldg_p(row_ptr(c0.x), keep); v[1] = ldg_p(row_ptr(c0.y), keep); v[2] = ldg_p(row_ptr(c0.z), keep); v[3] = ldg_p(row_ptr(c0.w), keep);
            v[4] = ldg_p(row_ptr(c1.x), keep); v[5] = ldg_p(row_ptr(c1.y), keep); v[6] = ldg_p(row_ptr(c1.z), keep); v[7] = ldg_p(row_ptr(c1.w), keep);
        } else {
            v[0] = ldg_p_if(row_ptr(c0.x), c0.x, keep); v[1] = ldg_p_if(row_ptr(c0.y), c0.y, keep);
            v[2] = ldg_p_if(row_ptr(c0.z), c0.z, keep); v[3] = ldg_p_if(row_ptr(c0.w), c0.w, keep);
            v[4] = ldg_p_if(row_ptr(c1.x), c1.x, keep); v[5] = ldg_p_if(row_ptr(c1.y), c1.y, keep);
            v[6] = ldg_p_if(row_ptr(c1.z), c1.z, keep); v[7] = ldg_p_if(row_ptr(c1.w), c1.w, keep);
        }
        // the slot weights are a coalesced load whose address depends on the batch only: issued with the gathers they
        // arrive with them, and not double-buffering them keeps the weighted kernel at 64 registers (4 CTAs / SM)
        float4 wc0 = make_float4(0.f, 0.f, 0.f, 0.f), wc1 = wc0;
        if (WEIGHTED) { wc0 = __ldg(wp); wc1 = __ldg(wp + 1); wp += 2; }
        // ---- next batch: column ids, the descriptor after it, its row scale
        cp += 2; dp += 1;
        c0 = __ldg(cp); c1 = __ldg(cp + 1);
        const int d_n2 = __ldg(dp + 1);
        const float rs_nxt = scale_of(d_nxt);
        // ---- accumulate
        if (WEIGHTED) {
            fma_p(acc, wc0.x, v[0]); fma_p(acc, wc0.y, v[1]); fma_p(acc, wc0.z, v[2]); fma_p(acc, wc0.w, v[3]);
            fma_p(acc, wc1.x, v[4]); fma_p(acc, wc1.y, v[5]); fma_p(acc, wc1.z, v[6]); fma_p(acc, wc1.w, v[7]);
        } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) add_p(acc, v[u]);
        }
        // ---- end of a row (or of this worker's piece of it)
        if (d_cur < 0) {
            int row = d_cur & kDescId;
            float rs = rs_cur;
            bool write = true;
            float4 o = to_f4(acc);
            if (d_cur & kDescPiece) {
                const int piece = row;
                stg4(a.scratch + (int64_t)piece * a.feat + sl * 4, o);
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
                    o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
                    for (int p = 0; p < np; ++p)
                        add4(o, __ldcg(reinterpret_cast<const float4*>(a.scratch + (int64_t)(p0 + p) * a.feat) + sl));
                    row = __ldg(a.split_row + h);
                    rs = has_scale ? __ldg(a.row_scale + row) : 1.0f;
                }
            }
            if (write) {
                if (has_tail) {                  // the few entries of the second CSR (this step's negative pairs)
                    const int k0 = __ldg(a.tail_rowptr + row), k1 = __ldg(a.tail_rowptr + row + 1);
                    for (int k = k0; k < k1; ++k)
                        fma4(o, __ldg(a.tail_val + k), __ldg(reinterpret_cast<const float4*>(row_ptr(__ldg(a.tail_col + k)))));
                }
                if (has_scale) { o.x *= rs; o.y *= rs; o.z *= rs; o.w *= rs; }
                if (has_self) fma4(o, a.self_coef, __ldg(reinterpret_cast<const float4*>(row_ptr(row))));
                if (has_bias) add4(o, __ldg(reinterpret_cast<const float4*>(a.bias) + sl));
                float4* op = reinterpret_cast<float4*>(ol + (unsigned long long)(unsigned)row * opitch);
                if (has_acc) add4(o, *op);
                stg4_hint(op, o, once);
            }
            acc = f4p_zero();
        }
        d_cur = d_nxt; d_nxt = d_n2; rs_cur = rs_nxt;
    }
}

template <int LANES, bool WEIGHTED>
static int resident_workers() {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spmm_batched_kernel<LANES, WEIGHTED, FLUSH_ANY>, 256, 0) != cudaSuccess) {
        cudaGetLastError();
        per_sm = 0;
    }
    if (per_sm <= 0) per_sm = 4;                 // no device in this process (plan built for a later run): 64 registers
    return kNumSMs * per_sm * 8 * (32 / LANES);
}

template <int LANES, bool WEIGHTED>
static int launch_flags(const BArgs& a, unsigned blocks, cudaStream_t stream) {
    const int need = (a.row_scale ? FLUSH_SCALE : 0) | (a.bias ? FLUSH_BIAS : 0) | (a.self_coef != 0.f ? FLUSH_SELF : 0) |
                     (a.accumulate ? FLUSH_ACC : 0) | (a.tail_rowptr ? FLUSH_TAIL : 0);
    // the flush variants of the Del-training epoch are compiled without the unused tests: GCN forward
    // (scale + bias), transpose-backward / loss gather (nothing), loss gather + this step's negatives (tail);
    // everything else takes the generic variant
    if (need == 0) GD_CUDA(launch_pdl(spmm_batched_kernel<LANES, WEIGHTED, 0>, blocks, 256, 0, stream, a));
    else if (need == (FLUSH_SCALE | FLUSH_BIAS)) GD_CUDA(launch_pdl(spmm_batched_kernel<LANES, WEIGHTED, FLUSH_SCALE | FLUSH_BIAS>, blocks, 256, 0, stream, a));
    else if (need == FLUSH_TAIL) GD_CUDA(launch_pdl(spmm_batched_kernel<LANES, WEIGHTED, FLUSH_TAIL>, blocks, 256, 0, stream, a));
    else GD_CUDA(launch_pdl(spmm_batched_kernel<LANES, WEIGHTED, FLUSH_ANY>, blocks, 256, 0, stream, a));
    GD_LAUNCH_CHECK();
    return GD_OK;
}

template <int LANES>
static int launch_batched(const BArgs& a, int64_t workers, bool weighted, cudaStream_t stream) {
    const int per_cta = 8 * (32 / LANES);
    const unsigned blocks = (unsigned)ceil_div<int64_t>(workers, per_cta);
    if (blocks == 0) return GD_OK;
    return weighted ? launch_flags<LANES, true>(a, blocks, stream) : launch_flags<LANES, false>(a, blocks, stream);
}

}  // namespace gd

using namespace gd;

extern "C" int32_t gd_spmm_batched_workers(int32_t feat, int32_t weighted) {
    switch (feat) {
        case 128: return weighted ? resident_workers<32, true>() : resident_workers<32, false>();
        case 64: return weighted ? resident_workers<16, true>() : resident_workers<16, false>();
        case 32: return weighted ? resident_workers<8, true>() : resident_workers<8, false>();
        default: return 0;
    }
}

extern "C" int gd_spmm_batched(const gd_spmm_bplan_t* plan, const float* valp, const float* row_scale, const float* x,
                               int64_t ldx, int32_t feat, float self_coef, const float* bias, float* out, int64_t ldo,
                               float* scratch, int32_t accumulate, gd_stream_t stream_) {
    return gd_spmm_batched_tail(plan, valp, nullptr, nullptr, nullptr, row_scale, x, ldx, feat, self_coef, bias, out, ldo, scratch,
                                accumulate, stream_);
}

extern "C" int gd_spmm_batched_tail(const gd_spmm_bplan_t* plan, const float* valp, const int32_t* tail_rowptr,
                                    const int32_t* tail_col, const float* tail_val, const float* row_scale, const float* x,
                                    int64_t ldx, int32_t feat, float self_coef, const float* bias, float* out, int64_t ldo,
                                    float* scratch, int32_t accumulate, gd_stream_t stream_) {
    cudaStream_t stream = as_stream(stream_);
    GD_CHECK_ARG(plan != nullptr, "null plan");
    GD_CHECK_ARG(feat == 32 || feat == 64 || feat == 128, "feat must be 32, 64 or 128");
    if (plan->num_rows == 0 || plan->num_batches == 0) return GD_OK;
    GD_CHECK_ARG(plan->desc && plan->colp && x && out, "null pointer");
    GD_CHECK_ARG(plan->num_workers > 0 && plan->batches_per_worker > 0 &&
                     (int64_t)plan->num_workers * plan->batches_per_worker >= plan->num_batches, "inconsistent worker partition");
    GD_CHECK_ARG(plan->num_piece == 0 || (scratch && plan->piece_split && plan->split_row && plan->split_piece_beg &&
                                          plan->split_npiece && plan->split_ticket), "split rows without scratch / ticket arrays");
    GD_CHECK_ARG(ldx >= feat && ldo >= feat && ldx % 4 == 0 && ldo % 4 == 0 && ldo * 4 < (int64_t)1 << 32,
                 "leading dimensions must be multiples of 4 and >= feat");
    GD_CHECK_ARG((((uintptr_t)x | (uintptr_t)out | (uintptr_t)scratch | (uintptr_t)bias | (uintptr_t)valp | (uintptr_t)plan->colp) % 16) == 0,
                 "operands must be 16-byte aligned");
    GD_CHECK_ARG(ldx * 4 < (int64_t)1 << 32 && plan->num_rows < kDescId, "row pitch / row count out of range");
    GD_CHECK_ARG(!tail_rowptr || (tail_col && tail_val), "tail CSR without columns / values");
    BArgs a;
    a.desc = plan->desc; a.colp = reinterpret_cast<const int4*>(plan->colp); a.valp = reinterpret_cast<const float4*>(valp);
    a.row_scale = row_scale; a.x = x; a.bias = bias; a.out = out; a.scratch = scratch;
    a.piece_split = plan->piece_split; a.split_row = plan->split_row; a.split_piece_beg = plan->split_piece_beg;
    a.split_npiece = plan->split_npiece; a.split_ticket = plan->split_ticket;
    a.tail_rowptr = tail_rowptr; a.tail_col = tail_col; a.tail_val = tail_val;
    a.ldx = ldx; a.ldo = ldo;
    a.num_batches = (int32_t)plan->num_batches; a.per_worker = plan->batches_per_worker; a.feat = feat; a.accumulate = accumulate;
    a.self_coef = self_coef;
    const bool weighted = valp != nullptr;
    if (feat == 128) return launch_batched<32>(a, plan->num_workers, weighted, stream);
    if (feat == 64) return launch_batched<16>(a, plan->num_workers, weighted, stream);
    return launch_batched<8>(a, plan->num_workers, weighted, stream);
}
