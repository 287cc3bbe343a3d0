/*
 * libgnndelete_b200.so — C ABI of the B200 (sm_100a) kernels behind GNNDelete's
 * unlearning hot path.
 *
 * The reference (mims-harvard/GNNDelete) has no FFI boundary of its own: its hot
 * path is Python that calls PyTorch-Geometric operators.  Each
 * entry point below therefore cites the reference call site / PyG operator whose
 * arithmetic it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - all buffers (inputs, outputs, workspaces) are owned by the caller;
 *   - feature matrices are row-major fp32 with an explicit leading dimension;
 *   - node / edge ids are int32 inside the library; the COO builders take the
 *     reference's int64 `edge_index` rows;
 *   - every call is asynchronous on `stream` (a cudaStream_t) and re-entrant;
 *   - return 0 on success, negative on error; `gd_last_error()` gives the text of
 *     the calling thread's last failure.
 */
#ifndef GNNDELETE_B200_H
#define GNNDELETE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* gd_stream_t; /* cudaStream_t */

#define GD_OK 0
#define GD_ERR_INVALID (-1)
#define GD_ERR_CUDA (-2)
#define GD_ERR_WORKSPACE (-3)

int gd_version(void);
const char* gd_last_error(void);
/* kernels launched by this library since load (process-wide; evidence for bench.py's
 * gpu_launches). */
long long gd_launch_count(void);

/* ------------------------------------------------------------------ graph build
 * Replaces the per-call index plumbing of torch_geometric MessagePassing
 * (`add_remaining_self_loops`, `gcn_norm`, gather/scatter index handling) used by
 * GCNConv/GATConv/GINConv/RGCNConv at framework/models/gcn.py:11-12, gat.py:11-12,
 * gin.py:11-12, rgcn.py:17-22.  Built once per (edge set) and cached by the host.
 */

/* Destination-major CSR view of a message-passing edge set. */
typedef struct gd_csr {
    int64_t num_rows;           /* destinations */
    int64_t nnz;
    const int32_t* rowptr;      /* [num_rows + 1] */
    const int32_t* col;         /* [nnz] source ids; ascending (relation, source) within a row */
    /* long-row split plan from gd_spmm_plan_build (all zero / NULL: no splitting) */
    int32_t seg_len;
    int32_t num_heavy;
    int32_t num_seg;
    const int32_t* heavy_row;     /* [num_heavy] */
    const int32_t* heavy_seg_beg; /* [num_heavy] first segment of the row */
    const int32_t* heavy_nseg;    /* [num_heavy] */
    const int32_t* seg_row;       /* [num_seg] */
    const int32_t* seg_beg;       /* [num_seg] absolute offset into col */
    const int32_t* seg_heavy;     /* [num_seg] index of the segment's row in heavy_* */
    int32_t* heavy_ticket;        /* [num_heavy] zero-initialised; self re-arming completion counters */
    /* optional visiting order of the rows (e.g. degree-sorted inside windows); NULL = identity */
    const int32_t* row_perm;      /* [num_rows] */
    /* reserved (row groups of a removed streaming kernel): NULL / 0 */
    const int32_t* grp_row;       /* [num_grp + 1] */
    int32_t num_grp;
} gd_csr_t;

size_t gd_csr_workspace_bytes(int64_t num_edges, int64_t num_nodes);

/* COO (int64 rows of the reference's edge_index: src = edge_index[0], dst =
 * edge_index[1]) -> CSR sorted by (dst, [rel,] src).
 *   self_loops = 1: PyG add_remaining_self_loops — existing (v,v) entries are
 *                   dropped and one (v,v) per node is inserted (GCNConv, GATConv);
 *   self_loops = 0: entries kept as they are (GINConv, RGCNConv, loss incidence);
 *   self_loops = 2: as 0, but only the ROWS are ordered (stable sort on the destination alone, fewer radix passes):
 *                   entries of a row keep their input order.  For per-step structures such as the negative-pair
 *                   incidence, where any fixed order is as good as the sorted one.  No relations.
 * `rel` (nullable, values in [0, num_rel)) adds the relation as a secondary sort key
 * and is written per entry to `rel_out`.
 * `eid[k]` = column of the input COO that landed at CSR position k, or
 * num_edges + v for the inserted self loop of node v.  nnz = rowptr[num_nodes]; `col`/`eid`/`rel_out` must hold
 * num_edges (+ num_nodes when self_loops) entries.
 * `status` (device int32[2]) receives {nnz, number of out-of-range endpoints}. */
int gd_csr_from_coo(const int64_t* src, const int64_t* dst, const int64_t* rel, int64_t num_edges,
                    int64_t num_nodes, int32_t num_rel, int32_t self_loops, int32_t* rowptr,
                    int32_t* col, int32_t* eid, int32_t* rel_out, int32_t* status, void* workspace,
                    size_t workspace_bytes, gd_stream_t stream);

/* dst[perm[k]] = k for k in [0,n) where perm[k] >= 0. */
int gd_invert_perm(const int32_t* perm, int64_t n, int32_t* inv, gd_stream_t stream);

/* gcn_norm (GCNConv defaults, gcn.py:11-12): dinv[i] = deg(i)^-1/2 with
 * deg = in-degree incl. the inserted self loop (unit weights), inf -> 0. */
int gd_gcn_dinv(const int32_t* rowptr, int64_t num_nodes, float* dinv, gd_stream_t stream);

/* Long-row split plan: rows with more than seg_len entries are cut into segments of
 * seg_len so a power-law hub never serialises on one warp.  Array capacities:
 * heavy_* >= nnz / seg_len + 1, seg_* >= 2 * nnz / seg_len + 2.
 * `counts` (device int32[2]) receives {num_heavy, num_seg}. */
int gd_spmm_plan_build(const int32_t* rowptr, int64_t num_rows, int32_t seg_len, int32_t* heavy_row,
                       int32_t* heavy_seg_beg, int32_t* heavy_nseg, int32_t* seg_row,
                       int32_t* seg_beg, int32_t* seg_heavy, int32_t* counts, gd_stream_t stream);

/* ------------------------------------------------------------- aggregation (1)
 * out[i,:] = row_scale[i] * ( sum_{k in row i} val[k] * col_scale[col[k]] * x[col[k],:] )
 *            + self_coef * x[i,:] + bias[:]
 * val / col_scale / row_scale / bias nullable.  Covers
 *   GCNConv propagate  (A_hat = D^-1/2 (A+I) D^-1/2 via row/col scale), gcn.py:11-12;
 *   GINConv propagate  (unit weights, self_coef = 1 + eps), gin.py:11-12;
 *   their transpose-backward (same call on the transposed CSR);
 *   the loss backward  (val = d loss / d logit per incidence entry).
 * `scratch` holds num_seg * feat floats when the CSR carries a split plan. */
int gd_spmm(const gd_csr_t* csr, const float* val, const float* col_scale, const float* row_scale,
            const float* x, int64_t ldx, int32_t feat, float self_coef, const float* bias,
            float* out, int64_t ldo, float* scratch, gd_stream_t stream);

/* gd_spmm with out += result when `accumulate` is non-zero (the loss gradient is the sum of the
 * gather over the fixed pairs and the gather over this step's negative pairs). */
int gd_spmm_acc(const gd_csr_t* csr, const float* val, const float* col_scale, const float* row_scale,
                const float* x, int64_t ldx, int32_t feat, float self_coef, const float* bias,
                float* out, int64_t ldo, float* scratch, int32_t accumulate, gd_stream_t stream);

/* Batch plan of the same aggregation (gnndelete_b200/csrc/spmm_batched.cu): every row is cut into
 * batches of 8 column slots (padded with -1), the flat batch list is cut into `num_workers` equal
 * contiguous ranges (one per resident sub-warp of feat/4 lanes), rows straddling a range boundary
 * become pieces that are reduced through `scratch` by a ticket counter.  Built once per edge set by
 * the host (gnndelete_b200/graph.py::BatchPlan); all arrays are device pointers owned by the caller.
 *   desc[b]  : bit 31 = batch ends its row / piece (flush); bit 30 = the flush is a piece;
 *              low 30 bits = row id, or piece id when bit 30 is set;
 *   colp     : [num_batches][8] source ids, -1 = padding (only at the end of a row's last batch).
 * desc, colp (and the caller's valp) must be readable for TWO batches past num_batches (the kernel loads the
 * next batch unconditionally); the values there are ignored. */
typedef struct gd_spmm_bplan {
    int64_t num_rows;
    int64_t num_batches;
    int32_t num_workers;
    int32_t batches_per_worker;
    const int32_t* desc;             /* [num_batches + 2] */
    const int32_t* colp;             /* [(num_batches + 2) * 8], 16-byte aligned */
    int32_t num_split;               /* rows cut into pieces */
    int32_t num_piece;
    const int32_t* piece_split;      /* [num_piece] index of the piece's row in split_* */
    const int32_t* split_row;        /* [num_split] */
    const int32_t* split_piece_beg;  /* [num_split] first piece of the row (pieces of a row are consecutive) */
    const int32_t* split_npiece;     /* [num_split] */
    int32_t* split_ticket;           /* [num_split] zero-initialised; self re-arming */
} gd_spmm_bplan_t;

/* Native plan builder (gnndelete_b200/csrc/bplan_build.cu).  gd_spmm_bplan_count writes {num_batches, num_split,
 * num_piece} to `sizes` (device int32[3]) and leaves its per-row scans in `workspace`; the caller reads the sizes,
 * allocates desc [num_batches + 2], colp [(num_batches + 2) * 8], slot_of_entry [nnz] (padded slot of every CSR
 * entry), piece_split [num_piece], split_* [num_split] and calls gd_spmm_bplan_fill with the SAME workspace and
 * batches_per_worker = ceil(num_batches / min(num_workers, num_batches)). */
size_t gd_spmm_bplan_workspace_bytes(int64_t num_rows);
int gd_spmm_bplan_count(const int32_t* rowptr, int64_t num_rows, int32_t num_workers, int32_t* sizes,
                        void* workspace, size_t workspace_bytes, gd_stream_t stream);
int gd_spmm_bplan_fill(const int32_t* rowptr, const int32_t* col, int64_t num_rows, int32_t num_batches,
                       int32_t batches_per_worker, int32_t* desc, int32_t* colp, int64_t* slot_of_entry,
                       int32_t* piece_split, int32_t* split_row, int32_t* split_piece_beg, int32_t* split_npiece,
                       const void* workspace, gd_stream_t stream);

/* Number of workers (sub-warps of feat/4 lanes) the device keeps resident for the batched kernel:
 * the plan is balanced for exactly this many.  feat in {32, 64, 128}; 0 otherwise. */
int32_t gd_spmm_batched_workers(int32_t feat, int32_t weighted);

/* out[i,:] = row_scale[i] * ( sum_{slots s of row i} valp[s] * x[colp[s],:] ) + self_coef * x[i,:] + bias[:]
 * (out += ... when `accumulate`).  `valp` (nullable = unit weights) is in the PADDED slot layout of the
 * plan ([num_batches * 8], padding slots ignored).  `scratch` holds num_piece * feat floats.
 * Same call sites as gd_spmm: GCNConv/GINConv propagate (gcn.py:11-12, gin.py:11-12), the transpose
 * backward and the loss backward gather. */
int gd_spmm_batched(const gd_spmm_bplan_t* plan, const float* valp, const float* row_scale, const float* x,
                    int64_t ldx, int32_t feat, float self_coef, const float* bias, float* out, int64_t ldo,
                    float* scratch, int32_t accumulate, gd_stream_t stream);

/* gd_spmm_batched plus a second, plain-CSR operand whose (few) entries are added to every row when it is
 * flushed:  out[i,:] += sum_{k in tail row i} tail_val[k] * x[tail_col[k],:]  (inside the row scale).
 * The loss gradient uses it for this step's negative pairs, whose incidence is rebuilt every epoch
 * (framework/trainer/gnndelete.py:221-225) and therefore has no batch plan.  tail_* nullable. */
int gd_spmm_batched_tail(const gd_spmm_bplan_t* plan, const float* valp, const int32_t* tail_rowptr,
                         const int32_t* tail_col, const float* tail_val, const float* row_scale, const float* x,
                         int64_t ldx, int32_t feat, float self_coef, const float* bias, float* out, int64_t ldo,
                         float* scratch, int32_t accumulate, gd_stream_t stream);

/* gd_spmm_batched_tail with a bf16 SOURCE matrix x (fp32 accumulation and output) - the "bf16-gather" mode
 * (gnndelete_b200/csrc/spmm_batched_bf16.cu): halves the bytes every non-zero gathers; it is the wire format of the
 * row-partitioned epoch's halo blocks.  ldx in bf16 elements (multiple of 8); feat in {64, 128}; the plan must be
 * balanced for gd_spmm_batched_bf16_workers(feat, weighted) workers.  Results differ from the fp32 kernel by the
 * bf16 rounding of the source rows (tolerance 2e-2, stated in the tests). */
int32_t gd_spmm_batched_bf16_workers(int32_t feat, int32_t weighted);
int gd_spmm_batched_bf16(const gd_spmm_bplan_t* plan, const float* valp, const int32_t* tail_rowptr,
                         const int32_t* tail_col, const float* tail_val, const float* row_scale,
                         const void* x_bf16, int64_t ldx, int32_t feat, float self_coef, const float* bias,
                         float* out, int64_t ldo, float* scratch, int32_t accumulate, gd_stream_t stream);

/* out_bf16[r, :feat] = bf16(row_scale[r] * x[r, :feat]), round to nearest even; row_scale nullable; feat multiple of
 * 8, ldo multiple of 8 elements. */
int gd_cast_bf16(const float* x, int64_t ldx, int64_t rows, int32_t feat, const float* row_scale, void* out_bf16,
                 int64_t ldo, gd_stream_t stream);

/* dst[dst_idx[i]] = src[src_idx[i]] for i < n (either index list nullable = identity).  The row-partitioned epoch
 * uses it to drop the all-gathered DEC coefficients into the slots of its incidence plan. */
int gd_move_f32(const float* src, const int32_t* src_idx, float* dst, const int32_t* dst_idx, int64_t n,
                gd_stream_t stream);

/* GATConv(heads=1) edge-softmax aggregation (gat.py:11-12; defaults negative_slope=0.2,
 * add_self_loops=True — the CSR must be built with self_loops=1):
 *   a_src[k] = <h_k, att_src>, a_dst[i] = <h_i, att_dst>                 (gd_gat_scores)
 *   e_ik = LeakyReLU(a_src[k] + a_dst[i]);  alpha = softmax over the row (+1e-16 in the
 *   denominator);  out_i = sum_k alpha_ik h_k + bias                     (gd_gat_fwd)
 * rowmax / rowden (row max and softmax denominator) are saved for the backward.
 * Backward w.r.t. h (the attention parameters are frozen on the Del path):
 *   gd_gat_bwd_dst: per edge alpha and d e_ik * LeakyReLU', written at the entry's position
 *                   in the transposed CSR (`tinv[fwd position] = transposed position`),
 *                   and d a_dst[i];
 *   gd_gat_bwd_src: dh_k = sum_i alpha_ik g_i + d a_src[k] att_src + d a_dst[k] att_dst.
 * out_channels must be 32, 64 or 128.
 * Rows longer than the CSR's split length (gd_spmm_plan_build) are processed as segments by separate sub-warps and
 * merged in order (online softmax for the forward) through `scratch`: gd_gat_scratch_floats(csr, channels) floats,
 * nullable when the CSR carries no split plan (whole rows per sub-warp then). */
int gd_gat_scores(const float* h, int64_t ldh, int64_t num_nodes, int32_t channels,
                  const float* att_src, const float* att_dst, float* a_src, float* a_dst,
                  gd_stream_t stream);
size_t gd_gat_scratch_floats(const gd_csr_t* csr, int32_t channels);
int gd_gat_fwd(const gd_csr_t* csr, const float* h, int64_t ldh, int32_t channels, const float* a_src,
               const float* a_dst, const float* bias, float negative_slope, float* out, int64_t ldo,
               float* rowmax, float* rowden, float* scratch, gd_stream_t stream);
int gd_gat_bwd_dst(const gd_csr_t* csr, const int32_t* tinv, const float* h, int64_t ldh,
                   int32_t channels, const float* a_src, const float* a_dst, const float* rowmax,
                   const float* rowden, const float* gout, int64_t ldg, const float* out, int64_t ldo,
                   const float* bias, float negative_slope, float* alpha_t, float* dpre_t,
                   float* da_dst, float* scratch, gd_stream_t stream);
int gd_gat_bwd_src(const gd_csr_t* csr_t, const float* alpha_t, const float* dpre_t, const float* gout,
                   int64_t ldg, int32_t channels, const float* att_src, const float* att_dst,
                   const float* da_dst, float* dh, int64_t lddh, float* da_src, float* scratch,
                   gd_stream_t stream);

/* RGCNConv(in, out, R, num_blocks=B|None), aggr='mean', root_weight, bias (rgcn.py:17-22):
 *   out_i = sum_r mean_{k in N_r(i)} x_k . W_r + x_i . root + bias
 * on a CSR built with `rel` as secondary key (entries of a row sorted by relation).
 *   gd_rgcn_norm : entry_weight[k] = 1 / |N_rel(k)(dst(k))|  (the per-(dst, relation) mean);
 *   gd_rgcn_conv : one kernel for all relations (replaces PyG's Python loop over R masked
 *                  propagate calls).  weight is [R, B, in/B, out/B] (B = 1: dense
 *                  [R, in, out]), root [in, out].  transposed = 1 computes the gradient
 *                  w.r.t. x: pass the TRANSPOSED CSR, its relation array, the forward entry
 *                  weights permuted to transposed order (gd_permute_f32 with tinv) and
 *                  x = d out; W_r^T / root^T are applied, bias is ignored.
 * in / out must be in {32, 64, 128}. */
int gd_rgcn_norm(const int32_t* rowptr, const int32_t* rel, int64_t num_rows, float* entry_weight,
                 gd_stream_t stream);
int gd_rgcn_conv(const gd_csr_t* csr, const int32_t* rel, const float* entry_weight, const float* x,
                 int64_t ldx, const float* weight, const float* root, const float* bias,
                 int32_t num_rel, int32_t num_blocks, int32_t in_dim, int32_t out_dim,
                 int32_t transposed, float* out, int64_t ldo, gd_stream_t stream);
/* RGCNConv with 4 block-diagonal relation weights as ONE pass over relation-sorted edge tiles
 * (gnndelete_b200/csrc/rgcn_edge.cu; rgcn.py:17-22 with num_blocks = 4): no [R N, out] intermediate, no dense
 * expansion of the blocks.  The host plan (gnndelete_b200/ops.py::_RgcnEdgePlan) sorts the entries of every tile of
 * `tile_rows` destination rows by (relation, destination) and cuts them into work items [item_beg, item_end):
 *   ent_src[e]  source row,  ent_meta[e] = relation << 5 | destination row inside its tile,
 *   ent_w[e]    the mean weight 1 / |N_r(i)| of the entry (gd_rgcn_norm).
 * gd_rgcn_edge_conv writes scratch[item, row in tile, :] = sum_e ent_w[e] * (x[ent_src[e], block] . W_rel[block]) for
 * the edges of the item; gd_rgcn_edge_reduce adds the items of every tile, in order, ONTO out (which the caller has
 * filled with x . root + bias).  weight: [R, 4, in_block, out_block] of the direction computed (pass W^T blocks and
 * the transposed edge list for the gradient w.r.t. x).  (in_block, out_block) in {(32,16), (16,32), (32,32)};
 * gd_rgcn_edge_tile_rows returns the tile height the kernel uses for a shape (0: not covered). */
int32_t gd_rgcn_edge_tile_rows(int32_t in_block, int32_t out_block);
int gd_rgcn_edge_conv(const int32_t* item_tile, const int32_t* item_beg, const int32_t* item_end, int64_t num_items,
                      const int32_t* ent_src, const int32_t* ent_meta, const float* ent_w, const float* x,
                      int64_t ldx, const float* weight, int32_t in_block, int32_t out_block, float* scratch,
                      gd_stream_t stream);
int gd_rgcn_edge_reduce(const int32_t* tile_item_ptr, int64_t num_rows, int32_t tile_rows, int32_t out_dim,
                        const float* scratch, float* out, int64_t ldo, gd_stream_t stream);
/* dst[perm[i]] = src[i] */
int gd_permute_f32(const float* src, const int32_t* perm, int64_t n, float* dst, gd_stream_t stream);
/* dst[i,:] = src[idx[i],:]  — nn.Embedding lookup `node_emb(x)` (rgcn.py:30); `status`
 * (device int32) counts out-of-range indices. */
int gd_gather_rows(const float* src, int64_t lds, int64_t src_rows, const int64_t* idx, int64_t m,
                   int32_t feat, float* dst, int64_t ldd, int32_t* status, gd_stream_t stream);

/* ------------------------------------------------------ dense contractions (2)
 * out[r(i),:] = epi( pro(a[r(i),:]) . B ),  i in [0,m),  r(i) = rows ? rows[i] : i
 *   pro : optional ReLU;           B : b_is_nk ? B[n,k] (nn.Linear weight, x @ B^T)
 *                                                : B[k,n] (deletion_weight / root, x @ B)
 *   epi : + bias[n], * out_scale[r(i)], optional ReLU, optional gate (kept where
 *         gate[r(i), c] > 0 else 0 — the ReLU backward of the producing layer).
 * Replaces nn.Linear inside GCNConv/GATConv/GINConv and
 * DeletionLayer.forward's `new_rep[mask] = new_rep[mask] @ deletion_weight`
 * (framework/models/deletion.py:24-25) when `rows` lists the masked rows. */
int gd_gemm_rows(const float* a, int64_t lda, const int32_t* rows, int64_t m, int32_t k,
                 const float* b, int32_t b_is_nk, int32_t n, const float* bias,
                 const float* out_scale, const float* gate, int64_t ldgate, int32_t relu_in,
                 int32_t relu_out, float* out, int64_t ldo, gd_stream_t stream);

/* Same contract as gd_gemm_rows on the tensor cores: tcgen05.mma (kind::tf32) with TMEM
 * accumulators and a 3xTF32 operand split, so results match the fp32 path to ~1e-6 relative.
 * Supported when gd_gemm_rows_tc_supported() returns 1 (k <= ~128, n in {32, 64, 96, 128},
 * 16-byte aligned operands); one persistent CTA per SM.  For k a multiple of 32 (no fp32 gate, no bias combined with a row
 * scale) the product is computed transposed with B (hi and lo) resident in TENSOR MEMORY as the A operand and whole 64-row
 * tiles of `a` streaming through shared memory (csrc/gemm_tc_wt.cu); otherwise B stays resident in shared memory
 * (csrc/gemm_tc.cu; GD_GEMM_ROWS=ring forces it).
 * Two optional bit-packed ReLU helpers ([row][n/32] words, bit c%32 of word c/32):
 *   relu_mask_out : written with (out[r,c] > 0) — the forward records ReLU's mask for free;
 *   gate_bits     : out[r,c] is kept where the bit is set, else 0 — the backward's ReLU'
 *                   gate at 1/32 of the traffic of reading the fp32 activations again. */
int gd_gemm_rows_tc_supported(int32_t k, int32_t n, int64_t lda, int64_t ldo);
int gd_gemm_rows_tc(const float* a, int64_t lda, const int32_t* rows, int64_t m, int32_t k,
                    const float* b, int32_t b_is_nk, int32_t n, const float* bias,
                    const float* out_scale, const float* gate, int64_t ldgate, int32_t relu_in,
                    int32_t relu_out, float* out, int64_t ldo, uint32_t* relu_mask_out,
                    const uint32_t* gate_bits, gd_stream_t stream);
/* `batch` products of the SAME rows with different B matrices, out_i = a . B_i (B_i = b + i * b_stride,
 * out_i = out + i * out_stride): the per-relation transforms Y_r = x . W_r of RGCNConv (rgcn.py:17-22)
 * issued from one call.  No row list, bias, scale or gate. */
int gd_gemm_rows_tc_batch(const float* a, int64_t lda, int64_t m, int32_t k, const float* b, int64_t b_stride,
                          int32_t b_is_nk, int32_t n, float* out, int64_t out_stride, int64_t ldo, int32_t batch,
                          gd_stream_t stream);

/* c[k1,n2] = sum_i a_scale[r(i)] * a[r(i),:k1]^T (outer) g[r(i),:n2] — weight gradient
 * of the contraction above over the (gathered) rows; relu_a applies ReLU to the `a`
 * rows first, a_scale is nullable.  Deterministic two-stage reduction through
 * `workspace`. */
size_t gd_gemm_tn_workspace_bytes(int64_t m, int32_t k1, int32_t n2);
int gd_gemm_tn_rows(const float* a, int64_t lda, const float* g, int64_t ldg, const int32_t* rows,
                    int64_t m, int32_t k1, int32_t n2, int32_t relu_a, const float* a_scale, float* c,
                    void* workspace, size_t workspace_bytes, gd_stream_t stream);

/* gd_gemm_tn_rows on the tensor cores (tcgen05, 3xTF32).  Default (csrc/gemm_tn_wt.cu): the rows of `a` go through a raw
 * shared tile into TENSOR MEMORY as the A operand [k1 lanes x 64 rows], only `g` is staged in operand form, partial sums
 * are added group after group with vector reductions by one thread per element (deterministic).  GD_GEMM_TN=ring
 * (csrc/gemm_tc.cu): rows are transposed while they are
 * staged, every CTA accumulates a contiguous range of rows in TMEM and flushes to its own
 * partial every 256 rows, a second kernel adds the partials in order (deterministic). */
int gd_gemm_tn_rows_tc_supported(int32_t k1, int32_t n2, int64_t lda, int64_t ldg);
size_t gd_gemm_tn_tc_workspace_bytes(int32_t k1, int32_t n2);
int gd_gemm_tn_rows_tc(const float* a, int64_t lda, const float* g, int64_t ldg, const int32_t* rows,
                       int64_t m, int32_t k1, int32_t n2, int32_t relu_a, const float* a_scale, float* c,
                       void* workspace, size_t workspace_bytes, gd_stream_t stream);

/* Input gradient of a gathered-row linear map chained into the weight gradient of the DeletionLayer in front of it
 * (csrc/gemm_dxdw_wt.cu; autograd of framework/models/deletion.py:23-33 `torch.matmul(x[mask], deletion_weight)` under
 * the ReLU and GCNConv.lin of framework/models/gcn.py:15-19):
 *   dX[r, :n] = gate[r, :] (.) ((in_scale[r] x[r, :k]) . B),    c[k1, n] = sum_r a[r, :k1]^T (x) dX[r, :]
 * over r in `rows` (or 0..m).  dX stays in tensor memory (the transposed accumulator of the first product is the A operand
 * of the second); same results as gd_gemm_rows_tc + gd_gemm_tn_rows_tc.  k in {32, 64}; n, k1 multiples of 32, <= 128.
 * gate_bits: [row][n / 32] bit masks (optional).  workspace: gd_gemm_tn_tc_workspace_bytes(k1, n). */
int gd_gemm_dxdw_tc_supported(int32_t k, int32_t n, int32_t k1, int64_t ldx, int64_t lda);
int gd_gemm_dxdw_tc(const float* x, int64_t ldx, const float* b, int32_t b_is_nk, int32_t k, int32_t n,
                    const float* in_scale, const uint32_t* gate_bits, const float* a, int64_t lda, int32_t k1,
                    const int32_t* rows, int64_t m, float* c, void* workspace, size_t workspace_bytes, gd_stream_t stream);

/* dst[rows[i], :] = src[rows[i], :]  (the unmasked rows of DeletionLayer.forward's clone). */
int gd_copy_rows(const float* src, int64_t lds, const int32_t* rows, int64_t m, int32_t feat,
                 float* dst, int64_t ldd, gd_stream_t stream);

/* dst[rows[i], :] = row_scale[rows[i]] * src[rows[i], :]  (row_scale may be NULL).  Backward of the GCN epoch: the
 * D^-1/2 factor of the transpose aggregation's source rows is applied where dA2 is produced (here for the unmasked
 * rows, in the Del GEMM's epilogue for the masked ones), so the aggregation itself runs without per-entry weights. */
int gd_copy_rows_scaled(const float* src, int64_t lds, const int32_t* rows, int64_t m, int32_t feat,
                        const float* row_scale, float* dst, int64_t ldd, gd_stream_t stream);

/* out = max(x, 0)  (F.relu between the two convs, deletion.py:67; used where the ReLU
 * cannot be folded into the next contraction's prologue: GIN / RGCN aggregate first) */
int gd_relu_fwd(const float* x, int64_t count, float* out, gd_stream_t stream);

/* out = grad * (pre > 0)   (ReLU backward, elementwise over [rows, feat]) */
int gd_relu_bwd(const float* grad, const float* pre, int64_t count, float* out, gd_stream_t stream);

/* ------------------------------------------------ decoder + DEC / NI losses (3)
 * Pairs [0,n_df) are the Df edges, [n_df, 2 n_df) the supplied negatives and
 * [2 n_df, 2 n_df + n_ni) the S_Df edges with u < v whose original logits are
 * `target`.  logits[p] = <z[u_p], z[v_p]>          (GCN.decode, gcn.py:26-35)
 *   loss_r = mean_i (logits[i] - logits[n_df+i])^2  (gnndelete.py:227-228, :362-363)
 *   loss_l = mean_j (logits[2n_df+j] - target[j])^2 (gnndelete.py:379-386)
 *   loss   = alpha * loss_r + (1-alpha) * loss_l    (gnndelete.py:249-250, :390-398)
 * losses[3] = {loss, loss_r, loss_l}.  d loss / d logits[p] is scattered to
 * inc_val[pos_u[p]] and inc_val[pos_v[p]], the two incidence entries of pair p, so
 * that dz = gd_spmm(incidence CSR, val = inc_val, x = z) is a deterministic gather. */
size_t gd_edge_loss_workspace_bytes(int64_t num_pairs);
int gd_edge_loss_fwd(const float* z, int64_t ldz, int32_t dim, const int32_t* pair_u,
                     const int32_t* pair_v, int64_t n_df, int64_t n_ni, const float* target,
                     float alpha, const int32_t* pos_u, const int32_t* pos_v, float* logits,
                     float* inc_val, float* losses, void* workspace, size_t workspace_bytes,
                     gd_stream_t stream);

/* Node-side fusion of the same losses with their gradient (gnndelete_b200/csrc/node_loss.cu): ONE pass over the
 * node -> incident-pair lists produces dz and loss_l.  `plan` is a batch plan (gd_spmm_bplan_t) over the incidence
 * rows: NI pairs listed from both endpoints, DEC pairs (Df / negatives) from both endpoints; colp = partner node.
 *   bmeta[b] : row of batch b in the low 24 bits (num_rows < 2^24), bit 24 + s set when slot s of the batch is an
 *              NI entry; readable for two batches past num_batches like desc / colp;
 *   valp[s]  : NI entry: the pair's target logit; DEC entry: d loss / d logit of the pair, as written by
 *              gd_edge_loss_fwd (n_ni = 0) through pos_u / pos_v; padding slots must hold 0;
 *   tail_*   : optional plain CSR of further given-coefficient entries (this step's negative pairs).
 * For plan row r:  dz[r,:] = sum_NI c_l (<zself[r'], z[x]> - target) z[x,:] + sum_DEC valp z[x,:] with
 * r' = row_slot[r] (or r), c_l = (1 - alpha) * 2 / norm_ni.  losses[3] = {alpha loss_r + (1 - alpha) loss_l, loss_r,
 * loss_l} with loss_r = dec_losses[1] (nullable = 0) and loss_l = (sum over NI entries of residual^2) / (2 norm_ni)
 * (every NI pair is listed twice); ni_sq_sum (nullable) receives the raw sum for callers that reduce it over ranks.
 * `bf16` != 0: z / zself hold bf16 rows (ld in elements; the wire format of the row-partitioned epoch), feat in
 * {64, 128}; otherwise fp32 rows, feat in {32, 64, 128}.  No float atomics: bitwise reproducible. */
int32_t gd_node_loss_workers(int32_t feat, int32_t bf16);
size_t gd_node_loss_workspace_bytes(int32_t num_workers);
int gd_node_loss_fwd_bwd(const gd_spmm_bplan_t* plan, const int32_t* bmeta, const float* valp,
                         const int32_t* tail_rowptr, const int32_t* tail_col, const float* tail_val,
                         const void* z, int64_t ldz, const void* zself, int64_t ldself, int32_t bf16,
                         const int32_t* row_slot, int32_t feat, int64_t norm_ni, float alpha,
                         const float* dec_losses, float* dz, int64_t ldo, float* scratch, float* losses,
                         float* ni_sq_sum, void* workspace, size_t workspace_bytes, gd_stream_t stream);

/* dz[u_i,:] += coef[i] * z[v_i,:] and dz[v_i,:] += coef[i] * z[u_i,:] for num_pairs given-coefficient pairs: the
 * gradient of this step's NEGATIVE pairs (resampled every epoch, gnndelete.py:221-225), added onto the dz of
 * gd_node_loss_fwd_bwd with vector float reductions - no per-step sort / incidence rebuild.  The order in which a
 * row's few contributions are added is not fixed (as in the reference's index_add scatter). */
int gd_pair_scatter_add(const float* z, int64_t ldz, int32_t feat, const int32_t* pair_u, const int32_t* pair_v,
                        const float* coef, int64_t num_pairs, float* dz, int64_t ldo, gd_stream_t stream);

/* DEC residuals of a SHARE of the Df items (row-partitioned epoch, gnndelete_b200/dist.py): item i = Df pair
 * (pair_u[i], pair_v[i]) with its negative (pair_u[n_items + i], pair_v[n_items + i]);
 *   r_i = <z_u, z_v> - <z_nu, z_nv>  (gcn.py:26-35, gnndelete.py:227-228);  coef_pos[i] = alpha * 2 / norm_df * r_i,
 *   coef_neg[i] = -coef_pos[i]  (d loss / d logit of the two pairs);  loss_r_part[0] = sum_i r_i^2 / norm_df (this
 *   caller's share of loss_r; the shares add up over ranks);  logits (nullable) [2 n_items].
 * z rows fp32 or bf16 (`bf16`), ldz in elements. */
size_t gd_dec_items_workspace_bytes(void);
int gd_dec_items_fwd(const void* z, int64_t ldz, int32_t bf16, int32_t feat, const int32_t* pair_u,
                     const int32_t* pair_v, int64_t n_items, int64_t norm_df, float alpha, float* coef_pos,
                     float* coef_neg, float* logits, float* loss_r_part, void* workspace, size_t workspace_bytes,
                     gd_stream_t stream);

/* Row-partitioned variant: this caller holds a SUBSET of the Df items / NI pairs (those touching
 * its rows).  Only the first own_df items / own_ni pairs contribute their squared residual to
 * `losses` (every pair is counted by exactly one rank) and the means use the GLOBAL counts
 * norm_df / norm_ni, so summing `losses` over ranks gives the global loss. */
int gd_edge_loss_fwd_part(const float* z, int64_t ldz, int32_t dim, const int32_t* pair_u,
                          const int32_t* pair_v, int64_t n_df, int64_t n_ni, const float* target,
                          float alpha, const int32_t* pos_u, const int32_t* pos_v, float* logits,
                          float* inc_val, float* losses, int64_t own_df, int64_t own_ni,
                          int64_t norm_df, int64_t norm_ni, void* workspace, size_t workspace_bytes,
                          gd_stream_t stream);

/* Dense-block Neighbourhood-Influence loss of train_fullbatch (gnndelete.py:163-193, 239-241):
 *   sum over pairs i > j inside the S_Df node set S (minus the excluded Df pairs) of
 *   (sigmoid(<z_i, z_j>) - sigmoid(logits_ori[i, j]))^2
 * on the S x S block only, tile by tile, nothing N x N is materialised.
 *   zs        [n_s, 64]   rows of z for the nodes of S (ascending node id)
 *   tgt_sig   [n_s, n_s]  sigmoid(logits_ori[S][:, S])
 *   excl_bits n_s*n_s bits (row-major), 1 = pair excluded (set for both orders of a Df pair)
 *   dzs       [n_s, 64]   = coef_scale * d(sum of squared residuals)/d zs   (coef_scale = weight / |M|)
 *   loss_sum  device float, the un-normalised sum.  Deterministic (no atomics). */
size_t gd_dense_ni_workspace_bytes(int64_t n_s);
int gd_dense_ni_fwd_bwd(const float* zs, int64_t ldz, int32_t dim, int64_t n_s, const float* tgt_sig,
                        int64_t ldt, const uint32_t* excl_bits, float coef_scale, float* dzs,
                        int64_t lddz, float* loss_sum, void* workspace, size_t workspace_bytes,
                        gd_stream_t stream);
/* The same loss on the tensor cores (tcgen05.mma kind::tf32 with the 3xTF32 operand split, accumulators in TMEM): one CTA
 * per 128 rows of z_S sweeps the 64-row column blocks - logit tile, sigmoid / residual / coefficient epilogue out of
 * TMEM, gradient contraction accumulated in TMEM; nothing N x N is formed.  The target block is packed once
 * (gd_dense_ni_tc_pack_target: tile-major, coalesced for the epilogue's row-per-thread reads; excluded pairs, the
 * diagonal and the padding carry the sentinel -1, so the kernel reads no bitmap).  dim must be 64
 * (gd_dense_ni_tc_supported); same outputs and the same coef_scale as gd_dense_ni_fwd_bwd. */
int gd_dense_ni_tc_supported(int32_t dim);
size_t gd_dense_ni_tc_target_bytes(int64_t n_s);
int gd_dense_ni_tc_pack_target(const float* tgt_sig, int64_t ldt, const uint32_t* excl_bits, int64_t n_s,
                               float* packed_target, gd_stream_t stream);
size_t gd_dense_ni_tc_workspace_bytes(int64_t n_s);
int gd_dense_ni_tc_fwd_bwd(const float* zs, int64_t ldz, int64_t n_s, const float* packed_target, float coef_scale,
                           float* dzs, int64_t lddz, float* loss_sum, void* workspace, size_t workspace_bytes,
                           gd_stream_t stream);

/* dst[rows[i], :] += src[i, :]   (rows unique) */
int gd_add_rows(const float* src, int64_t lds, const int32_t* rows, int64_t m, int32_t feat, float* dst,
                int64_t ldd, gd_stream_t stream);

/* Node-embedding MSE terms of the layer-wise Del losses, forward + gradient in one pass
 * (framework/trainer/gnndelete_nodeemb.py:196-212 `GNNDeleteNodeembTrainer.train_fullbatch`,
 * :770-781 `KGGNNDeleteNodeembTrainer.train`):
 *   term 0  loss_r = MSE(cat(z[h], z[t]), cat(z_ref[h'], z_ref[t']))   (Deleted Edge Consistency)
 *   term 1  loss_l = MSE(z[rows], z_ref[rows])                          (Neighbourhood Influence)
 * given as ONE destination-major incidence over the rows of z: `rowptr[n + 1]`, and per entry
 * `code >= 0` -> term 0 against z_ref[code], `code < 0` -> term 1 against z_ref[-1 - code].
 *   S_t       = sum over the entries of term t of |z[r] - z_ref[partner]|^2
 *   losses[1] = w0 * S_0, losses[2] = w1 * S_1      (w_t = 1 / (M_t * dim) for 'mse_mean', 1 for 'mse_sum')
 *   losses[0] = a0 * losses[1] + a1 * losses[2]     (the mixed objective, e.g. a0 = alpha, a1 = 1 - alpha)
 *   dz[r, :]  = d losses[0] / d z[r, :]             (every row written, zeros where a row has no entry;
 *                                                    dz NULL: forward only)
 * Deterministic: no atomics, summation order fixed by the incidence and the launch shape. */
size_t gd_row_mse_workspace_bytes(int64_t n);
int gd_row_mse_fwd_bwd(const float* z, int64_t ldz, const float* z_ref, int64_t ldref, int32_t dim, int64_t n,
                       const int32_t* rowptr, const int32_t* code, float w0, float w1, float a0, float a1,
                       float* dz, int64_t lddz, float* losses, void* workspace, size_t workspace_bytes,
                       gd_stream_t stream);

/* logits[p] = sum_d z[u_p,d] * w[t_p,d] * z[v_p,d]  (w, t nullable => plain dot).
 * GCN.decode (gcn.py:26-35) / RGCN.decode DistMult (rgcn.py:40-47). */
int gd_pair_decode(const float* z, int64_t ldz, int32_t dim, const int32_t* pair_u,
                   const int32_t* pair_v, int64_t num_pairs, const float* rel_weight,
                   const int32_t* pair_rel, float* logits, gd_stream_t stream);

/* Adam step with torch.optim.Adam semantics (delete_gnn.py:229-241: lr 1e-3,
 * betas (0.9, 0.999), eps 1e-8, weight_decay 0, amsgrad off) on a flat parameter.
 * `step` is a device float holding the step count BEFORE this call; it is
 * incremented by the kernel so the call can be replayed inside a CUDA graph. */
int gd_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* step,
                 int64_t count, float lr, float beta1, float beta2, float eps, gd_stream_t stream);

/* ------------------------------------------------------ deletion masks (4)
 * torch_geometric.utils.k_hop_subgraph(seeds, num_hops, edge_index, num_nodes,
 * flow='source_to_target') as used at delete_gnn.py:128-151, seeds = endpoints of the
 * edges selected by `seed_edge_mask` (= edge_index[:, df_mask].flatten().unique()).
 * Each hop adds the SOURCES of edges whose TARGET is in the previous frontier
 * (src = edge_index[0], dst = edge_index[1]); the result is the induced-edge mask of
 * the union of all frontiers and the node mask of those edges' endpoints
 * (delete_gnn.py:141-145).  Frontiers are N-bit bitmaps; one streaming pass over the
 * edge list per hop.  `status` (device int32) counts out-of-range endpoints. */
size_t gd_khop_workspace_bytes(int64_t num_nodes);
int gd_khop_masks(const int64_t* src, const int64_t* dst, int64_t num_edges, int64_t num_nodes,
                  const uint8_t* seed_edge_mask, int32_t num_hops, uint8_t* edge_mask,
                  uint8_t* node_mask, int32_t* status, void* workspace, size_t workspace_bytes,
                  gd_stream_t stream);

/* torch_geometric.utils.to_undirected(edge_index, [a, b]) (reduce='add') as used at
 * delete_gnn.py:175-182: concatenate the flipped list, sort by row * N + col, merge
 * duplicates summing the int attributes.  Outputs hold up to 2 * num_edges entries;
 * `out_count` (device int64) receives the number of distinct entries. */
size_t gd_to_undirected_workspace_bytes(int64_t num_edges);
int gd_to_undirected(const int64_t* src, const int64_t* dst, int64_t num_edges, int64_t num_nodes,
                     const int32_t* attr_a, const int32_t* attr_b, int64_t* out_row,
                     int64_t* out_col, int32_t* out_a, int32_t* out_b, int64_t* out_count,
                     void* workspace, size_t workspace_bytes, gd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GNNDELETE_B200_H */
