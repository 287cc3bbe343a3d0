"""Functional wrappers (raw kernels) and ``torch.autograd.Function``s over the C ABI.

The raw wrappers (``spmm``, ``gemm_rows`` …) write into caller-provided or freshly
allocated CUDA tensors and are what the fused epoch engine (``engine.py``) calls; the
autograd Functions make the drop-in modules differentiable for arbitrary user losses.
Gradients nobody can consume on the Del path (frozen conv weights) are only computed
when the corresponding input ``requires_grad``.
"""
from __future__ import annotations

import os

import torch

from . import _lib as L
from .graph import CSR


# 'tc': tcgen05 tensor-core path where the shape allows, 'simt': exact-fp32 CUDA-core path only
GEMM_BACKEND = os.environ.get('GD_GEMM', 'tc')
# A 128-wide aggregation whose source matrix is about the size of the L2 runs as two 64-wide column passes
# (GD_SPMM_SPLIT128=0: never, =force: at every size).  The 120 MB source of the Collab layer-1 aggregation only just fits
# the 126 MB L2 (50 % hit rate, 410 MB of DRAM reads for 120 MB of compulsory source bytes,
# profiles/r1_spmm_batched_ncu_full.md); each half-pass gathers 256-byte half rows from a 60 MB footprint like the 64-wide
# launches.  Measured 121.3 -> 112.7 us, epoch 1118 -> 1137 epochs/s (profiles/r1_split128_ab.md).  Sources far below the
# L2 size gain nothing from a second launch, sources far above it (config 5) do not become L2 resident by halving.
_SPLIT128_MODE = os.environ.get('GD_SPMM_SPLIT128', '1')
SPLIT128 = _SPLIT128_MODE != '0'
L2_BYTES = 126 << 20
SPLIT128_SOURCE_BYTES = (0, float('inf')) if _SPLIT128_MODE == 'force' else (L2_BYTES // 2, 2 * L2_BYTES)


def _f32(t):
    if t.dtype != torch.float32:
        raise RuntimeError(f'gnndelete_b200 kernels compute in fp32; got {t.dtype}')
    return t.contiguous()


# ------------------------------------------------------------------ raw kernels
def spmm(csr: CSR, x, out=None, val=None, col_scale=None, row_scale=None, self_coef=0.0, bias=None,
         accumulate=False, valp=None, tail=None):
    """``out = row_scale * (A . x) + self_coef * x + bias`` with optional per-entry weights.

    Default kernel: the batched aggregation on the CSR's batch plan (``gd_spmm_batched``).  ``valp`` are
    per-slot weights already in the plan's padded layout (``csr.bplan(f, True).slot_of_entry``); a
    constant ``col_scale`` is folded into cached padded weights once.  Per-entry ``val`` in CSR order
    (or a CSR without a plan-able width) goes through the row-walking kernel ``gd_spmm_acc``.  ``tail`` =
    ``(rowptr, col, val)`` of a second plain CSR whose entries are added row by row (batched path only)."""
    if x.dtype == torch.bfloat16:
        return _spmm_bf16(csr, x, out, val, col_scale, row_scale, self_coef, bias, accumulate, valp, tail)
    x = _f32(x)
    n, f = csr.num_rows, x.shape[1]
    if out is None:
        assert not accumulate, 'accumulate needs an output buffer'
        out = torch.empty(n, f, dtype=torch.float32, device=x.device)
    weighted = valp is not None or col_scale is not None
    bp = csr.bplan(f, weighted) if val is None and x.stride(0) % 4 == 0 and out.stride(0) % 4 == 0 else None
    if bp is not None and SPLIT128 and f == 128 and valp is None and tail is None \
            and SPLIT128_SOURCE_BYTES[0] < x.shape[0] * x.stride(0) * 4 <= SPLIT128_SOURCE_BYTES[1] \
            and csr.bplan(64, weighted) is not None:
        bp64 = csr.bplan(64, weighted)
        v64 = bp64.col_scale_weights(col_scale) if col_scale is not None else None
        for off in (0, 256):                         # byte offset of the column half inside every row
            L.call('gd_spmm_batched_tail', bp64.ref, L.ptr(v64), None, None, None, L.ptr(row_scale), L.ptr(x) + off,
                   x.stride(0), 64, float(self_coef), None if bias is None else L.ptr(bias) + off, L.ptr(out) + off,
                   out.stride(0), L.ptr(bp64.scratch(64)), int(bool(accumulate)), L.stream())
        return out
    if bp is not None:
        if valp is None and col_scale is not None:
            valp = bp.col_scale_weights(col_scale)
        t_rp, t_col, t_val = tail if tail is not None else (None, None, None)
        L.call('gd_spmm_batched_tail', bp.ref, L.ptr(valp), L.ptr(t_rp), L.ptr(t_col), L.ptr(t_val), L.ptr(row_scale), L.ptr(x),
               x.stride(0), f, float(self_coef), L.ptr(bias), L.ptr(out), out.stride(0), L.ptr(bp.scratch(f)),
               int(bool(accumulate)), L.stream())
        return out
    assert valp is None and tail is None, 'padded weights / a tail CSR need a batch plan'
    L.call('gd_spmm_acc', csr.ref, L.ptr(val), L.ptr(col_scale), L.ptr(row_scale), L.ptr(x), x.stride(0), f,
           float(self_coef), L.ptr(bias), L.ptr(out), out.stride(0), L.ptr(csr.scratch(f)), int(bool(accumulate)),
           L.stream())
    return out


def _spmm_bf16(csr, x, out, val, col_scale, row_scale, self_coef, bias, accumulate, valp, tail):
    """The bf16-gather mode of :func:`spmm`: bf16 source rows, fp32 accumulation and output (``gd_spmm_batched_bf16``).
    There is no row-walking fallback for bf16 sources: the width must be 64 or 128."""
    x = x.contiguous() if x.stride(1) != 1 else x
    n, f = csr.num_rows, x.shape[1]
    if val is not None:
        raise RuntimeError('per-entry values in CSR order are not supported with a bf16 source (use valp)')
    weighted = valp is not None or col_scale is not None
    bp = csr.bplan(f, weighted, bf16=True)
    if bp is None:
        raise RuntimeError(f'gd_spmm_batched_bf16 covers widths 64 / 128 on a non-empty static CSR; got width {f}')
    if out is None:
        assert not accumulate, 'accumulate needs an output buffer'
        out = torch.empty(n, f, dtype=torch.float32, device=x.device)
    if valp is None and col_scale is not None:
        valp = bp.col_scale_weights(col_scale)
    t_rp, t_col, t_val = tail if tail is not None else (None, None, None)
    L.call('gd_spmm_batched_bf16', bp.ref, L.ptr(valp), L.ptr(t_rp), L.ptr(t_col), L.ptr(t_val), L.ptr(row_scale),
           x.data_ptr(), x.stride(0), f, float(self_coef), L.ptr(bias), L.ptr(out), out.stride(0), L.ptr(bp.scratch(f)),
           int(bool(accumulate)), L.stream())
    return out


def cast_bf16(x, out=None, row_scale=None):
    """``out = bf16(row_scale[:, None] * x)`` (round to nearest even) - the wire / gather format of the bf16-gather mode."""
    x = _f32(x)
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    L.call('gd_cast_bf16', L.ptr(x), x.stride(0), x.shape[0], x.shape[1], L.ptr(row_scale), out.data_ptr(), out.stride(0),
           L.stream())
    return out


def gemm_rows(a, b, b_is_nk, out=None, rows=None, bias=None, out_scale=None, gate=None,
              relu_in=False, relu_out=False, relu_mask_out=None, gate_bits=None):
    """out[r,:] = epi(pro(a[r,:]) . B) over all rows or the ``rows`` list (in place in ``out``).
    ``relu_mask_out`` / ``gate_bits`` (int32 ``[rows, n/32]`` bit masks) are tensor-core-path only."""
    a = _f32(a)
    b = _f32(b)
    k = a.shape[1]
    n = b.shape[0] if b_is_nk else b.shape[1]
    assert (b.shape[1] if b_is_nk else b.shape[0]) == k, 'inner dimensions differ'
    m = a.shape[0] if rows is None else rows.numel()
    if out is None:
        out = torch.empty(a.shape[0], n, dtype=torch.float32, device=a.device)
    fn = 'gd_gemm_rows'
    if GEMM_BACKEND != 'simt' and L.load().gd_gemm_rows_tc_supported(k, n, a.stride(0), out.stride(0)) \
            and (gate is None or gate.stride(0) % 4 == 0):
        fn = 'gd_gemm_rows_tc'          # tcgen05 + TMEM, 3xTF32 split (fp32-level accuracy)
    args = [L.ptr(a), a.stride(0), L.ptr(rows), m, k, L.ptr(b), int(b_is_nk), n,
            L.ptr(bias), L.ptr(out_scale), L.ptr(gate), gate.stride(0) if gate is not None else 0,
            int(relu_in), int(relu_out), L.ptr(out), out.stride(0)]
    if fn == 'gd_gemm_rows_tc':
        args += [L.ptr(relu_mask_out), L.ptr(gate_bits)]
    elif relu_mask_out is not None or gate_bits is not None:
        raise RuntimeError('bit-packed ReLU masks need the tcgen05 GEMM path (shape not supported)')
    L.call(fn, *args, L.stream())
    return out


def gemm_tc_available(k, n, lda, ldo):
    return GEMM_BACKEND != 'simt' and bool(L.load().gd_gemm_rows_tc_supported(k, n, lda, ldo))


_tn_ws = {}


def gemm_tn_rows(a, g, rows=None, relu_a=False, a_scale=None, out=None):
    """c[k1,n2] = sum_r a_scale[r] * a[r,:]^T (x) g[r,:] over all rows or the ``rows`` list."""
    a = _f32(a)
    g = _f32(g)
    k1, n2 = a.shape[1], g.shape[1]
    m = a.shape[0] if rows is None else rows.numel()
    if out is None:
        out = torch.empty(k1, n2, dtype=torch.float32, device=a.device)
    fn = 'gd_gemm_tn_rows'
    nbytes = L.load().gd_gemm_tn_workspace_bytes(m, k1, n2)
    if GEMM_BACKEND != 'simt' and L.load().gd_gemm_tn_rows_tc_supported(k1, n2, a.stride(0), g.stride(0)):
        fn = 'gd_gemm_tn_rows_tc'
        nbytes = L.load().gd_gemm_tn_tc_workspace_bytes(k1, n2)
    key = (a.device, nbytes)
    ws = _tn_ws.get(key)
    if ws is None:
        ws = torch.empty(max(nbytes, 4), dtype=torch.uint8, device=a.device)
        _tn_ws[key] = ws
    L.call(fn, L.ptr(a), a.stride(0), L.ptr(g), g.stride(0), L.ptr(rows), m, k1, n2,
           int(relu_a), L.ptr(a_scale), L.ptr(out), L.ptr(ws), nbytes, L.stream())
    return out


def gemm_dxdw_supported(x, b, b_is_nk, a):
    """Whether :func:`gemm_dxdw` (the chained input-gradient / weight-gradient kernel) takes these operands."""
    k = x.shape[1]
    n = b.shape[0] if b_is_nk else b.shape[1]
    return (GEMM_BACKEND != 'simt' and x.dtype == torch.float32 and a.dtype == torch.float32
            and bool(L.load().gd_gemm_dxdw_tc_supported(k, n, a.shape[1], x.stride(0), a.stride(0))))


def gemm_dxdw(x, b, b_is_nk, a, rows=None, in_scale=None, gate_bits=None, out=None):
    """``out[k1,n] = sum_r a[r,:]^T (x) (gate[r,:] * ((in_scale[r] x[r,:]) @ B))`` over all rows or the ``rows`` list:
    the DeletionLayer input gradient chained into its weight gradient through tensor memory (the [rows, n] gradient is
    never written).  Same results as ``gemm_rows(..., out_scale, gate_bits)`` followed by ``gemm_tn_rows``."""
    k = x.shape[1]
    n = b.shape[0] if b_is_nk else b.shape[1]
    k1 = a.shape[1]
    m = x.shape[0] if rows is None else rows.numel()
    if out is None:
        out = torch.empty(k1, n, dtype=torch.float32, device=x.device)
    nbytes = L.load().gd_gemm_tn_tc_workspace_bytes(k1, n)
    key = (x.device, nbytes)
    ws = _tn_ws.get(key)
    if ws is None:
        ws = torch.empty(max(nbytes, 4), dtype=torch.uint8, device=x.device)
        _tn_ws[key] = ws
    L.call('gd_gemm_dxdw_tc', L.ptr(x), x.stride(0), L.ptr(_f32(b)), int(b_is_nk), k, n, L.ptr(in_scale), L.ptr(gate_bits),
           L.ptr(a), a.stride(0), k1, L.ptr(rows), m, L.ptr(out), L.ptr(ws), nbytes, L.stream())
    return out


def copy_rows(src, dst, rows, row_scale=None):
    """``dst[rows] = src[rows]`` (``* row_scale[rows]`` when given)."""
    L.call('gd_copy_rows_scaled', L.ptr(src), src.stride(0), L.ptr(rows), rows.numel(), src.shape[1],
           L.ptr(row_scale), L.ptr(dst), dst.stride(0), L.stream())
    return dst


def relu_fwd(x, out=None):
    x = _f32(x)
    if out is None:
        out = torch.empty_like(x)
    L.call('gd_relu_fwd', L.ptr(x), x.numel(), L.ptr(out), L.stream())
    return out


def relu_bwd(grad, pre, out=None):
    grad = _f32(grad)
    if out is None:
        out = torch.empty_like(grad)
    L.call('gd_relu_bwd', L.ptr(grad), L.ptr(pre), grad.numel(), L.ptr(out), L.stream())
    return out


def pair_decode(z, pu, pv, rel_weight=None, pair_rel=None):
    z = _f32(z)
    out = torch.empty(pu.numel(), dtype=torch.float32, device=z.device)
    L.call('gd_pair_decode', L.ptr(z), z.stride(0), z.shape[1], L.ptr(pu, 'i32'), L.ptr(pv, 'i32'),
           pu.numel(), L.ptr(rel_weight), L.ptr(pair_rel), L.ptr(out), L.stream())
    return out


def adam_step(param, grad, exp_avg, exp_avg_sq, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    L.call('gd_adam_step', L.ptr(param), L.ptr(grad), L.ptr(exp_avg), L.ptr(exp_avg_sq), L.ptr(step),
           param.numel(), float(lr), float(beta1), float(beta2), float(eps), L.stream())


# ------------------------------------------------------------- autograd Functions
class SpMMFn(torch.autograd.Function):
    """out = R A C x + self_coef x + bias  (R/C: optional diagonal row/col scales)."""

    @staticmethod
    def forward(ctx, x, bias, plan, col_scale, row_scale, self_coef):
        ctx.plan, ctx.col_scale, ctx.row_scale, ctx.self_coef = plan, col_scale, row_scale, self_coef
        ctx.has_bias = bias is not None
        return spmm(plan.fwd, x, col_scale=col_scale, row_scale=row_scale, self_coef=self_coef,
                    bias=None if bias is None else bias.detach())

    @staticmethod
    def backward(ctx, gout):
        gout = gout.contiguous()
        gx = gb = None
        if ctx.needs_input_grad[0]:
            # (R A C)^T = C A^T R
            gx = spmm(ctx.plan.bwd, gout, col_scale=ctx.row_scale, row_scale=ctx.col_scale,
                      self_coef=ctx.self_coef)
        if ctx.has_bias and ctx.needs_input_grad[1]:
            gb = gout.sum(0)
        return gx, gb, None, None, None, None


class LinearFn(torch.autograd.Function):
    """out = (relu?(x) W^T + b) * out_scale   with W = nn.Linear weight [out, in]."""

    @staticmethod
    def forward(ctx, x, weight, bias, out_scale, relu_in):
        ctx.save_for_backward(x, weight)
        ctx.out_scale, ctx.relu_in, ctx.has_bias = out_scale, relu_in, bias is not None
        return gemm_rows(x, weight.detach(), True, bias=None if bias is None else bias.detach(),
                         out_scale=out_scale, relu_in=relu_in)

    @staticmethod
    def backward(ctx, gout):
        x, weight = ctx.saved_tensors
        gout = gout.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            # gx = (gout * s) W, gated by relu'(x)
            gx = gemm_rows(gout, weight.detach(), False, out_scale=ctx.out_scale,
                           gate=x if ctx.relu_in else None)
        if ctx.needs_input_grad[1]:
            # gw[o, i] = sum_r s_r gout[r, o] relu?(x)[r, i]
            gw = gemm_tn_rows(gout, x, a_scale=ctx.out_scale) if not ctx.relu_in else \
                gemm_tn_rows(x, gout, relu_a=True, a_scale=ctx.out_scale).t().contiguous()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = (gout * ctx.out_scale.view(-1, 1)).sum(0) if ctx.out_scale is not None else gout.sum(0)
        return gx, gw, gb, None, None


class ReLUFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _f32(x)
        ctx.save_for_backward(x)
        return relu_fwd(x)

    @staticmethod
    def backward(ctx, gout):
        (x,) = ctx.saved_tensors
        return relu_bwd(gout.contiguous(), x)


class DeletionFn(torch.autograd.Function):
    """y = x.clone(); y[rows] = x[rows] @ W   (framework/models/deletion.py:24-25) without
    materialising the clone-then-overwrite: masked rows come out of the gathered-row GEMM,
    the complement is copied."""

    @staticmethod
    def forward(ctx, x, weight, rows, comp):
        x = _f32(x)
        # like autograd's matmul, keep W only if the input needs a gradient: the KG trainer
        # steps deletion1's optimizer between the two backward passes of one forward
        # (gnndelete_nodeemb.py:788-796), which must not invalidate this node
        ctx.save_for_backward(x, weight if ctx.needs_input_grad[0] else None)
        ctx.rows, ctx.comp = rows, comp
        out = torch.empty_like(x)
        gemm_rows(x, weight.detach(), False, out=out, rows=rows)
        copy_rows(x, out, comp)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, weight = ctx.saved_tensors
        gout = gout.contiguous()
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(gout)
            gemm_rows(gout, weight.detach(), True, out=gx, rows=ctx.rows)     # gout[rows] @ W^T
            copy_rows(gout, gx, ctx.comp)
        if ctx.needs_input_grad[1]:
            gw = gemm_tn_rows(x, gout, rows=ctx.rows)                           # x[rows]^T gout[rows]
        return gx, gw, None, None


class PairDecodeFn(torch.autograd.Function):
    """logits[p] = sum_d z[u_p,d] w[t_p,d] z[v_p,d]; differentiable w.r.t. z through the
    incidence CSR of the pair list (deterministic gather, no atomics)."""

    @staticmethod
    def forward(ctx, z, pairs):
        ctx.pairs = pairs
        ctx.save_for_backward(z)
        return pair_decode(z, pairs.pu, pairs.pv, pairs.rel_weight, pairs.pair_rel)

    @staticmethod
    def backward(ctx, glogits):
        (z,) = ctx.saved_tensors
        pairs = ctx.pairs
        if pairs.rel_weight is not None:
            raise NotImplementedError('DistMult decode backward is not on the unlearning hot path')
        val = glogits.contiguous()[pairs.inc_pair.long()]
        return spmm(pairs.inc, z, val=val), None


# 'edge' (default): one pass over relation-sorted edge tiles with register-resident weight slices (gd_rgcn_edge_conv,
# block-diagonal weights); 'transform': transform-then-gather through a [R N, out] intermediate (the path for dense
# relation weights); 'tile': the generic one-kernel relation-tile path (gd_rgcn_conv, any shape in {32, 64, 128})
RGCN_MODE = os.environ.get('GD_RGCN', 'edge')
RGCN_EDGE_CHUNK = 512                                    # edges per work item of the edge path
RGCN_Y_LIMIT_BYTES = 32 << 30                            # per-call workspace cap of the transform-then-gather path


def _rgcn_dense_weights(plan, weight):
    """Per-relation dense ``[R, in, out]`` weights (block-diagonal blocks expanded), cached on the plan until the
    parameter changes (the relation weights are frozen on the Del path)."""
    key = (weight.data_ptr(), weight._version, tuple(weight.shape))
    hit = getattr(plan, '_rgcn_dense', None)
    if hit is None or hit[0] != key:
        if weight.dim() == 4:
            r, b, ib, ob = weight.shape
            d = torch.zeros(r, b * ib, b * ob, dtype=torch.float32, device=weight.device)
            for k in range(b):
                d[:, k * ib:(k + 1) * ib, k * ob:(k + 1) * ob] = weight[:, k]
        else:
            d = weight.to(torch.float32)
        plan._rgcn_dense = hit = (key, d.contiguous())
    return hit[1]


class _RgcnEdgePlan:
    """Work items of ``gd_rgcn_edge_conv`` for one direction of a relation-typed CSR: the entries of every tile of
    ``tile_rows`` destination rows sorted by (relation, destination), cut into chunks of ``RGCN_EDGE_CHUNK`` edges."""

    def __init__(self, csr, entry_weight, num_nodes, tile_rows):
        dev = csr.rowptr.device
        n, T = int(num_nodes), int(tile_rows)
        rp = csr.rowptr.long()
        deg = rp[1:] - rp[:-1]
        dst = torch.repeat_interleave(torch.arange(n, device=dev), deg)
        rel = csr.rel.long()
        num_rel = int(rel.max().item()) + 1 if csr.nnz else 1
        if num_rel >= (1 << 26):
            raise ValueError('relation ids do not fit the packed edge word')
        key = torch.div(dst, T, rounding_mode='floor') * (num_rel * T) + rel * T + dst % T
        order = torch.argsort(key, stable=True)
        self.ent_src = csr.col[order].to(torch.int32).contiguous()
        self.ent_meta = ((rel[order] << 5) | (dst[order] % T)).to(torch.int32).contiguous()
        self.ent_w = entry_weight[:csr.nnz][order].contiguous()
        tiles = -(-n // T)
        starts = torch.arange(0, tiles + 1, device=dev) * T
        tile_ptr = rp[starts.clamp(max=n)]
        ne = tile_ptr[1:] - tile_ptr[:-1]
        nchunk = torch.clamp((ne + RGCN_EDGE_CHUNK - 1) // RGCN_EDGE_CHUNK, min=1)
        tip = torch.zeros(tiles + 1, dtype=torch.int64, device=dev)
        torch.cumsum(nchunk, 0, out=tip[1:])
        self.num_items = int(tip[-1].item())
        item_tile = torch.repeat_interleave(torch.arange(tiles, device=dev), nchunk)
        k_in = torch.arange(self.num_items, device=dev) - tip[item_tile]
        beg = tile_ptr[item_tile] + k_in * RGCN_EDGE_CHUNK
        end = torch.minimum(beg + RGCN_EDGE_CHUNK, tile_ptr[item_tile + 1])
        i32 = lambda t: t.to(torch.int32).contiguous()                           # noqa: E731
        self.item_tile, self.item_beg, self.item_end = i32(item_tile), i32(beg), i32(end)
        self.tile_item_ptr = i32(tip)
        self.tile_rows, self.num_rows = T, n
        self._scratch = {}

    def scratch(self, out_dim):
        buf = self._scratch.get(out_dim)
        if buf is None:
            buf = torch.empty(max(self.num_items, 1) * self.tile_rows * out_dim, dtype=torch.float32, device=self.ent_w.device)
            self._scratch[out_dim] = buf
        return buf


def _rgcn_edge_plan(plan, transposed, tile_rows):
    cache = plan.__dict__.setdefault('_rgcn_edge', {})
    key = (bool(transposed), int(tile_rows))
    if key not in cache:
        w_fwd, w_bwd = plan.rgcn_weights
        csr = plan.bwd if transposed else plan.fwd
        cache[key] = _RgcnEdgePlan(csr, w_bwd if transposed else w_fwd, plan.num_nodes, tile_rows)
    return cache[key]


def _rgcn_block_weights(plan, weight, transposed):
    """Block weights in the layout the edge kernel reads: ``[R, B, in_block, out_block]`` of the direction being
    computed (the transposed direction needs ``W_r^T`` blocks), cached until the parameter changes."""
    if not transposed:
        return weight
    cache = plan.__dict__.setdefault('_rgcn_wt', {})
    hit = cache.get(weight.data_ptr())
    # the entry holds a tensor on the same storage (keeps the address from being reused) and the storage's version
    if hit is None or hit[0].shape != weight.shape or hit[1] != weight._version:
        if len(cache) > 8:
            cache.clear()
        cache[weight.data_ptr()] = hit = (weight, weight._version, weight.transpose(2, 3).contiguous())
    return hit[2]


def rgcn_conv(plan, x, weight, root, bias, transposed=False, out=None):
    """RGCNConv forward (``transposed``: gradient w.r.t. the layer input).

    Block-diagonal weights (``num_blocks = 4``, rgcn.py:17-22) take the edge path: ``x . root + bias`` on the tensor
    cores, then ONE pass over relation-sorted edge tiles (``gd_rgcn_edge_conv`` / ``_reduce``) that keeps a relation's
    weight slice in registers - no ``[R N, out]`` intermediate, no dense expansion of the blocks.

    Dense relation weights (``GD_RGCN=transform`` forces it for blocks too), transform-then-gather: ``Y_r = x . W_r`` for every relation on the tensor cores
    (``gd_gemm_rows_tc``, one call per relation into one ``[R N, out]`` buffer), then ONE batched weighted
    aggregation over the virtual source index ``rel * N + col`` with the per-(destination, relation) mean
    weights (``gd_spmm_batched`` accumulating onto ``x . root + bias``).  With degree ~ R per node (BioKG:
    108 entries, 102 relations) pre-aggregating per relation saves nothing, so the per-edge products are
    done as dense GEMMs instead of SIMT tile products.  ``GD_RGCN=tile`` selects the one-kernel path."""
    x = _f32(x)
    weight = weight.detach().contiguous()
    root = root.detach().contiguous()
    if weight.dim() == 4:
        num_rel, blocks, ib, ob = weight.shape
        in_dim, out_dim = blocks * ib, blocks * ob
    else:
        num_rel, in_dim, out_dim = weight.shape
        blocks = 1
    fout = in_dim if transposed else out_dim
    n = x.shape[0]
    b = None if (bias is None or transposed) else bias.detach().contiguous()
    if RGCN_MODE == 'edge' and weight.dim() == 4 and blocks == 4 and n == plan.num_nodes and plan.fwd.rel is not None \
            and x.stride(0) % 4 == 0:
        ibe, obe = (ob, ib) if transposed else (ib, ob)
        tile_rows = L.load().gd_rgcn_edge_tile_rows(int(ibe), int(obe))
        if tile_rows:
            ep = _rgcn_edge_plan(plan, transposed, tile_rows)
            wb = _rgcn_block_weights(plan, weight, transposed)
            out = gemm_rows(x, root, transposed, out=out, bias=b)        # x . root + bias   (g . root^T)
            scratch = ep.scratch(fout)
            L.call('gd_rgcn_edge_conv', L.ptr(ep.item_tile), L.ptr(ep.item_beg), L.ptr(ep.item_end), ep.num_items,
                   L.ptr(ep.ent_src), L.ptr(ep.ent_meta), L.ptr(ep.ent_w), L.ptr(x), x.stride(0), L.ptr(wb), int(ibe), int(obe),
                   L.ptr(scratch), L.stream())
            L.call('gd_rgcn_edge_reduce', L.ptr(ep.tile_item_ptr), n, ep.tile_rows, fout, L.ptr(scratch), L.ptr(out),
                   out.stride(0), L.stream())
            return out
    vcsr = plan.rgcn_virtual(transposed) if RGCN_MODE != 'tile' else None
    bp = vcsr.bplan(fout, True) if vcsr is not None and num_rel * n * fout * 4 <= RGCN_Y_LIMIT_BYTES else None
    if bp is not None and n == plan.num_nodes:
        wd = _rgcn_dense_weights(plan, weight)
        y = torch.empty(num_rel * n, fout, dtype=torch.float32, device=x.device)
        k_in = x.shape[1]
        if gemm_tc_available(k_in, fout, x.stride(0), fout):        # Y_r = x . W_r   (transposed: g . W_r^T), one C call
            L.call('gd_gemm_rows_tc_batch', L.ptr(x), x.stride(0), n, k_in, L.ptr(wd), wd.stride(0), int(transposed), fout,
                   L.ptr(y), n * fout, fout, num_rel, L.stream())
        else:
            for r in range(num_rel):
                gemm_rows(x, wd[r], transposed, out=y[r * n:(r + 1) * n])
        out = gemm_rows(x, root, transposed, out=out, bias=b)        # x . root + bias   (g . root^T)
        valp = plan.rgcn_virtual_weights(transposed, bp)
        L.call('gd_spmm_batched', bp.ref, L.ptr(valp), None, L.ptr(y), y.stride(0), fout, 0.0, None, L.ptr(out),
               out.stride(0), L.ptr(bp.scratch(fout)), 1, L.stream())
        return out
    w_fwd, w_bwd = plan.rgcn_weights
    csr = plan.bwd if transposed else plan.fwd
    if out is None:
        out = torch.empty(n, fout, dtype=torch.float32, device=x.device)
    L.call('gd_rgcn_conv', csr.ref, L.ptr(csr.rel), L.ptr(w_bwd if transposed else w_fwd), L.ptr(x), x.stride(0),
           L.ptr(weight), L.ptr(root), L.ptr(b), num_rel, blocks, in_dim, out_dim, int(transposed), L.ptr(out),
           out.stride(0), L.stream())
    return out


def gather_rows(src, idx):
    """``src[idx]`` (nn.Embedding lookup) with range checking on the device."""
    src = _f32(src)
    idx = idx.contiguous()
    out = torch.empty(idx.numel(), src.shape[1], dtype=torch.float32, device=src.device)
    status = torch.zeros(1, dtype=torch.int32, device=src.device)
    L.call('gd_gather_rows', L.ptr(src), src.stride(0), src.shape[0], L.ptr(idx, 'i64'), idx.numel(), src.shape[1],
           L.ptr(out), out.stride(0), L.ptr(status), L.stream())
    return out, status


class RGCNConvFn(torch.autograd.Function):
    """RGCNConv forward; differentiable w.r.t. the input features only (relation weights,
    root and bias are frozen on the Del path)."""

    @staticmethod
    def forward(ctx, x, weight, root, bias, plan):
        ctx.plan = plan
        ctx.save_for_backward(weight, root)
        return rgcn_conv(plan, x, weight, root, bias)

    @staticmethod
    def backward(ctx, gout):
        if any(ctx.needs_input_grad[1:4]):
            raise NotImplementedError('gradients of the RGCN relation weights are outside the Del hot path')
        if not ctx.needs_input_grad[0]:
            return (None,) * 5
        weight, root = ctx.saved_tensors
        return rgcn_conv(ctx.plan, gout.contiguous(), weight, root, None, transposed=True), None, None, None, None


class GATAggregateFn(torch.autograd.Function):
    """out_i = sum_k softmax_i(LeakyReLU(a_src[k] + a_dst[i])) h_k + bias over the self-looped
    CSR (GATConv heads=1).  On the Del path only ``h`` needs a gradient (attention vectors and bias are
    frozen); when the original model is trained (``Trainer.train_fullbatch``,
    base.py:75-142) the attention-vector gradients ``h^T d a_src`` / ``h^T d a_dst`` and the bias gradient
    come from the per-node score gradients the backward kernels already produce."""

    @staticmethod
    def forward(ctx, h, att_src, att_dst, bias, plan, slope):
        h = _f32(h)
        n, c = h.shape
        dev = h.device
        a_src = torch.empty(n, dtype=torch.float32, device=dev)
        a_dst = torch.empty(n, dtype=torch.float32, device=dev)
        att_src = att_src.detach().reshape(-1).contiguous()
        att_dst = att_dst.detach().reshape(-1).contiguous()
        L.call('gd_gat_scores', L.ptr(h), h.stride(0), n, c, L.ptr(att_src), L.ptr(att_dst), L.ptr(a_src),
               L.ptr(a_dst), L.stream())
        out = torch.empty(n, c, dtype=torch.float32, device=dev)
        rowmax = torch.empty(n, dtype=torch.float32, device=dev)
        rowden = torch.empty(n, dtype=torch.float32, device=dev)
        b = None if bias is None else bias.detach().contiguous()
        L.call('gd_gat_fwd', plan.fwd.ref, L.ptr(h), h.stride(0), c, L.ptr(a_src), L.ptr(a_dst), L.ptr(b),
               float(slope), L.ptr(out), out.stride(0), L.ptr(rowmax), L.ptr(rowden), L.ptr(plan.fwd.gat_scratch(c)), L.stream())
        ctx.plan, ctx.slope = plan, float(slope)
        ctx.save_for_backward(h, att_src, att_dst, a_src, a_dst, rowmax, rowden, out, b if b is not None else h.new_empty(0))
        ctx.has_bias = b is not None
        return out

    @staticmethod
    def backward(ctx, gout):
        h, att_src, att_dst, a_src, a_dst, rowmax, rowden, out, b = ctx.saved_tensors
        if not any(ctx.needs_input_grad[:3]):
            return None, None, None, (gout.sum(0) if ctx.needs_input_grad[3] else None), None, None
        plan = ctx.plan
        gout = gout.contiguous()
        n, c = h.shape
        dev = h.device
        nnz = plan.fwd.nnz
        alpha_t = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)
        dpre_t = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)
        da_dst = torch.empty(n, dtype=torch.float32, device=dev)
        da_src = torch.empty(n, dtype=torch.float32, device=dev)
        dh = torch.empty(n, c, dtype=torch.float32, device=dev)
        L.call('gd_gat_bwd_dst', plan.fwd.ref, L.ptr(plan.tinv), L.ptr(h), h.stride(0), c, L.ptr(a_src),
               L.ptr(a_dst), L.ptr(rowmax), L.ptr(rowden), L.ptr(gout), gout.stride(0), L.ptr(out), out.stride(0),
               L.ptr(b) if ctx.has_bias else None, ctx.slope, L.ptr(alpha_t), L.ptr(dpre_t), L.ptr(da_dst),
               L.ptr(plan.fwd.gat_scratch(c)), L.stream())
        L.call('gd_gat_bwd_src', plan.bwd.ref, L.ptr(alpha_t), L.ptr(dpre_t), L.ptr(gout), gout.stride(0), c,
               L.ptr(att_src), L.ptr(att_dst), L.ptr(da_dst), L.ptr(dh), dh.stride(0), L.ptr(da_src),
               L.ptr(plan.bwd.gat_scratch(c)), L.stream())
        g_src = torch.mv(h.t(), da_src).view(1, 1, -1) if ctx.needs_input_grad[1] else None    # a_src = h att_src
        g_dst = torch.mv(h.t(), da_dst).view(1, 1, -1) if ctx.needs_input_grad[2] else None
        g_bias = gout.sum(0) if ctx.needs_input_grad[3] else None
        return (dh if ctx.needs_input_grad[0] else None), g_src, g_dst, g_bias, None, None
