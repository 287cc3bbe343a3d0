"""Fused full-graph Del-training epoch for ``GCNDelete`` (BASELINE configs 1, 3, 5).

One epoch = forward of the Delete model on the fixed ``sdf``-masked edge set ->
decode on (Df, supplied negatives) -> ``0.5 * MSE(pos, neg) + 0.5 * NI(edge form)`` ->
backward to ``deletion{1,2}.deletion_weight`` -> Adam (reference
``framework/trainer/gnndelete.py:215-259`` with the edge-form NI of ``:379-386``).

Compared with running the drop-in modules under autograd this engine
  * keeps every activation / gradient buffer resident and writes in place (no clone,
    no per-epoch allocation, no host sync — losses stay on the device);
  * hoists the frozen, input-constant layer-1 conv out of the loop (optional; both
    "recompute" and "hoisted" are reported by bench.py);
  * computes only the gradients somebody consumes: dX1 only on the S1 rows, no conv
    weight/bias gradients;
  * can be captured into one CUDA graph (``capture()``) so an epoch is a single launch.
The kernels are the same C-ABI calls the autograd Functions use, so parity of the
engine against the oracle is also parity of the modules.
"""
from __future__ import annotations

import os

import torch

from . import ops
from .graph import plan_for, rows_of
from .losses import DenseNIPlan, EdgeLossPlan


class GCNDeleteEngine:
    conv_kind = 'gcn'

    def __init__(self, model, data, neg_edge_index, z_ori=None, ni_target=None, hoist_layer1=True,
                 lr=1e-3, betas=(0.9, 0.999), eps=1e-8, alpha=0.5, logits_ori=None, static_negatives=False,
                 deterministic=False):
        """``logits_ori`` (dense ``[N, N]``, the original model's ``z z^T`` as saved in
        ``pred_proba.pt``) selects ``train_fullbatch``'s dense-block NI loss instead of the edge form.
        ``static_negatives``: the supplied negatives are fixed for the run, so the loss
        gradient is one gather over one incidence; leave False when negatives are replaced every epoch
        (``set_negatives`` / ``capture(dynamic_negatives=True)``).  ``deterministic``: replaceable negatives go through a
        sorted per-step incidence instead of float reductions (bitwise reproducible epochs, see ``EdgeLossPlan``)."""
        self.model = model
        dev = data.x.device
        self.x = data.x.contiguous()
        n = self.x.shape[0]
        self.n = n
        ei = data.train_pos_edge_index
        self.plan = plan_for(ei[:, data.sdf_mask].contiguous(), n, self.conv_kind)
        self.rows1, self.comp1 = rows_of(data.sdf_node_1hop_mask.to(dev), n)
        self.rows2, self.comp2 = rows_of(data.sdf_node_2hop_mask.to(dev), n)
        sdf = ei[:, data.sdf_mask]
        ni = sdf[:, sdf[0] < sdf[1]]                                   # gnndelete.py:379-381
        hid, out = model.conv1.out_channels, model.conv2.out_channels
        self.dense = None
        if logits_ori is not None:                                     # gnndelete.py:163-193, 239-241
            self.dense = DenseNIPlan(data.sdf_node_2hop_mask, ei[:, data.df_mask], logits_ori, n, out, weight=1.0 - alpha)
            ni = ni[:, :0]
        self.loss = EdgeLossPlan(ei[:, data.df_mask], neg_edge_index, ni, n, z_ori=z_ori,
                                 target=ni_target, alpha=alpha, static_negatives=static_negatives,
                                 deterministic=deterministic)
        self.losses_total = torch.zeros(3, dtype=torch.float32, device=dev)
        self.alpha = float(alpha)
        f32 = dict(dtype=torch.float32, device=dev)
        self.h0 = torch.empty(n, hid, **f32)
        self.a1 = torch.empty(n, hid, **f32)
        self.x1 = torch.empty(n, hid, **f32)
        self.h1 = torch.empty(n, out, **f32)
        self.a2 = torch.empty(n, out, **f32)
        self.z = torch.empty(n, out, **f32)
        self.dz = torch.empty(n, out, **f32)
        self.da2 = torch.empty(n, out, **f32)
        self.dh1 = torch.empty(n, out, **f32)
        # ReLU mask of x1 on the S1 rows as bits (written by the Del1 GEMM, read by the dX1 GEMM)
        self.bitmask = ops.gemm_tc_available(hid, hid, hid, hid) and ops.gemm_tc_available(out, hid, out, hid) \
            and hid % 32 == 0
        self.x1_bits = torch.zeros(n, hid // 32, dtype=torch.int32, device=dev) if self.bitmask else None
        # dX1 chained into dW_del1 through tensor memory (gemm_dxdw_wt.cu); GD_FUSED_DXDW=0 keeps the two kernels
        self.fused_dxdw = self.bitmask and os.environ.get('GD_FUSED_DXDW', '1') != '0' \
            and ops.gemm_dxdw_supported(self.dh1, torch.empty(out, hid, **f32), False, self.a1)
        self.dx1 = None if self.fused_dxdw else torch.zeros(n, hid, **f32)       # only the S1 rows are ever written / read
        self.hoist = bool(hoist_layer1)
        self._layer1_done = False
        self.params = [model.deletion1.deletion_weight, model.deletion2.deletion_weight]
        for p in self.params:
            p.grad = torch.zeros_like(p)
        self.lr, self.betas, self.eps = lr, betas, eps
        self.state = [dict(m=torch.zeros_like(p), v=torch.zeros_like(p), step=torch.zeros(1, **f32))
                      for p in self.params]
        self.graph = None
        self._graph_dynamic_neg = False
        self.launches_per_epoch = None
        # the complement-row copies of the Del layers touch rows the gathered-row GEMM does not: they run on a side
        # stream next to it (fork / join; captured as parallel graph branches)
        self.side = torch.cuda.Stream()

    def _del_rows(self, src, dst, comp, gemm, row_scale=None):
        """``dst[rows] = gemm(src[rows])`` on the current stream, ``dst[comp] = (row_scale *) src[comp]`` concurrently."""
        cur = torch.cuda.current_stream()
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side):
            ops.copy_rows(src, dst, comp, row_scale=row_scale)
        gemm()
        cur.wait_stream(self.side)

    # ------------------------------------------------------------------ forward
    def layer1(self):
        c1, p = self.model.conv1, self.plan
        ops.gemm_rows(self.x, c1.lin.weight.detach(), True, out=self.h0, out_scale=p.dinv)
        ops.spmm(p.fwd, self.h0, out=self.a1, row_scale=p.dinv, bias=c1.bias.detach())
        self._layer1_done = True

    def forward(self):
        m, p = self.model, self.plan
        if not (self.hoist and self._layer1_done):
            self.layer1()
        w1 = m.deletion1.deletion_weight.detach()
        w2 = m.deletion2.deletion_weight.detach()
        self._del_rows(self.a1, self.x1, self.comp1, lambda: ops.gemm_rows(                   # Del1 on S1
            self.a1, w1, False, out=self.x1, rows=self.rows1, relu_mask_out=self.x1_bits))
        ops.gemm_rows(self.x1, m.conv2.lin.weight.detach(), True, out=self.h1, out_scale=p.dinv, relu_in=True)
        ops.spmm(p.fwd, self.h1, out=self.a2, row_scale=p.dinv, bias=m.conv2.bias.detach())
        self._del_rows(self.a2, self.z, self.comp2, lambda: ops.gemm_rows(                    # Del2 on S2
            self.a2, w2, False, out=self.z, rows=self.rows2))
        return self.loss.forward(self.z, dz_out=self.dz)       # node mode: dz comes out of the same pass

    # ----------------------------------------------------------------- backward
    def backward(self):
        m, p = self.model, self.plan
        g1, g2 = self.params[0].grad, self.params[1].grad
        w2 = m.deletion2.deletion_weight.detach()
        self.loss.backward(self.z, out=self.dz)
        if self.dense is not None:
            loss_l = self.dense.forward_backward(self.z, self.dz)
            self.losses_total[1:2].copy_(self.loss.losses[1:2])
            self.losses_total[2:3].copy_(loss_l)
            torch.add(self.loss.losses[0:1], loss_l, alpha=1.0 - self.alpha, out=self.losses_total[0:1])
        ops.gemm_tn_rows(self.a2, self.dz, rows=self.rows2, out=g2)            # dW_del2
        # da2 <- D^-1/2 dA2: the source-side factor of the transpose aggregation is applied where its operand is produced
        # (GEMM epilogue / complement copy), so the aggregation runs without per-entry weights
        self._del_rows(self.dz, self.da2, self.comp2, lambda: ops.gemm_rows(                  # D^-1/2 (dz[S2] @ W2^T)
            self.dz, w2, True, out=self.da2, rows=self.rows2, out_scale=p.dinv), row_scale=p.dinv)
        ops.spmm(p.bwd, self.da2, out=self.dh1)                                    # A^T (D^-1/2 dA2)
        w_lin2 = m.conv2.lin.weight.detach()
        if self.fused_dxdw:
            # dX1 = ReLU' (D^-1/2 dH1) W_2 on S1 chained into dW_del1 = a1[S1]^T dX1 through tensor memory
            ops.gemm_dxdw(self.dh1, w_lin2, False, self.a1, rows=self.rows1, in_scale=p.dinv, gate_bits=self.x1_bits, out=g1)
            return
        ops.gemm_rows(self.dh1, w_lin2, False, out=self.dx1, rows=self.rows1,
                      out_scale=p.dinv, gate=None if self.bitmask else self.x1,
                      gate_bits=self.x1_bits)                                      # ReLU' (D^-1/2 dH1) W_2 on S1
        ops.gemm_tn_rows(self.a1, self.dx1, rows=self.rows1, out=g1)               # dW_del1

    def forward_backward(self):
        losses = self.forward()
        self.backward()
        return self.losses_total if self.dense is not None else losses

    def adam_step(self):
        for p, st in zip(self.params, self.state):
            ops.adam_step(p.data, p.grad, st['m'], st['v'], st['step'], self.lr, self.betas[0], self.betas[1], self.eps)

    def epoch(self):
        """forward + backward + Adam; returns the persistent device tensor (loss, loss_r, loss_l)."""
        if self.graph is not None:
            self.graph.replay()
            return self.losses_total if self.dense is not None else self.loss.losses
        losses = self.forward_backward()
        self.adam_step()
        return losses

    # --------------------------------------------------------------- CUDA graph
    def capture(self, warmup=2, dynamic_negatives=False):
        """Capture one epoch into a CUDA graph.  Warm-up epochs run first (module loading
        and workspace allocation are not capturable) and are undone, so the captured
        graph starts from the current parameters and optimizer state."""
        if dynamic_negatives and self.loss.static:
            raise ValueError('capture(dynamic_negatives=True) needs an engine built with static_negatives=False')
        snap_p = [p.detach().clone() for p in self.params]
        snap_s = [{k: v.clone() for k, v in st.items()} for st in self.state]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.forward_backward()
                self.adam_step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            if dynamic_negatives:
                self.loss.update_negatives()          # rebuild the negative incidence from the staging buffer
            self.forward_backward()
            self.adam_step()
        self._graph_dynamic_neg = bool(dynamic_negatives)
        self._restore(snap_p, snap_s)
        self.graph = g
        return g

    def _restore(self, snap_p, snap_s):
        with torch.no_grad():
            for p, sp in zip(self.params, snap_p):
                p.copy_(sp)
            for st, ss in zip(self.state, snap_s):
                for k in st:
                    st[k].copy_(ss[k])

    def set_negatives(self, neg_edge_index):
        """New supplied negatives (the reference resamples them every epoch, gnndelete.py:221-225).
        Eager mode: the negative incidence is rebuilt in place right away.  After ``capture(
        dynamic_negatives=True)`` the rebuild is part of the graph: only the staging buffer is
        overwritten here (host or device source) and the next ``epoch()`` picks it up."""
        if self.graph is not None and self._graph_dynamic_neg:
            self.loss.neg_buf.copy_(neg_edge_index, non_blocking=True)
        else:
            if self.loss.static:
                self.graph = None                     # the incidence buffers are rebuilt: a captured graph is stale
            self.loss.update_negatives(neg_edge_index)


class GATDeleteEngine(GCNDeleteEngine):
    """The same fused epoch for ``GATDelete`` (deletion.py:80-105 over gat.py:7-24, BASELINE config 2): the two
    aggregations are PyG ``GATConv(heads=1)`` edge-softmax aggregations (``gd_gat_scores`` / ``gd_gat_fwd`` and the two
    backward kernels) instead of normalised sums; everything else - Del GEMMs, fused loss, Adam, CUDA-graph capture,
    negative handling - is inherited.  The attention vectors and conv weights are frozen on the Del path, so only the
    gradient w.r.t. the layer-2 input is computed."""
    conv_kind = 'gat'

    def __init__(self, model, data, neg_edge_index, **kw):
        super().__init__(model, data, neg_edge_index, **kw)
        from . import _lib as L
        self.L = L
        dev, n = self.x.device, self.n
        hid, out = model.conv1.out_channels, model.conv2.out_channels
        f32 = dict(dtype=torch.float32, device=dev)
        nnz = self.plan.fwd.nnz
        self.att = {}
        for name, conv, c in (('c1', model.conv1, hid), ('c2', model.conv2, out)):
            self.att[name] = dict(src=conv.att_src.detach().reshape(-1).contiguous(), dst=conv.att_dst.detach().reshape(-1).contiguous(),
                                  a_src=torch.empty(n, **f32), a_dst=torch.empty(n, **f32), rowmax=torch.empty(n, **f32),
                                  rowden=torch.empty(n, **f32), slope=float(conv.negative_slope))
        self.alpha_t = torch.empty(max(nnz, 1), **f32)
        self.dpre_t = torch.empty(max(nnz, 1), **f32)
        self.da_dst = torch.empty(n, **f32)
        self.da_src = torch.empty(n, **f32)
        self.tinv = self.plan.tinv

    def _gat(self, key, h, bias, out):
        L, a, p = self.L, self.att[key], self.plan
        c = h.shape[1]
        L.call('gd_gat_scores', L.ptr(h), h.stride(0), self.n, c, L.ptr(a['src']), L.ptr(a['dst']), L.ptr(a['a_src']),
               L.ptr(a['a_dst']), L.stream())
        L.call('gd_gat_fwd', p.fwd.ref, L.ptr(h), h.stride(0), c, L.ptr(a['a_src']), L.ptr(a['a_dst']), L.ptr(bias),
               a['slope'], L.ptr(out), out.stride(0), L.ptr(a['rowmax']), L.ptr(a['rowden']), L.ptr(p.fwd.gat_scratch(c)), L.stream())

    def layer1(self):
        c1 = self.model.conv1
        ops.gemm_rows(self.x, c1.lin_src.weight.detach(), True, out=self.h0)
        self._gat('c1', self.h0, c1.bias.detach(), self.a1)
        self._layer1_done = True

    def forward(self):
        m = self.model
        if not (self.hoist and self._layer1_done):
            self.layer1()
        w1 = m.deletion1.deletion_weight.detach()
        w2 = m.deletion2.deletion_weight.detach()
        self._del_rows(self.a1, self.x1, self.comp1, lambda: ops.gemm_rows(                   # Del1 on S1
            self.a1, w1, False, out=self.x1, rows=self.rows1, relu_mask_out=self.x1_bits))
        ops.gemm_rows(self.x1, m.conv2.lin_src.weight.detach(), True, out=self.h1, relu_in=True)
        self._gat('c2', self.h1, m.conv2.bias.detach(), self.a2)
        self._del_rows(self.a2, self.z, self.comp2, lambda: ops.gemm_rows(                    # Del2 on S2
            self.a2, w2, False, out=self.z, rows=self.rows2))
        return self.loss.forward(self.z, dz_out=self.dz)

    def backward(self):
        m, p, L = self.model, self.plan, self.L
        g1, g2 = self.params[0].grad, self.params[1].grad
        w2 = m.deletion2.deletion_weight.detach()
        self.loss.backward(self.z, out=self.dz)
        if self.dense is not None:
            loss_l = self.dense.forward_backward(self.z, self.dz)
            self.losses_total[1:2].copy_(self.loss.losses[1:2])
            self.losses_total[2:3].copy_(loss_l)
            torch.add(self.loss.losses[0:1], loss_l, alpha=1.0 - self.alpha, out=self.losses_total[0:1])
        ops.gemm_tn_rows(self.a2, self.dz, rows=self.rows2, out=g2)                # dW_del2
        self._del_rows(self.dz, self.da2, self.comp2, lambda: ops.gemm_rows(                  # dz[S2] @ W2^T
            self.dz, w2, True, out=self.da2, rows=self.rows2))
        a, c2 = self.att['c2'], m.conv2
        out = self.h1.shape[1]
        bias = c2.bias.detach()
        L.call('gd_gat_bwd_dst', p.fwd.ref, L.ptr(self.tinv), L.ptr(self.h1), self.h1.stride(0), out, L.ptr(a['a_src']),
               L.ptr(a['a_dst']), L.ptr(a['rowmax']), L.ptr(a['rowden']), L.ptr(self.da2), self.da2.stride(0), L.ptr(self.a2),
               self.a2.stride(0), L.ptr(bias), a['slope'], L.ptr(self.alpha_t), L.ptr(self.dpre_t), L.ptr(self.da_dst),
               L.ptr(p.fwd.gat_scratch(out)), L.stream())
        L.call('gd_gat_bwd_src', p.bwd.ref, L.ptr(self.alpha_t), L.ptr(self.dpre_t), L.ptr(self.da2), self.da2.stride(0), out,
               L.ptr(a['src']), L.ptr(a['dst']), L.ptr(self.da_dst), L.ptr(self.dh1), self.dh1.stride(0), L.ptr(self.da_src),
               L.ptr(p.bwd.gat_scratch(out)), L.stream())
        w_src2 = c2.lin_src.weight.detach()
        if self.fused_dxdw:
            ops.gemm_dxdw(self.dh1, w_src2, False, self.a1, rows=self.rows1, gate_bits=self.x1_bits, out=g1)
            return
        ops.gemm_rows(self.dh1, w_src2, False, out=self.dx1, rows=self.rows1,
                      gate=None if self.bitmask else self.x1, gate_bits=self.x1_bits)          # ReLU' (dH1 W_2) on S1
        ops.gemm_tn_rows(self.a1, self.dx1, rows=self.rows1, out=g1)               # dW_del1


class EpochPipeline:
    """Double-buffered host <-> device plumbing around a captured engine (``capture(dynamic_negatives=True)``):
    step k's negatives travel pinned host -> device on a copy stream while step k-1 computes, and step k's
    losses travel device -> pinned host while step k+1 computes; the host only ever waits for the losses of
    the PREVIOUS step.  Every step still moves its own inputs and its own result - nothing is skipped, the
    transfers are just overlapped with compute the way a training loop prefetches its next batch."""

    def __init__(self, engine):
        if engine.graph is None or not engine._graph_dynamic_neg:
            raise ValueError('EpochPipeline needs engine.capture(dynamic_negatives=True)')
        self.eng = engine
        nb = engine.loss.neg_buf
        self.h2d, self.d2h = torch.cuda.Stream(), torch.cuda.Stream()     # separate queues: an upload never waits behind a read-back
        self.stage = [torch.empty_like(nb) for _ in range(2)]
        self.host = [torch.empty(3, dtype=torch.float32).pin_memory() for _ in range(2)]
        self.dev_out = [torch.zeros(3, dtype=torch.float32, device=nb.device) for _ in range(2)]
        self.staged = [torch.cuda.Event() for _ in range(2)]
        self.done = [torch.cuda.Event() for _ in range(2)]
        self.read = [torch.cuda.Event() for _ in range(2)]
        self.k = 0

    def submit(self, neg_host):
        """Queue one epoch on ``neg_host`` (pinned ``[2, n_df]`` int64); returns its step index."""
        k, s = self.k, self.k & 1
        main = torch.cuda.current_stream()
        with torch.cuda.stream(self.h2d):
            if k >= 2:
                self.h2d.wait_event(self.done[s])             # step k-2 has consumed this staging buffer
            self.stage[s].copy_(neg_host, non_blocking=True)
            self.staged[s].record(self.h2d)
        main.wait_event(self.staged[s])
        self.eng.loss.neg_buf.copy_(self.stage[s], non_blocking=True)
        losses = self.eng.epoch()                             # one graph launch (rebuilds the negative incidence first)
        if k >= 2:
            main.wait_event(self.read[s])                     # step k-2's losses have left this device slot
        self.dev_out[s].copy_(losses, non_blocking=True)      # the graph's loss buffer is overwritten by the next step
        self.done[s].record(main)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(self.done[s])
            self.host[s].copy_(self.dev_out[s], non_blocking=True)
            self.read[s].record(self.d2h)
        self.k += 1
        return k

    def result(self, k):
        """(loss, loss_r, loss_l) of step ``k`` (must be one of the last two submitted); blocks until it is on the host."""
        self.read[k & 1].synchronize()
        return self.host[k & 1]
