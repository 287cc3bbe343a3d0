"""Decoder + Deleted-Edge-Consistency / Neighbourhood-Influence losses on the device.

``PairPlan`` holds a fixed list of node pairs together with its *incidence CSR* (for
every node, the pairs it takes part in and the partner node).  The forward kernel
(``gd_edge_loss_fwd``) writes d loss / d logit of each pair into the pair's two
incidence slots; the gradient w.r.t. the embeddings is then a deterministic CSR gather
(``gd_spmm``) instead of the atomics-based ``index_add`` scatter autograd performs for
``z[edge_index[0]] * z[edge_index[1]]`` (reference gcn.py:26-35, gnndelete.py:227-250,
:362-398).
"""
from __future__ import annotations

import os

import torch

from . import _lib as L
from . import ops
from .graph import build_csr, invert_perm


class PairPlan:
    def __init__(self, pu, pv, num_nodes, rel_weight=None, pair_rel=None):
        pu = pu.to(torch.int64).contiguous()
        pv = pv.to(torch.int64).contiguous()
        self.num_pairs = pu.numel()
        self.num_nodes = int(num_nodes)
        self.pu = pu.to(torch.int32)
        self.pv = pv.to(torch.int32)
        self.rel_weight, self.pair_rel = rel_weight, pair_rel
        # incidence entry e < P: node u_e receives partner v_e; entry e >= P: node v receives u
        dst = torch.cat([pu, pv])
        src = torch.cat([pv, pu])
        self.inc = build_csr(src, dst, num_nodes, self_loops=False)
        pos = invert_perm(self.inc.eid, 2 * self.num_pairs)
        self.pos_u = pos[:self.num_pairs].contiguous()
        self.pos_v = pos[self.num_pairs:].contiguous()
        self._inc_pair = None

    @property
    def inc_pair(self):
        """pair id of every incidence entry (only needed by the generic decode backward)."""
        if self._inc_pair is None:
            self._inc_pair = (self.inc.eid % max(self.num_pairs, 1)).to(torch.int32)
        return self._inc_pair


class EdgeLossPlan:
    """Pairs ``[Df | supplied negatives | S_Df edges with u<v]`` + NI targets.

    ``target`` are the original model's logits on the NI pairs
    (``(z_ori[row] * z_ori[col]).sum(-1)``, gnndelete.py:383); they are constant over
    the run, so they are computed once (with the decode kernel) instead of per step.

    Two execution modes, chosen by the embedding width at the first :meth:`forward`:

    ``node`` (widths 32 / 64 / 128, the hot path): losses and ``dz`` come out of ONE pass over the
    node -> incident-pair lists (``gd_node_loss_fwd_bwd``, csrc/node_loss.cu).  A node keeps its own
    row in registers, gathers one partner row per incident pair, recomputes the NI logit from its
    side (``g = c_l (<z_w, z_x> - target)``) and accumulates ``g z_x``; the DEC pairs (Df pair i and
    its negative share one residual) get their coefficient from a small pre-pass over the ``2 n_df``
    DEC pairs (``gd_edge_loss_fwd`` with no NI pairs).  One row gather per incidence entry in total,
    instead of two per pair in the forward plus one per entry in the backward.

    ``pair`` (any other width): ``gd_edge_loss_fwd`` over all pairs writes d loss / d logit into the
    incidence slots and ``dz`` is a weighted CSR gather.

    In both modes the incidence is split in two: a FIXED part over the Df + NI pairs (built once)
    and a small CSR over the negative pairs that :meth:`update_negatives` rebuilds in place —
    the reference draws new negatives every epoch (gnndelete.py:221-225).  The update is a
    radix sort of ``2 n_df`` keys into preallocated buffers with no host synchronisation, so it
    can live inside the epoch's CUDA graph."""

    NODE_ROW_LIMIT = 1 << 24          # gd_node_loss_fwd_bwd packs the row id of a batch into 24 bits

    def __init__(self, df_edges, neg_edges, ni_edges, num_nodes, z_ori=None, target=None, alpha=0.5,
                 static_negatives=False, deterministic=False):
        """``static_negatives``: the supplied negatives never change (SURVEY.md §8(d)'s epoch definition), so
        they join the fixed incidence; :meth:`update_negatives` then rebuilds the whole incidence (slow path).
        Default: negatives are replaceable every step without a rebuild.  In ``node`` mode the gradient of the
        replaceable negative pairs is then ADDED onto ``dz`` with vector float reductions (``gd_pair_scatter_add``):
        no per-step sort, but the order of a row's few negative contributions is not fixed (like the reference's
        ``index_add``); ``deterministic=True`` keeps the sorted per-step incidence (tail CSR) instead - bitwise
        reproducible, ~0.2 ms slower per Collab-sized epoch."""
        dev = df_edges.device
        n_df, n_ni, n = df_edges.shape[1], ni_edges.shape[1], int(num_nodes)
        assert neg_edges.shape[1] == n_df, 'one negative per Df entry (gnndelete.py:221-228)'
        self.n_df, self.n_ni, self.num_nodes = n_df, n_ni, n
        self.static = bool(static_negatives)
        self.deterministic = bool(deterministic)
        self.neg_atomic = False
        P = 2 * n_df + n_ni
        self.num_pairs = P
        self.pu = torch.cat([df_edges[0], neg_edges[0], ni_edges[0]]).to(torch.int32).contiguous()
        self.pv = torch.cat([df_edges[1], neg_edges[1], ni_edges[1]]).to(torch.int32).contiguous()
        if self.pu.numel() == 0:
            self.pu = torch.zeros(1, dtype=torch.int32, device=dev)
            self.pv = torch.zeros(1, dtype=torch.int32, device=dev)
        self.pos_u = torch.zeros(max(P, 1), dtype=torch.int32, device=dev)
        self.pos_v = torch.zeros(max(P, 1), dtype=torch.int32, device=dev)
        # ---- negative incidence: persistent buffers, rebuilt by update_negatives (dynamic mode only)
        m = 0 if self.static else 2 * n_df
        self.neg_m = m
        self.neg_src = torch.zeros(max(m, 1), dtype=torch.int64, device=dev)
        self.neg_dst = torch.zeros(max(m, 1), dtype=torch.int64, device=dev)
        self.neg_rowptr = torch.zeros(n + 1, dtype=torch.int32, device=dev)
        self.neg_col = torch.zeros(max(m, 1), dtype=torch.int32, device=dev)
        self.neg_eid = torch.zeros(max(m, 1), dtype=torch.int32, device=dev)
        self.neg_pos = torch.zeros(max(m, 1), dtype=torch.int32, device=dev)
        self.neg_status = torch.zeros(2, dtype=torch.int32, device=dev)
        self.neg_ws_bytes = L.load().gd_csr_workspace_bytes(max(m, 1), n)
        self.neg_ws = torch.empty(self.neg_ws_bytes, dtype=torch.uint8, device=dev)
        from .graph import CSR
        self.inc_neg = CSR(self.neg_rowptr, self.neg_col, self.neg_eid, None, n, m)
        self.inc_neg.dynamic = True                  # rebuilt in place every step: no cached batch plan
        self.neg_buf = neg_edges.clone()             # staging buffer a caller may overwrite (H2D) before a graph replay
        if target is None:
            if n_ni > 0:
                target = ops.pair_decode(z_ori, self.pu[2 * n_df:P].contiguous(), self.pv[2 * n_df:P].contiguous())
            else:
                target = torch.zeros(1, dtype=torch.float32, device=dev)
        self.target = target.contiguous()
        self.alpha = float(alpha)
        self.logits = torch.empty(max(P, 1), dtype=torch.float32, device=dev)
        self.losses = torch.zeros(3, dtype=torch.float32, device=dev)
        self.dec_losses = torch.zeros(3, dtype=torch.float32, device=dev)
        self.ni_sq_sum = torch.zeros(1, dtype=torch.float32, device=dev)
        self.ws_bytes = L.load().gd_edge_loss_workspace_bytes(P)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self._feat = None
        self._dz = None
        self.mode = None
        self._layout(z_ori.shape[1] if z_ori is not None else 0)

    # ------------------------------------------------------------------ plan construction
    def _mode_for(self, feat):
        from .graph import BATCHED
        ok = BATCHED and feat in (32, 64, 128) and 0 < self.num_nodes < self.NODE_ROW_LIMIT and self.num_pairs > 0
        return 'node' if ok else 'pair'

    def _fixed_pairs(self):
        """Indices (into ``pu`` / ``pv``) of the DEC pairs that never change: Df, plus the negatives in static mode."""
        dev, n_df = self.pu.device, self.n_df
        return torch.arange(2 * n_df if self.static else n_df, device=dev)

    def _build_fixed_pair(self):
        """``pair`` mode: fixed incidence CSR (node -> (partner, pair side)) over the pairs that never change:
        Df + NI, plus the negatives in static mode.  Entry e < Pf is the u side of fixed pair e, entry Pf + e its v side."""
        dev, n_df, P = self.pu.device, self.n_df, self.num_pairs
        if self.static:
            self.fixed_idx = torch.arange(P, device=dev)
        else:
            self.fixed_idx = torch.cat([torch.arange(n_df, device=dev), torch.arange(2 * n_df, P, device=dev)])
        fu, fv = self.pu[:P][self.fixed_idx].long(), self.pv[:P][self.fixed_idx].long()
        Pf = int(self.fixed_idx.numel())
        self.inc_fixed = build_csr(torch.cat([fv, fu]), torch.cat([fu, fv]), self.num_nodes, self_loops=False)
        self._posf = invert_perm(self.inc_fixed.eid, max(2 * Pf, 1))
        self.nnz_fixed = 2 * Pf

    def _layout_pair(self, feat):
        """``pair`` mode: per-entry gradient buffer ``inc_val`` = [fixed incidence | negative incidence].
        When the batched aggregation covers ``feat`` the fixed part is in the padded slot layout of its
        batch plan (the loss kernel scatters d loss / d logit through ``pos_u`` / ``pos_v``, so the
        remap is free); otherwise it is in CSR order for the row-walking kernel."""
        self._build_fixed_pair()
        n_df, Pf = self.n_df, int(self.fixed_idx.numel())
        bp = self.inc_fixed.bplan(feat, True) if (feat and self.nnz_fixed) else None
        self.bplan_fixed = bp
        posf = self._posf
        if bp is not None:
            posf = bp.slot_of_entry[posf[:2 * Pf].long()].to(torch.int32)
            self.val_off = bp.num_slots                        # incl. the plan's two batches of slack
        else:
            self.val_off = max(self.nnz_fixed, 1)
        self.pos_u[self.fixed_idx] = posf[:Pf]
        self.pos_v[self.fixed_idx] = posf[Pf:2 * Pf]
        self.inc_val = torch.zeros(self.val_off + max(self.neg_m, 1), dtype=torch.float32, device=self.pos_u.device)

    def _layout_node(self, feat):
        """``node`` mode: batch plan over the node -> incident-pair lists (NI pairs from both endpoints, fixed DEC
        pairs from both endpoints), per-slot values (NI target | DEC coefficient slot), per-batch meta words."""
        from .graph import BatchPlan
        dev, n, n_df, n_ni = self.pu.device, self.num_nodes, self.n_df, self.n_ni
        P = self.num_pairs
        nu, nv = self.pu[2 * n_df:P].long(), self.pv[2 * n_df:P].long()
        fidx = self._fixed_pairs()
        fu, fv = self.pu[:P][fidx].long(), self.pv[:P][fidx].long()
        Pf = int(fidx.numel())
        dst = torch.cat([nu, nv, fu, fv])
        src = torch.cat([nv, nu, fv, fu])
        inc = build_csr(src, dst, n, self_loops=False)
        self.inc_fixed = inc
        self.nnz_fixed = inc.nnz
        posf = invert_perm(inc.eid, max(inc.nnz, 1)).long()
        workers = L.load().gd_node_loss_workers(int(feat), 0)
        bp = BatchPlan(inc.rowptr, inc.col, n, inc.nnz, workers)
        bp.colp.clamp_(min=0)       # padding slots gather a valid row (coefficient 0): the kernel's loads are unpredicated
        self.bplan_fixed = bp
        slot = bp.slot_of_entry[posf[:inc.nnz]] if inc.nnz else posf[:0]
        self.val_off = bp.num_slots
        self.inc_val = torch.zeros(self.val_off + max(self.neg_m, 1), dtype=torch.float32, device=dev)
        if self.neg_atomic:          # negative pair i: both "slots" are element val_off + i (its coefficient, read by the scatter)
            ar = torch.arange(n_df, dtype=torch.int32, device=dev) + self.val_off
            self.pos_u[n_df:2 * n_df] = ar
            self.pos_v[n_df:2 * n_df] = ar
        flag = torch.zeros(bp.num_slots, dtype=torch.bool, device=dev)
        if n_ni:
            s_ni = slot[:2 * n_ni]
            self.inc_val[s_ni] = torch.cat([self.target[:n_ni], self.target[:n_ni]])
            flag[s_ni] = True
        self._ni_slots = slot[:2 * n_ni]
        self.pos_u[fidx] = slot[2 * n_ni:2 * n_ni + Pf].to(torch.int32)
        self.pos_v[fidx] = slot[2 * n_ni + Pf:].to(torch.int32)
        # meta word of every batch: row (low 24 bits) | NI flag of each of its 8 slots (high 8 bits)
        deg = (inc.rowptr[1:] - inc.rowptr[:-1]).long()
        nbr = torch.clamp((deg + 7) // 8, min=1)
        brow = torch.repeat_interleave(torch.arange(n, device=dev), nbr)
        assert brow.numel() == bp.num_batches
        bits = (flag.view(-1, 8).long() << torch.arange(8, device=dev)).sum(1)          # [num_batches + 2]
        meta = bits << 24
        meta[:bp.num_batches] |= brow
        self.bmeta = torch.where(meta >= (1 << 31), meta - (1 << 32), meta).to(torch.int32).contiguous()
        self.nl_ws_bytes = L.load().gd_node_loss_workspace_bytes(bp.num_workers)
        self.nl_ws = torch.empty(self.nl_ws_bytes, dtype=torch.uint8, device=dev)

    def _layout(self, feat):
        if self._feat == feat:
            return
        self._feat = feat
        self.mode = self._mode_for(feat)
        self.neg_atomic = self.mode == 'node' and not self.static and not self.deterministic and self.n_df > 0
        if self.mode == 'node':
            self._layout_node(feat)
        else:
            self._layout_pair(feat)
        self.update_negatives()

    def update_negatives(self, neg_edges=None):
        """Install new negatives (default: whatever is in ``self.neg_buf``).  Dynamic mode: no allocation,
        no host sync (capturable).  Static mode: rebuilds the whole incidence (not capturable)."""
        n_df = self.n_df
        if n_df == 0:
            return
        if neg_edges is not None:
            self.neg_buf.copy_(neg_edges)
        nu, nv = self.neg_buf[0], self.neg_buf[1]
        self.pu[n_df:2 * n_df].copy_(nu)
        self.pv[n_df:2 * n_df].copy_(nv)
        if self.static:
            if neg_edges is not None:
                bad = int(((self.neg_buf < 0) | (self.neg_buf >= self.num_nodes)).sum().item())
                self.neg_status[1] = bad
                if bad:
                    return
                feat, self._feat = self._feat, None
                self._layout(feat)
            return
        if self.neg_atomic:
            return
        self.neg_dst[:n_df].copy_(nu); self.neg_dst[n_df:].copy_(nv)
        self.neg_src[:n_df].copy_(nv); self.neg_src[n_df:].copy_(nu)
        m = 2 * n_df
        L.call('gd_csr_from_coo', L.ptr(self.neg_src), L.ptr(self.neg_dst), None, m, self.num_nodes, 1, 2,   # rows ordered only
               L.ptr(self.neg_rowptr), L.ptr(self.neg_col), L.ptr(self.neg_eid), None, L.ptr(self.neg_status),
               L.ptr(self.neg_ws), self.neg_ws_bytes, L.stream())
        L.call('gd_invert_perm', L.ptr(self.neg_eid), m, L.ptr(self.neg_pos), L.stream())
        torch.add(self.neg_pos[:n_df], self.val_off, out=self.pos_u[n_df:2 * n_df])
        torch.add(self.neg_pos[n_df:m], self.val_off, out=self.pos_v[n_df:2 * n_df])

    def check_negatives(self):
        """Host-side validation of the last update (synchronises): raises on out-of-range endpoints."""
        bad = int(self.neg_status[1].item())
        if bad:
            raise ValueError(f'negative edges hold {bad} endpoints outside [0, {self.num_nodes})')

    # ------------------------------------------------------------------------- execution
    def forward(self, z, dz_out=None):
        """Fills ``self.losses`` = (loss, loss_r, loss_l) and returns it (a persistent device tensor, no host sync).
        ``node`` mode also produces ``dz`` here (into ``dz_out`` when given); ``self.logits`` then holds the DEC
        logits only (``[:2 n_df]``)."""
        self._layout(z.shape[1])
        if self.mode == 'pair':
            L.call('gd_edge_loss_fwd', L.ptr(z, 'f32'), z.stride(0), z.shape[1], L.ptr(self.pu), L.ptr(self.pv),
                   self.n_df, self.n_ni, L.ptr(self.target), self.alpha, L.ptr(self.pos_u), L.ptr(self.pos_v),
                   L.ptr(self.logits), L.ptr(self.inc_val), L.ptr(self.losses), L.ptr(self.ws), self.ws_bytes,
                   L.stream())
            return self.losses
        if self.n_df:            # DEC pre-pass: logits of the Df pairs and their negatives -> coefficient slots, loss_r
            L.call('gd_edge_loss_fwd', L.ptr(z, 'f32'), z.stride(0), z.shape[1], L.ptr(self.pu), L.ptr(self.pv),
                   self.n_df, 0, None, self.alpha, L.ptr(self.pos_u), L.ptr(self.pos_v), L.ptr(self.logits),
                   L.ptr(self.inc_val), L.ptr(self.dec_losses), L.ptr(self.ws), self.ws_bytes, L.stream())
        dz = dz_out
        if dz is None:
            if self._dz is None or self._dz.shape != z.shape:
                self._dz = torch.empty_like(z)
            dz = self._dz
        self._dz_last = dz
        bp = self.bplan_fixed
        tail = self.neg_m > 0 and not self.neg_atomic
        L.call('gd_node_loss_fwd_bwd', bp.ref, L.ptr(self.bmeta), L.ptr(self.inc_val),
               L.ptr(self.neg_rowptr) if tail else None, L.ptr(self.neg_col) if tail else None,
               L.ptr(self.inc_val[self.val_off:]) if tail else None,
               L.ptr(z, 'f32'), z.stride(0), L.ptr(z, 'f32'), z.stride(0), 0, None, z.shape[1], self.n_ni, self.alpha,
               L.ptr(self.dec_losses), L.ptr(dz), dz.stride(0), L.ptr(bp.scratch(z.shape[1])), L.ptr(self.losses),
               L.ptr(self.ni_sq_sum), L.ptr(self.nl_ws), self.nl_ws_bytes, L.stream())
        if self.neg_atomic:
            n_df = self.n_df
            L.call('gd_pair_scatter_add', L.ptr(z, 'f32'), z.stride(0), z.shape[1], L.ptr(self.pu[n_df:2 * n_df]),
                   L.ptr(self.pv[n_df:2 * n_df]), L.ptr(self.inc_val[self.val_off:]), n_df, L.ptr(dz), dz.stride(0),
                   L.stream())
        return self.losses

    def backward(self, z, out=None):
        """dz = d loss / d z for the ``z`` last given to :meth:`forward`."""
        if self.mode == 'node':
            dz = self._dz_last
            if out is None:
                return dz.clone() if dz is self._dz else dz
            if out.data_ptr() != dz.data_ptr():
                out.copy_(dz)
            return out
        if self.bplan_fixed is not None:
            # one gather: the fixed incidence through its batch plan, this step's negative incidence as the tail CSR
            tail = (self.neg_rowptr, self.neg_col, self.inc_val[self.val_off:]) if self.neg_m > 0 else None
            return ops.spmm(self.inc_fixed, z, out=out, valp=self.inc_val[:self.val_off], tail=tail)
        else:
            out = ops.spmm(self.inc_fixed, z, out=out, val=self.inc_val[:self.val_off])
        if self.neg_m > 0:
            ops.spmm(self.inc_neg, z, out=out, val=self.inc_val[self.val_off:], accumulate=True)
        return out


class EdgeLossFn(torch.autograd.Function):
    """(loss, loss_r, loss_l) as one differentiable op; only ``loss`` carries gradient
    (``loss_r`` / ``loss_l`` are logged, gnndelete.py:264-270)."""

    @staticmethod
    def forward(ctx, z, plan):
        z = z.contiguous()
        ctx.plan = plan
        ctx.save_for_backward(z)
        out = plan.forward(z).clone()
        ctx.mark_non_differentiable()
        return out

    @staticmethod
    def backward(ctx, gout):
        (z,) = ctx.saved_tensors
        dz = ctx.plan.backward(z)
        # d(out[0]) only; gout[0] is the upstream scale (1 for loss.backward())
        return dz * gout[0], None


def edge_loss(z, plan):
    out = EdgeLossFn.apply(z, plan)
    return out[0], out[1].detach(), out[2].detach()


class DenseNIPlan:
    """Dense-block Neighbourhood-Influence term of ``train_fullbatch`` (gnndelete.py:163-193,
    239-241): ``MSE(sigmoid(z z^T)[M], sigmoid(logits_ori)[M])`` with ``M`` = strictly-lower node
    pairs inside the 2-hop node set minus the Df pairs.  Only the ``S x S`` block is ever
    touched (``gd_dense_ni_fwd_bwd``); the target block ``sigmoid(logits_ori[S][:, S])`` and the
    excluded-pair bitmap are built once."""

    def __init__(self, node_mask, df_edges, logits_ori, num_nodes, dim, weight=0.5):
        dev = df_edges.device
        node_mask = node_mask.to(dev)
        self.S = node_mask.nonzero().squeeze(1)
        self.S32 = self.S.to(torch.int32)
        n_s = self.S.numel()
        self.n_s, self.dim = n_s, int(dim)
        pos = torch.full((int(num_nodes),), -1, dtype=torch.int64, device=dev)
        pos[self.S] = torch.arange(n_s, device=dev)
        lo = logits_ori.to(dev) if logits_ori.device != dev else logits_ori
        self.tgt = torch.sigmoid(lo[self.S][:, self.S].float()).contiguous()
        pu, pv = pos[df_edges[0]], pos[df_edges[1]]
        ok = (pu >= 0) & (pv >= 0) & (pu != pv)
        pu, pv = pu[ok], pv[ok]
        idx = torch.unique(torch.cat([pu * n_s + pv, pv * n_s + pu]))
        words = torch.zeros((n_s * n_s + 31) // 32 + 1, dtype=torch.int64, device=dev)
        words.index_add_(0, idx >> 5, torch.ones_like(idx) << (idx & 31))
        self.excl = (words & 0xFFFFFFFF).to(torch.int64)
        self.excl = torch.where(self.excl >= 2 ** 31, self.excl - 2 ** 32, self.excl).to(torch.int32).contiguous()
        self.num_pairs = n_s * (n_s - 1) // 2 - idx.numel() // 2
        self.weight = float(weight)
        self.scale = self.weight / self.num_pairs if self.num_pairs > 0 else 0.0
        self.zs = torch.empty(max(n_s, 1), self.dim, dtype=torch.float32, device=dev)
        self.dzs = torch.empty(max(n_s, 1), self.dim, dtype=torch.float32, device=dev)
        self.loss_sum = torch.zeros(1, dtype=torch.float32, device=dev)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        lib = L.load()
        # tensor-core kernel (tcgen05, 3xTF32) for the reference's 64-wide embeddings; GD_DENSE_NI=simt keeps the fp32
        # CUDA-core kernel (the only one for other widths)
        self.tensor_core = bool(lib.gd_dense_ni_tc_supported(self.dim)) and os.environ.get('GD_DENSE_NI', 'tc') != 'simt' and n_s > 0
        if self.tensor_core:
            self.packed = torch.empty(lib.gd_dense_ni_tc_target_bytes(n_s) // 4, dtype=torch.float32, device=dev)
            L.call('gd_dense_ni_tc_pack_target', L.ptr(self.tgt), self.tgt.stride(0), L.ptr(self.excl), n_s, L.ptr(self.packed), L.stream())
            torch.cuda.current_stream().synchronize()
            self.tgt = self.excl = None                  # the packed tiles carry both (sentinel -1 = pair not in M)
            self.ws_bytes = lib.gd_dense_ni_tc_workspace_bytes(n_s)
        else:
            self.ws_bytes = lib.gd_dense_ni_workspace_bytes(n_s)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)

    def forward_backward(self, z, dz):
        """Adds ``weight * d loss_l / d z`` into ``dz`` (rows of S) and returns ``loss_l`` (device scalar)."""
        if self.n_s == 0 or self.num_pairs <= 0:
            return self.loss_sum * 0.0
        L.call('gd_gather_rows', L.ptr(z, 'f32'), z.stride(0), z.shape[0], L.ptr(self.S, 'i64'), self.n_s, self.dim,
               L.ptr(self.zs), self.zs.stride(0), L.ptr(self.status), L.stream())
        if self.tensor_core:
            L.call('gd_dense_ni_tc_fwd_bwd', L.ptr(self.zs), self.zs.stride(0), self.n_s, L.ptr(self.packed), self.scale,
                   L.ptr(self.dzs), self.dzs.stride(0), L.ptr(self.loss_sum), L.ptr(self.ws), self.ws_bytes, L.stream())
        else:
            L.call('gd_dense_ni_fwd_bwd', L.ptr(self.zs), self.zs.stride(0), self.dim, self.n_s, L.ptr(self.tgt),
                   self.tgt.stride(0), L.ptr(self.excl), self.scale, L.ptr(self.dzs), self.dzs.stride(0), L.ptr(self.loss_sum),
                   L.ptr(self.ws), self.ws_bytes, L.stream())
        L.call('gd_add_rows', L.ptr(self.dzs), self.dzs.stride(0), L.ptr(self.S32), self.n_s, self.dim, L.ptr(dz),
               dz.stride(0), L.stream())
        return self.loss_sum / self.num_pairs


class DenseNIFn(torch.autograd.Function):
    """``weight * loss_l`` of :class:`DenseNIPlan` as a differentiable op (forward and ``dz`` come out of the same
    kernel pass); second output: the unweighted ``loss_l`` for logging."""

    @staticmethod
    def forward(ctx, z, plan):
        z = z.contiguous()
        dz = torch.zeros_like(z)
        loss_l = plan.forward_backward(z, dz).reshape(())
        ctx.dz = dz
        out = torch.stack([loss_l * plan.weight, loss_l])
        return out

    @staticmethod
    def backward(ctx, gout):
        return ctx.dz * gout[0], None


def dense_ni_loss(z, plan):
    out = DenseNIFn.apply(z, plan)
    return out[0], out[1].detach()


def row_mse_incidence(dst_r, src_r, rows_l, num_rows):
    """Destination-major incidence of the two node-embedding MSE terms (``gd_row_mse_fwd_bwd``).

    Term 0 (``loss_r``): row ``dst_r[i]`` of ``z`` is compared with row ``src_r[i]`` of ``z_ori``;
    term 1 (``loss_l``): row ``rows_l[j]`` with the same row of ``z_ori``.  Returns int32
    ``(rowptr[num_rows + 1], code[nnz])`` with ``code = src`` for term 0 and ``-1 - row`` for term 1;
    within a row the entries keep their input order, term 0 first (stable sort), which fixes the
    summation order.  Plain tensor ops: runs on whatever device the index tensors live on (setup time)."""
    dst_r, src_r, rows_l = dst_r.reshape(-1).long(), src_r.reshape(-1).long(), rows_l.reshape(-1).long()
    if dst_r.numel() != src_r.numel():
        raise ValueError('loss_r compares equally long row lists (gnndelete_nodeemb.py:200-204)')
    n = int(num_rows)
    for name, t in (('dst_r', dst_r), ('src_r', src_r), ('rows_l', rows_l)):
        if t.numel() and (int(t.min()) < 0 or int(t.max()) >= n):
            raise IndexError(f'{name} holds row ids outside [0, {n})')
    dst = torch.cat([dst_r, rows_l])
    code = torch.cat([src_r, -1 - rows_l])
    order = torch.argsort(dst, stable=True)
    counts = torch.bincount(dst, minlength=n)
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dst.device)
    torch.cumsum(counts, 0, out=rowptr[1:])
    return rowptr.to(torch.int32).contiguous(), code[order].to(torch.int32).contiguous()


class RowMSEPlan:
    """One layer's node-embedding losses against the frozen original embeddings ``z_ori``
    (gnndelete_nodeemb.py:196-212; KG step :770-781):

        loss_r = loss_fct(cat(z[pos[0]], z[pos[1]]), cat(z_ori[neg[0]], z_ori[neg[1]]))
        loss_l = loss_fct(z[node_mask], z_ori[node_mask])

    with ``loss_fct`` = ``nn.MSELoss(reduction)``.  ``mix = (a_r, a_l)`` are the weights of the objective
    that carries gradient (``alpha, 1 - alpha`` for the layer-wise types, ``alpha, 1`` for ``only*``).
    ``pos`` / ``neg`` / ``node_mask`` are fixed for the run (the reference samples ``neg_edge`` once,
    :186-189), so the incidence is built once; :meth:`set_pairs` rebuilds it for resampled negatives."""

    def __init__(self, pos, neg, node_mask, z_ori, mix=(0.5, 0.5), reduction='mean'):
        if reduction not in ('mean', 'sum'):
            raise NotImplementedError(reduction)
        self.z_ori = z_ori.detach().contiguous()
        if self.z_ori.dtype != torch.float32:
            raise RuntimeError('gnndelete_b200 kernels compute in fp32')
        self.n, self.dim = self.z_ori.shape
        self.reduction = reduction
        self.mix = (float(mix[0]), float(mix[1]))
        dev = self.z_ori.device
        self.rows_l = node_mask.nonzero().squeeze(1) if node_mask.dtype == torch.bool else node_mask.reshape(-1).long()
        self.losses = torch.zeros(3, dtype=torch.float32, device=dev)
        self.ws_bytes = L.load().gd_row_mse_workspace_bytes(self.n)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.set_pairs(pos, neg)

    def set_pairs(self, pos, neg):
        dev = self.z_ori.device
        dst_r = torch.cat([pos[0], pos[1]]).to(dev)
        src_r = torch.cat([neg[0], neg[1]]).to(dev)
        self.m_r, self.m_l = dst_r.numel(), self.rows_l.numel()
        self.rowptr, self.code = row_mse_incidence(dst_r, src_r, self.rows_l.to(dev), self.n)
        if self.code.numel() == 0:
            self.code = torch.zeros(1, dtype=torch.int32, device=dev)
        nan = float('nan')                           # MSELoss('mean') of an empty selection is nan in the reference too
        if self.reduction == 'mean':
            self.w = (1.0 / (self.m_r * self.dim) if self.m_r else nan, 1.0 / (self.m_l * self.dim) if self.m_l else nan)
        else:
            self.w = (1.0, 1.0)

    def forward_backward(self, z, dz=None, want_grad=True):
        """Fills ``self.losses`` = (a_r loss_r + a_l loss_l, loss_r, loss_l) and, unless ``want_grad`` is off,
        ``dz`` = d losses[0] / d z (every row written).  No host sync."""
        if z.shape != (self.n, self.dim):
            raise ValueError(f'expected embeddings of shape {(self.n, self.dim)}, got {tuple(z.shape)}')
        if want_grad and dz is None:
            dz = torch.empty(self.n, self.dim, dtype=torch.float32, device=z.device)
        L.call('gd_row_mse_fwd_bwd', L.ptr(z, 'f32'), z.stride(0), L.ptr(self.z_ori), self.z_ori.stride(0), self.dim, self.n,
               L.ptr(self.rowptr), L.ptr(self.code), self.w[0], self.w[1], self.mix[0], self.mix[1],
               L.ptr(dz) if want_grad else None, dz.stride(0) if want_grad else 0, L.ptr(self.losses),
               L.ptr(self.ws), self.ws_bytes, L.stream())
        return self.losses, dz


class RowMSEFn(torch.autograd.Function):
    """(a_r loss_r + a_l loss_l, loss_r, loss_l) of one layer as one differentiable op; only the first carries
    gradient.  Forward and gradient come out of the same kernel pass, so ``backward`` is a scale — it may be
    called repeatedly (``retain_graph=True`` in the layer-wise schedule, gnndelete_nodeemb.py:224-232)."""

    @staticmethod
    def forward(ctx, z, plan):
        z = z.contiguous()
        losses, dz = plan.forward_backward(z, want_grad=ctx.needs_input_grad[0])
        ctx.dz = dz
        return losses.clone()

    @staticmethod
    def backward(ctx, gout):
        return ctx.dz * gout[0], None


def row_mse(z, plan):
    out = RowMSEFn.apply(z, plan)
    return out[0], out[1].detach(), out[2].detach()
