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
    the run, so they are computed once (with the decode kernel) instead of per step."""

    def __init__(self, df_edges, neg_edges, ni_edges, num_nodes, z_ori=None, target=None, alpha=0.5):
        self.n_df = df_edges.shape[1]
        assert neg_edges.shape[1] == self.n_df, 'one negative per Df entry (gnndelete.py:221-228)'
        self.n_ni = ni_edges.shape[1]
        pu = torch.cat([df_edges[0], neg_edges[0], ni_edges[0]])
        pv = torch.cat([df_edges[1], neg_edges[1], ni_edges[1]])
        self.pairs = PairPlan(pu, pv, num_nodes)
        dev = pu.device
        if target is None:
            if self.n_ni > 0:
                p = self.pairs
                target = ops.pair_decode(z_ori, p.pu[2 * self.n_df:].contiguous(), p.pv[2 * self.n_df:].contiguous())
            else:
                target = torch.zeros(1, dtype=torch.float32, device=dev)
        self.target = target.contiguous()
        self.alpha = float(alpha)
        P = self.pairs.num_pairs
        self.logits = torch.empty(max(P, 1), dtype=torch.float32, device=dev)
        self.inc_val = torch.zeros(max(2 * P, 1), dtype=torch.float32, device=dev)
        self.losses = torch.zeros(3, dtype=torch.float32, device=dev)
        self.ws_bytes = L.load().gd_edge_loss_workspace_bytes(P)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)

    def forward(self, z):
        """Fills ``self.losses`` = (loss, loss_r, loss_l), ``self.logits`` and the incidence
        values; returns ``self.losses`` (a persistent device tensor, no host sync)."""
        p = self.pairs
        L.call('gd_edge_loss_fwd', L.ptr(z, 'f32'), z.stride(0), z.shape[1], L.ptr(p.pu), L.ptr(p.pv),
               self.n_df, self.n_ni, L.ptr(self.target), self.alpha, L.ptr(p.pos_u), L.ptr(p.pos_v),
               L.ptr(self.logits), L.ptr(self.inc_val), L.ptr(self.losses), L.ptr(self.ws), self.ws_bytes,
               L.stream())
        return self.losses

    def backward(self, z, out=None):
        """dz = d loss / d z for the ``z`` last given to :meth:`forward`."""
        return ops.spmm(self.pairs.inc, z, out=out, val=self.inc_val)


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
        self.ws_bytes = L.load().gd_dense_ni_workspace_bytes(n_s)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)

    def forward_backward(self, z, dz):
        """Adds ``weight * d loss_l / d z`` into ``dz`` (rows of S) and returns ``loss_l`` (device scalar)."""
        if self.n_s == 0 or self.num_pairs <= 0:
            return self.loss_sum * 0.0
        L.call('gd_gather_rows', L.ptr(z, 'f32'), z.stride(0), z.shape[0], L.ptr(self.S, 'i64'), self.n_s, self.dim,
               L.ptr(self.zs), self.zs.stride(0), L.ptr(self.status), L.stream())
        L.call('gd_dense_ni_fwd_bwd', L.ptr(self.zs), self.zs.stride(0), self.dim, self.n_s, L.ptr(self.tgt),
               self.tgt.stride(0), L.ptr(self.excl), self.scale, L.ptr(self.dzs), self.dzs.stride(0), L.ptr(self.loss_sum),
               L.ptr(self.ws), self.ws_bytes, L.stream())
        L.call('gd_add_rows', L.ptr(self.dzs), self.dzs.stride(0), L.ptr(self.S32), self.n_s, self.dim, L.ptr(dz),
               dz.stride(0), L.stream())
        return self.loss_sum / self.num_pairs
