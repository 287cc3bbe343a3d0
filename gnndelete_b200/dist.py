"""1-D row partition of the GCNDelete epoch over the GPUs of one node (SURVEY.md §8(e),
BASELINE config 5).  Net-new: the reference has no distributed code.

Rank p owns the contiguous destination rows ``[lo_p, hi_p)`` — their features, CSR rows,
Del masks and every loss pair that touches one of them.  Per epoch and rank:

    forward   H0_loc = D^-1/2 (X_loc W1^T)           all-gather -> H0 [N,128]   (hoistable)
              A1_loc = D^-1/2 SpMM(csr_loc, H0) + b1;  Del1;  H1_loc = D^-1/2 (relu(x1) W2^T)
                                                      all-gather -> H1 [N,64]
              A2_loc = D^-1/2 SpMM(csr_loc, H1) + b2;  Del2 -> z_loc
                                                      all-gather -> z  [N,64]
              decode + DEC/NI on the pairs touching local rows (a cross-partition pair is
              evaluated by both owners of its endpoints, counted once)          all-reduce 3 scalars
    backward  dz_loc (incidence gather over z), dW_del2, dA2_loc
                                                      all-gather -> D^-1/2 dA2 [N,64]
              dH1_loc = D^-1/2 SpMM(csr_loc, .)  (A_hat symmetric), dX1_loc, dW_del1
                                                      all-reduce dW_del1, dW_del2 (80 KB)
    Adam      replicated

``PartitionPlan`` is pure index logic (torch, any device) so it is tested on CPU with gloo;
``PartitionedGCNDeleteEngine`` runs the plan with the CUDA kernels and NCCL
(``torch.distributed`` all_gather_into_tensor / all_reduce over NVLink).
"""
from __future__ import annotations

import torch

from . import _lib as L
from . import ops
from .graph import CSR, build_csr, invert_perm


def row_bounds(num_nodes: int, world: int):
    per = -(-num_nodes // world)
    return [(min(r * per, num_nodes), min((r + 1) * per, num_nodes)) for r in range(world)]


class PartitionPlan:
    """Everything rank ``rank`` needs, as index tensors (global node ids unless named *_loc)."""

    def __init__(self, data, neg_edge_index, rank: int, world: int):
        n = int(data.num_nodes)
        self.n, self.rank, self.world = n, rank, world
        self.bounds = row_bounds(n, world)
        lo, hi = self.bounds[rank]
        self.lo, self.hi, self.n_loc = lo, hi, hi - lo
        self.per = self.bounds[0][1] - self.bounds[0][0]
        dev = data.train_pos_edge_index.device
        ei = data.train_pos_edge_index
        sdf = ei[:, data.sdf_mask]
        inside = lambda t: (t >= lo) & (t < hi)                                   # noqa: E731
        # ---- message passing: entries with a local destination, PyG self-loop handling
        m = inside(sdf[1]) & (sdf[0] != sdf[1])
        loops = torch.arange(lo, hi, device=dev, dtype=ei.dtype)
        self.mp_src = torch.cat([sdf[0][m], loops])
        self.mp_dst_loc = torch.cat([sdf[1][m] - lo, loops - lo])
        # ---- Del masks (local row ids)
        m1 = data.sdf_node_1hop_mask.to(dev)[lo:hi]
        m2 = data.sdf_node_2hop_mask.to(dev)[lo:hi]
        self.rows1_loc, self.comp1_loc = m1.nonzero().squeeze(1), (~m1).nonzero().squeeze(1)
        self.rows2_loc, self.comp2_loc = m2.nonzero().squeeze(1), (~m2).nonzero().squeeze(1)
        # ---- loss items: Df item i = (pair i, negative i); NI pair j = sdf edge with u < v
        df = ei[:, data.df_mask]
        ni = sdf[:, sdf[0] < sdf[1]]
        self.norm_df, self.norm_ni = df.shape[1], ni.shape[1]
        touch_df = inside(df[0]) | inside(df[1]) | inside(neg_edge_index[0]) | inside(neg_edge_index[1])
        own_df = inside(df[0])                                                  # counted by the owner of u
        sel_df = torch.cat([own_df.nonzero().squeeze(1), (touch_df & ~own_df).nonzero().squeeze(1)])
        self.own_df = int(own_df.sum())
        touch_ni = inside(ni[0]) | inside(ni[1])
        own_ni = inside(ni[0])
        sel_ni = torch.cat([own_ni.nonzero().squeeze(1), (touch_ni & ~own_ni).nonzero().squeeze(1)])
        self.own_ni = int(own_ni.sum())
        self.sel_df, self.sel_ni = sel_df, sel_ni
        self.n_df, self.n_ni = sel_df.numel(), sel_ni.numel()
        self.pu = torch.cat([df[0][sel_df], neg_edge_index[0][sel_df], ni[0][sel_ni]])
        self.pv = torch.cat([df[1][sel_df], neg_edge_index[1][sel_df], ni[1][sel_ni]])
        # ---- incidence entries of the LOCAL nodes: (node_loc, partner, pair, side)
        P = self.pu.numel()
        pid = torch.arange(P, device=dev)
        mu, mv = inside(self.pu), inside(self.pv)
        self.ent_node_loc = torch.cat([self.pu[mu] - lo, self.pv[mv] - lo])
        self.ent_partner = torch.cat([self.pv[mu], self.pu[mv]])
        self.ent_pair = torch.cat([pid[mu], pid[mv]])
        self.ent_side = torch.cat([torch.zeros(int(mu.sum()), dtype=torch.long, device=dev),
                                   torch.ones(int(mv.sum()), dtype=torch.long, device=dev)])

    def ni_pairs(self):
        s = 2 * self.n_df
        return self.pu[s:], self.pv[s:]


class PartitionedGCNDeleteEngine:
    """The epoch of ``engine.GCNDeleteEngine`` on one rank of the row partition."""

    def __init__(self, model, data, neg_edge_index, z_ori_full, group=None, hoist_layer1=False,
                 lr=1e-3, betas=(0.9, 0.999), eps=1e-8, alpha=0.5):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        self.model = model
        dev = data.x.device
        plan = PartitionPlan(data, neg_edge_index, rank, world)
        self.plan = plan
        n, nl, per = plan.n, plan.n_loc, plan.per
        self.x_loc = data.x[plan.lo:plan.hi].contiguous()
        # local CSR over global source ids (rows beyond n_loc are empty -> truncated view)
        full = build_csr(plan.mp_src, plan.mp_dst_loc, n, self_loops=False)
        self.csr = _truncate(full, nl)
        self.dinv = torch.empty(max(nl, 1), dtype=torch.float32, device=dev)
        L.call('gd_gcn_dinv', L.ptr(self.csr.rowptr), nl, L.ptr(self.dinv), L.stream())
        # D^-1/2 of every node (source-side factor of the transposed aggregation), gathered once
        dpad = torch.zeros(per, dtype=torch.float32, device=dev)
        dpad[:nl] = self.dinv[:nl]
        self.dinv_full = torch.empty(per * world, dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(self.dinv_full, dpad, group=group)
        i32 = lambda t: t.to(torch.int32).contiguous()                           # noqa: E731
        self.rows1, self.comp1 = i32(plan.rows1_loc), i32(plan.comp1_loc)
        self.rows2, self.comp2 = i32(plan.rows2_loc), i32(plan.comp2_loc)
        # loss: pairs index the gathered z (global ids); incidence CSR over local nodes
        self.pu, self.pv = i32(plan.pu), i32(plan.pv)
        P = plan.pu.numel()
        inc_full = build_csr(plan.ent_partner, plan.ent_node_loc, n, self_loops=False)
        self.inc = _truncate(inc_full, nl)
        pos = invert_perm(inc_full.eid, max(plan.ent_pair.numel(), 1))
        dummy = inc_full.nnz                                                      # slot for endpoints owned elsewhere
        self.pos_u = torch.full((max(P, 1),), dummy, dtype=torch.int32, device=dev)
        self.pos_v = torch.full((max(P, 1),), dummy, dtype=torch.int32, device=dev)
        side0 = plan.ent_side == 0
        self.pos_u[plan.ent_pair[side0]] = pos[side0.nonzero().squeeze(1)]
        self.pos_v[plan.ent_pair[~side0]] = pos[(~side0).nonzero().squeeze(1)]
        self.inc_val = torch.zeros(inc_full.nnz + 1, dtype=torch.float32, device=dev)
        nu, nv = plan.ni_pairs()
        self.target = ops.pair_decode(z_ori_full, i32(nu), i32(nv)) if plan.n_ni > 0 else torch.zeros(1, device=dev)
        self.logits = torch.empty(max(P, 1), dtype=torch.float32, device=dev)
        self.losses_loc = torch.zeros(3, dtype=torch.float32, device=dev)
        self.losses = torch.zeros(3, dtype=torch.float32, device=dev)
        self.ws_bytes = L.load().gd_edge_loss_workspace_bytes(P)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.alpha = float(alpha)
        hid, out = model.conv1.out_channels, model.conv2.out_channels
        f32 = dict(dtype=torch.float32, device=dev)
        pad = per * world                                                        # all_gather needs equal blocks
        self.h0_loc = torch.zeros(per, hid, **f32); self.h0 = torch.empty(pad, hid, **f32)
        self.a1 = torch.empty(nl, hid, **f32); self.x1 = torch.empty(nl, hid, **f32)
        self.h1_loc = torch.zeros(per, out, **f32); self.h1 = torch.empty(pad, out, **f32)
        self.a2 = torch.empty(nl, out, **f32)
        self.z_loc = torch.zeros(per, out, **f32); self.z = torch.empty(pad, out, **f32)
        self.dz = torch.empty(nl, out, **f32)
        self.da2_loc = torch.zeros(per, out, **f32); self.da2 = torch.empty(pad, out, **f32)
        self.dh1 = torch.empty(nl, out, **f32)
        self.dx1 = torch.zeros(nl, hid, **f32)
        self.hoist, self._layer1_done = bool(hoist_layer1), False
        self.params = [model.deletion1.deletion_weight, model.deletion2.deletion_weight]
        for p in self.params:
            p.grad = torch.zeros_like(p)
        self.gflat = torch.zeros(sum(p.numel() for p in self.params), **f32)
        self.lr, self.betas, self.eps = lr, betas, eps
        self.state = [dict(m=torch.zeros_like(p), v=torch.zeros_like(p), step=torch.zeros(1, **f32)) for p in self.params]
        self.nl = nl

    def _gather(self, loc, full):
        self.dist.all_gather_into_tensor(full, loc, group=self.group)

    def layer1(self):
        c1 = self.model.conv1
        ops.gemm_rows(self.x_loc, c1.lin.weight.detach(), True, out=self.h0_loc[:self.nl], out_scale=self.dinv)
        self._gather(self.h0_loc, self.h0)
        ops.spmm(self.csr, self.h0, out=self.a1, row_scale=self.dinv, bias=c1.bias.detach())
        self._layer1_done = True

    def forward(self):
        m, nl = self.model, self.nl
        if not (self.hoist and self._layer1_done):
            self.layer1()
        w1, w2 = m.deletion1.deletion_weight.detach(), m.deletion2.deletion_weight.detach()
        ops.gemm_rows(self.a1, w1, False, out=self.x1, rows=self.rows1)
        ops.copy_rows(self.a1, self.x1, self.comp1)
        ops.gemm_rows(self.x1, m.conv2.lin.weight.detach(), True, out=self.h1_loc[:nl], out_scale=self.dinv, relu_in=True)
        self._gather(self.h1_loc, self.h1)
        ops.spmm(self.csr, self.h1, out=self.a2, row_scale=self.dinv, bias=m.conv2.bias.detach())
        zl = self.z_loc[:nl]
        ops.gemm_rows(self.a2, w2, False, out=zl, rows=self.rows2)
        ops.copy_rows(self.a2, zl, self.comp2)
        self._gather(self.z_loc, self.z)
        p = self.plan
        L.call('gd_edge_loss_fwd_part', L.ptr(self.z), self.z.stride(0), self.z.shape[1], L.ptr(self.pu), L.ptr(self.pv),
               p.n_df, p.n_ni, L.ptr(self.target), self.alpha, L.ptr(self.pos_u), L.ptr(self.pos_v), L.ptr(self.logits),
               L.ptr(self.inc_val), L.ptr(self.losses_loc), p.own_df, p.own_ni, p.norm_df, p.norm_ni, L.ptr(self.ws),
               self.ws_bytes, L.stream())
        self.losses.copy_(self.losses_loc)
        self.dist.all_reduce(self.losses, group=self.group)
        return self.losses

    def backward(self):
        m, nl = self.model, self.nl
        g1, g2 = self.params[0].grad, self.params[1].grad
        w2 = m.deletion2.deletion_weight.detach()
        ops.spmm(self.inc, self.z, out=self.dz, val=self.inc_val)
        zl = self.z_loc[:nl]
        ops.gemm_tn_rows(self.a2, self.dz, rows=self.rows2, out=g2)
        dal = self.da2_loc[:nl]
        ops.gemm_rows(self.dz, w2, True, out=dal, rows=self.rows2)
        ops.copy_rows(self.dz, dal, self.comp2)
        self._gather(self.da2_loc, self.da2)
        ops.spmm(self.csr, self.da2, out=self.dh1, col_scale=self.dinv_full)                    # A^T = A (symmetric)
        ops.gemm_rows(self.dh1, m.conv2.lin.weight.detach(), False, out=self.dx1, rows=self.rows1,
                      out_scale=self.dinv, gate=self.x1)
        ops.gemm_tn_rows(self.a1, self.dx1, rows=self.rows1, out=g1)
        n1 = g1.numel()
        self.gflat[:n1].copy_(g1.view(-1)); self.gflat[n1:].copy_(g2.view(-1))
        self.dist.all_reduce(self.gflat, group=self.group)
        g1.view(-1).copy_(self.gflat[:n1]); g2.view(-1).copy_(self.gflat[n1:])
        del zl

    def adam_step(self):
        for p, st in zip(self.params, self.state):
            ops.adam_step(p.data, p.grad, st['m'], st['v'], st['step'], self.lr, self.betas[0], self.betas[1], self.eps)

    def epoch(self):
        losses = self.forward()
        self.backward()
        self.adam_step()
        return losses


def _truncate(csr: CSR, num_rows: int) -> CSR:
    """View of a CSR whose rows beyond ``num_rows`` are empty."""
    out = CSR(csr.rowptr[:num_rows + 1].contiguous(), csr.col, csr.eid, csr.rel, num_rows, csr.nnz,
              csr.plan if csr.plan else None, None, None)
    return out
