"""CPU, world_size 2 over gloo: the row-partition index logic (gnndelete_b200.dist.PartitionPlan) reproduces
the single-process oracle when every rank evaluates its part with plain torch math and the halo
exchanges are all_gathers — i.e. the decomposition the CUDA engine executes is correct."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from gnndelete_b200.dist import PartitionPlan
        from oracle import unlearn as OU
        from tests import util as U
        torch.set_num_threads(1)
        shape, raw, df, data, neg = U.make_case('cora', 0.03)
        om = U.oracle_model('gcn', shape, data, dtype=torch.float64)
        d64 = data.clone(); d64.x = data.x.double()
        with torch.no_grad():
            zo = om.get_original_embeddings(d64.x, d64.train_pos_edge_index[:, d64.dr_mask])
        loss_o, lr_o, ll_o, _ = OU.edge_form_loss(om, d64, neg, zo)
        loss_o.backward()

        plan = PartitionPlan(data, neg, rank, world)
        n, nl, lo, hi, per = plan.n, plan.n_loc, plan.lo, plan.hi, plan.per
        W1, b1 = om.conv1.lin.weight.detach(), om.conv1.bias.detach()
        W2, b2 = om.conv2.lin.weight.detach(), om.conv2.bias.detach()
        D1 = om.deletion1.deletion_weight.detach().clone().requires_grad_(True)
        D2 = om.deletion2.deletion_weight.detach().clone().requires_grad_(True)

        def gather(loc):
            pad = torch.zeros(per, loc.shape[1], dtype=loc.dtype)
            pad[:nl] = loc
            parts = [torch.zeros_like(pad) for _ in range(world)]
            dist.all_gather(parts, pad.detach())
            full = torch.cat(parts)
            full[lo:hi] = loc          # keep the autograd path through the local block
            return full

        deg = torch.zeros(nl, dtype=torch.float64).index_add_(0, plan.mp_dst_loc, torch.ones(plan.mp_src.numel(), dtype=torch.float64))
        dinv = deg.pow(-0.5)

        def agg(h_full):
            out = torch.zeros(nl, h_full.shape[1], dtype=torch.float64)
            return out.index_add(0, plan.mp_dst_loc, h_full[plan.mp_src])

        x_loc = d64.x[lo:hi]
        a1 = dinv.view(-1, 1) * agg(gather(dinv.view(-1, 1) * (x_loc @ W1.t()))) + b1
        x1 = a1.clone(); x1[plan.rows1_loc] = a1[plan.rows1_loc] @ D1
        a2 = dinv.view(-1, 1) * agg(gather(dinv.view(-1, 1) * (x1.relu() @ W2.t()))) + b2
        z_loc = a2.clone(); z_loc[plan.rows2_loc] = a2[plan.rows2_loc] @ D2
        z = gather(z_loc)
        logits = (z[plan.pu] * z[plan.pv]).sum(-1)
        ndf = plan.n_df
        r_dec = logits[:ndf] - logits[ndf:2 * ndf]
        nu, nv = plan.ni_pairs()
        r_ni = logits[2 * ndf:] - (zo[nu] * zo[nv]).sum(-1)
        loss_r = (r_dec[:plan.own_df] ** 2).sum() / plan.norm_df
        loss_l = (r_ni[:plan.own_ni] ** 2).sum() / plan.norm_ni
        # the gradient w.r.t. the LOCAL rows needs every touching pair, counted or not: build the full local objective
        obj = 0.5 * (r_dec ** 2).sum() / plan.norm_df + 0.5 * (r_ni ** 2).sum() / plan.norm_ni
        # d obj / d z_loc uses only the local rows' dependence (remote rows are constants from the gather)
        gz = torch.autograd.grad(obj, z_loc, retain_graph=True)[0]
        # incidence formulation the CUDA engine uses
        coef = torch.cat([r_dec, -r_dec]) * (0.5 * 2 / plan.norm_df)
        coef = torch.cat([coef, r_ni * (0.5 * 2 / plan.norm_ni)]).detach()
        gz_inc = torch.zeros_like(z_loc).index_add(0, plan.ent_node_loc, coef[plan.ent_pair].view(-1, 1) * z.detach()[plan.ent_partner])
        assert torch.allclose(gz, gz_inc, rtol=1e-9, atol=1e-12)
        # backward through the local layers with the halo exchange of dA2 done explicitly
        dW2 = a2[plan.rows2_loc].t().detach() @ gz_inc[plan.rows2_loc]
        da2 = gz_inc.clone(); da2[plan.rows2_loc] = gz_inc[plan.rows2_loc] @ D2.detach().t()
        da2_full = gather(dinv.view(-1, 1) * da2).detach()
        dh1 = dinv.view(-1, 1) * agg(da2_full)           # A_hat symmetric: same local rows serve the transpose
        dx1 = (dh1 @ W2) * (x1.detach() > 0)
        dW1 = a1[plan.rows1_loc].t().detach() @ dx1[plan.rows1_loc]
        red = torch.cat([dW1.flatten(), dW2.flatten(), torch.stack([0.5 * loss_r + 0.5 * loss_l, loss_r, loss_l]).detach()])
        dist.all_reduce(red)
        n1 = dW1.numel()
        ok = torch.allclose(red[:n1].view_as(dW1), om.deletion1.deletion_weight.grad, rtol=1e-8, atol=1e-12) and \
            torch.allclose(red[n1:n1 + dW2.numel()].view_as(dW2), om.deletion2.deletion_weight.grad, rtol=1e-8, atol=1e-12) and \
            torch.allclose(red[-3:], torch.stack([loss_o, lr_o, ll_o]).detach(), rtol=1e-9)
        # every pair is counted exactly once across ranks
        cnt = torch.tensor([plan.own_df, plan.own_ni])
        dist.all_reduce(cnt)
        ok = ok and cnt.tolist() == [plan.norm_df, plan.norm_ni]
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_row_partition_matches_oracle_world2():
    world = 2
    ctx = mp.get_context('spawn')
    ret = ctx.Manager().dict()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert dict(ret) == {0: True, 1: True}


def test_row_bounds_cover_all_rows():
    from gnndelete_b200.dist import row_bounds
    for n, w in [(10, 3), (235368, 8), (7, 8), (16, 4)]:
        b = row_bounds(n, w)
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert all(hi - lo <= b[0][1] - b[0][0] for lo, hi in b)
