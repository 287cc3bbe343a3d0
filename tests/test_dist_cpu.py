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


def _worker(rank, world, port, ret, balance):
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

        plan = PartitionPlan(data, neg, rank, world, balance=balance)
        n, nl, lo, hi, per = plan.n, plan.n_loc, plan.lo, plan.hi, plan.per
        W1, b1 = om.conv1.lin.weight.detach(), om.conv1.bias.detach()
        W2, b2 = om.conv2.lin.weight.detach(), om.conv2.bias.detach()
        D1 = om.deletion1.deletion_weight.detach().clone()
        D2 = om.deletion2.deletion_weight.detach().clone()

        def gather(loc):
            """local rows -> the [world * per, F] slot-space matrix every rank sees"""
            pad = torch.zeros(per, loc.shape[1], dtype=loc.dtype)
            pad[:nl] = loc
            parts = [torch.zeros_like(pad) for _ in range(world)]
            dist.all_gather(parts, pad)
            return torch.cat(parts)

        deg = torch.zeros(nl, dtype=torch.float64).index_add_(0, plan.mp_dst_loc, torch.ones(plan.mp_src_slot.numel(), dtype=torch.float64))
        dinv = deg.pow(-0.5)

        def agg(h_slots):
            out = torch.zeros(nl, h_slots.shape[1], dtype=torch.float64)
            return out.index_add(0, plan.mp_dst_loc, h_slots[plan.mp_src_slot])

        x_loc = d64.x[lo:hi]
        a1 = dinv.view(-1, 1) * agg(gather(dinv.view(-1, 1) * (x_loc @ W1.t()))) + b1
        x1 = a1.clone(); x1[plan.rows1_loc] = a1[plan.rows1_loc] @ D1
        a2 = dinv.view(-1, 1) * agg(gather(dinv.view(-1, 1) * (x1.relu() @ W2.t()))) + b2
        z_loc = a2.clone(); z_loc[plan.rows2_loc] = a2[plan.rows2_loc] @ D2
        zg = gather(z_loc)
        # DEC residuals of this rank's share of the Df items -> coefficients, all-gathered as [world, 2, per_items]
        ni_, pi = plan.n_items, plan.per_items
        lg = (zg[plan.dec_pu_slot] * zg[plan.dec_pv_slot]).sum(-1)
        r_dec = lg[:ni_] - lg[ni_:]
        c_r = 0.5 * 2 / plan.norm_df
        send = torch.zeros(2, pi, dtype=torch.float64)
        send[0, :ni_] = c_r * r_dec; send[1, :ni_] = -c_r * r_dec
        parts = [torch.zeros_like(send) for _ in range(world)]
        dist.all_gather(parts, send)
        coef_all = torch.stack(parts).reshape(-1)
        # node pass: every incident pair of a local node, NI logits recomputed from the node's side
        zo_slots = gather(zo[lo:hi])
        tgt = (zo[lo + plan.ni_node_loc] * zo_slots[plan.ni_partner_slot]).sum(-1)
        r_ni = (z_loc[plan.ni_node_loc] * zg[plan.ni_partner_slot]).sum(-1) - tgt
        c_l = 0.5 * 2 / plan.norm_ni
        dz = torch.zeros_like(z_loc)
        dz.index_add_(0, plan.ni_node_loc, (c_l * r_ni).view(-1, 1) * zg[plan.ni_partner_slot])
        dz.index_add_(0, plan.dec_node_loc, coef_all[plan.dec_coef_idx].view(-1, 1) * zg[plan.dec_partner_slot])
        loss_r_part = (r_dec ** 2).sum() / plan.norm_df
        ni_sq = (r_ni ** 2).sum()
        # backward through the local layers with the halo exchange of D^-1/2 dA2 done explicitly
        dW2 = a2[plan.rows2_loc].t() @ dz[plan.rows2_loc]
        da2 = dz.clone(); da2[plan.rows2_loc] = dz[plan.rows2_loc] @ D2.t()
        dh1 = dinv.view(-1, 1) * agg(gather(dinv.view(-1, 1) * da2))     # A_hat symmetric: same local rows serve the transpose
        dx1 = (dh1 @ W2) * (x1 > 0)
        dW1 = a1[plan.rows1_loc].t() @ dx1[plan.rows1_loc]
        red = torch.cat([dW1.flatten(), dW2.flatten(), torch.stack([loss_r_part, ni_sq])])
        dist.all_reduce(red)
        n1 = dW1.numel()
        loss_r, loss_l = red[-2], red[-1] * 0.5 / plan.norm_ni
        ok = torch.allclose(red[:n1].view_as(dW1), om.deletion1.deletion_weight.grad, rtol=1e-8, atol=1e-12) and \
            torch.allclose(red[n1:n1 + dW2.numel()].view_as(dW2), om.deletion2.deletion_weight.grad, rtol=1e-8, atol=1e-12) and \
            torch.allclose(torch.stack([0.5 * loss_r + 0.5 * loss_l, loss_r, loss_l]), torch.stack([loss_o, lr_o, ll_o]).detach(), rtol=1e-9)
        # every Df item belongs to exactly one rank; every aggregation / incidence entry to exactly one rank
        cnt = torch.tensor([plan.n_items, plan.ni_node_loc.numel(), plan.dec_node_loc.numel(), plan.mp_dst_loc.numel(), nl])
        dist.all_reduce(cnt)
        sdf = data.train_pos_edge_index[:, data.sdf_mask]
        want = [plan.norm_df, 2 * plan.norm_ni, 4 * plan.norm_df, int((sdf[0] != sdf[1]).sum()) + n, n]
        ok = ok and cnt.tolist() == want
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('balance', [False, True])
def test_row_partition_matches_oracle_world2(balance):
    world = 2
    ctx = mp.get_context('spawn')
    ret = ctx.Manager().dict()
    port = 29500 + (os.getpid() % 2000) + int(balance)
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret, balance)) for r in range(world)]
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


def test_balanced_bounds_equalise_work():
    from gnndelete_b200.dist import balanced_bounds
    g = torch.Generator().manual_seed(0)
    w = (torch.arange(1, 5001, dtype=torch.float32) ** -0.5) * 100 + torch.rand(5000, generator=g)     # power-law row work
    for world in (1, 2, 4, 8):
        b = balanced_bounds(w, world)
        assert b[0][0] == 0 and b[-1][1] == 5000 and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        tot = [float(w[lo:hi].sum()) for lo, hi in b]
        assert max(tot) <= 1.02 * sum(tot) / world + float(w.max())
