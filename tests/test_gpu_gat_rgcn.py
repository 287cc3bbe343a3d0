"""GPU parity: GATDelete (edge-softmax aggregation fwd/bwd) and RGCNDelete (relation-segmented
mean aggregation + block-diagonal weights fwd/bwd) against the CPU oracle in fp64."""
import pytest
import torch

from tests import util as U

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.mark.parametrize('in_dim', [128, 500])
def test_gat_delete_forward_and_grads(lib, in_dim):
    gat_delete_case(in_dim, 0.1)


def gat_delete_case(in_dim, scale):
    from gnndelete_b200 import models as M
    from oracle import unlearn as OU
    shape, raw, df, data, neg = U.make_case('pubmed', scale, in_dim=in_dim)
    om = U.oracle_model('gat', shape, data, dtype=torch.float64)
    d64 = data.clone()
    d64.x = data.x.double()
    with torch.no_grad():
        zo = om.get_original_embeddings(d64.x, d64.train_pos_edge_index[:, d64.dr_mask])
    loss_o, lr_o, ll_o, _ = OU.edge_form_loss(om, d64, neg, zo)
    loss_o.backward()
    z1_o, z2_o = om(d64.x, d64.train_pos_edge_index[:, d64.sdf_mask], return_all_emb=True)

    m = M.GATDelete(U.args_for(shape), data.sdf_node_1hop_mask, data.sdf_node_2hop_mask)
    m.load_state_dict({k: v.float() for k, v in om.state_dict().items()}, strict=True)
    m = m.to(DEV)
    dd = data.clone().to(DEV)
    ei = dd.train_pos_edge_index[:, dd.sdf_mask]
    z1, z2 = m(dd.x, ei, return_all_emb=True)
    U.assert_close(z1, z1_o, what='gat z1')
    U.assert_close(z2, z2_o, what='gat z2')
    zo_g = m.get_original_embeddings(dd.x, dd.train_pos_edge_index[:, dd.dr_mask])
    U.assert_close(zo_g, zo, what='gat z_ori')
    # reference-shaped loss through autograd
    negd = neg.to(DEV)
    n = int(dd.df_mask.sum())
    logits = m.decode(z2, dd.train_pos_edge_index[:, dd.df_mask], negd)
    loss_r = torch.nn.functional.mse_loss(logits[:n], logits[n:])
    lower = ei[0] < ei[1]
    pairs = torch.stack([ei[0][lower], ei[1][lower]])
    loss_l = torch.nn.functional.mse_loss(m.decode(z2, pairs), m.decode(zo_g, pairs).detach())
    (0.5 * loss_r + 0.5 * loss_l).backward()
    U.assert_close(loss_r, lr_o, what='gat loss_r')
    U.assert_close(loss_l, ll_o, what='gat loss_l')
    U.assert_close(m.deletion2.deletion_weight.grad, om.deletion2.deletion_weight.grad, what='gat dW_del2')
    U.assert_close(m.deletion1.deletion_weight.grad, om.deletion1.deletion_weight.grad, what='gat dW_del1')


def test_gat_isolated_and_hub_rows(lib):
    """Rows with only the self loop and a hub row much longer than a warp."""
    from gnndelete_b200 import models as M
    from oracle import models as OM
    import types
    n = 300
    g = torch.Generator().manual_seed(3)
    hub = torch.stack([torch.arange(1, 200), torch.zeros(199, dtype=torch.long)])      # 199 sources -> node 0
    rnd = torch.randint(200, 290, (2, 150), generator=g)                                # nodes 290.. isolated
    ei = torch.cat([hub, rnd], 1)
    args = types.SimpleNamespace(in_dim=64, hidden_dim=128, out_dim=64)
    torch.manual_seed(0)
    om = OM.GAT(args).double()
    U.randomize(om)
    x = torch.randn(n, 64, generator=g)
    ref1, ref2 = om(x.double(), ei, return_all_emb=True)
    m = M.GAT(args)
    m.load_state_dict({k: v.float() for k, v in om.state_dict().items()})
    m = m.to(DEV)
    z1, z2 = m(x.to(DEV), ei.to(DEV), return_all_emb=True)
    U.assert_close(z1, ref1, what='gat hub z1')
    U.assert_close(z2, ref2, what='gat hub z2')


def test_gat_hub_row_10k_neighbours(lib):
    """One destination with 10,000 in-neighbours (a row 300x longer than a sub-warp batch) next to ordinary rows:
    edge-softmax aggregation forward and its gradient w.r.t. the input features against the oracle."""
    from gnndelete_b200 import models as M
    from oracle import models as OM
    import types
    n = 12000
    g = torch.Generator().manual_seed(11)
    hub = torch.stack([torch.arange(1, 10001), torch.zeros(10000, dtype=torch.long)])
    rnd = torch.randint(0, n, (2, 30000), generator=g)
    ei = torch.unique(torch.cat([hub, rnd[:, rnd[0] != rnd[1]]], 1), dim=1)
    args = types.SimpleNamespace(in_dim=64, hidden_dim=128, out_dim=64)
    torch.manual_seed(0)
    om = OM.GAT(args).double()
    U.randomize(om)
    x = torch.randn(n, 64, generator=g)
    xo = x.double().requires_grad_(True)
    ref1, ref2 = om(xo, ei, return_all_emb=True)
    w = torch.randn(n, 64, generator=g)
    (ref2 * w.double()).sum().backward()
    m = M.GAT(args)
    m.load_state_dict({k: v.float() for k, v in om.state_dict().items()})
    m = m.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    z1, z2 = m(xg, ei.to(DEV), return_all_emb=True)
    U.assert_close(z1, ref1, what='gat 10k-hub z1')
    U.assert_close(z2, ref2, what='gat 10k-hub z2')
    (z2 * w.to(DEV)).sum().backward()
    U.assert_close(xg.grad, xo.grad, what='gat 10k-hub dx')


def test_gat_delete_engine_first_step(lib):
    """GATDeleteEngine (fused, capturable GATDelete epoch): losses and both Del gradients of the first step against the
    fp64 oracle, then the captured graph against eager steps."""
    from gnndelete_b200 import models as M
    from gnndelete_b200.engine import GATDeleteEngine
    from oracle import unlearn as OU
    shape, raw, df, data, neg = U.make_case('pubmed', 0.2, in_dim=500)
    om = U.oracle_model('gat', shape, data, dtype=torch.float64)
    d64 = data.clone(); d64.x = data.x.double()
    with torch.no_grad():
        zo = om.get_original_embeddings(d64.x, d64.train_pos_edge_index[:, d64.dr_mask])
    loss, lr, ll, _ = OU.edge_form_loss(om, d64, neg, zo)
    loss.backward()

    def fresh(**kw):
        m = M.GATDelete(U.args_for(shape), data.sdf_node_1hop_mask, data.sdf_node_2hop_mask)
        m.load_state_dict({k: v.float() for k, v in om.state_dict().items()}, strict=True)
        return GATDeleteEngine(m.to(DEV), data.clone().to(DEV), neg.to(DEV), z_ori=zo.float().to(DEV), **kw)

    for hoist in (False, True):
        eng = fresh(hoist_layer1=hoist, static_negatives=True)
        got = eng.forward_backward().clone()
        U.assert_close(got, torch.stack([loss, lr, ll]).detach(), what=f'gat engine losses (hoist={hoist})')
        U.assert_close(eng.params[0].grad, om.deletion1.deletion_weight.grad, what='gat engine dW_del1')
        U.assert_close(eng.params[1].grad, om.deletion2.deletion_weight.grad, what='gat engine dW_del2')
    a, b = fresh(), fresh()
    a.capture(dynamic_negatives=True)
    for _ in range(3):
        U.assert_close(a.epoch(), b.epoch(), tol=1e-5, what='captured vs eager GAT epochs')


@pytest.mark.parametrize('num_edge_type,mode', [(51, 'edge'), (51, 'transform'), (51, 'tile'), (9, 'edge'), (9, 'tile')])
def test_rgcn_delete_forward_and_grads(lib, num_edge_type, mode, monkeypatch):
    """num_edge_type 51 -> block-diagonal weights (num_blocks=4), 9 -> dense relation weights; every RGCN execution
    path (GD_RGCN): 'edge' (relation-sorted edge tiles, the default for block weights; dense weights fall through to
    transform-then-gather), 'transform', 'tile' (generic relation-tile kernel)."""
    from gnndelete_b200 import ops
    monkeypatch.setattr(ops, 'RGCN_MODE', mode)
    rgcn_delete_case(num_edge_type, 0.002)


def test_rgcn_edge_path_hub_rows_and_chunks(lib, monkeypatch):
    """Edge path with work items much shorter than a tile's edge list (hub rows spread over several chunks) against
    the generic tile kernel, forward and transposed, both layer shapes."""
    import dataclasses
    from gnndelete_b200 import graph as G
    from gnndelete_b200 import ops
    from gnndelete_b200 import synthetic as S
    shape = dataclasses.replace(S.SHAPES['biokg'].scaled(0.004), num_edge_type=51)
    raw = S.make_graph(shape, seed=1)
    ei = torch.cat([raw.train_pos_edge_index, raw.train_pos_edge_index.flip(0)], 1).to(DEV)
    et = torch.cat([raw.train_edge_type, raw.train_edge_type + 51]).to(DEV)
    n = shape.num_nodes
    g = torch.Generator().manual_seed(2)
    for (fin, fout) in ((128, 64), (128, 128)):
        w = (torch.randn(102, 4, fin // 4, fout // 4, generator=g) * 0.2).to(DEV)
        root = (torch.randn(fin, fout, generator=g) * 0.1).to(DEV)
        bias = torch.randn(fout, generator=g).to(DEV)
        x = torch.randn(n, fin, generator=g).to(DEV)
        gout = torch.randn(n, fout, generator=g).to(DEV)
        plan = G.GraphPlan(ei, n, False, et, 102)
        monkeypatch.setattr(ops, 'RGCN_MODE', 'tile')
        ref = ops.rgcn_conv(plan, x, w, root, bias)
        ref_t = ops.rgcn_conv(plan, gout, w, root, None, transposed=True)
        monkeypatch.setattr(ops, 'RGCN_MODE', 'edge')
        monkeypatch.setattr(ops, 'RGCN_EDGE_CHUNK', 64)
        got = ops.rgcn_conv(plan, x, w, root, bias)
        got_t = ops.rgcn_conv(plan, gout, w, root, None, transposed=True)
        assert '_rgcn_edge' in plan.__dict__, 'the edge path must have been taken'
        U.assert_close(got, ref, what=f'edge vs tile {fin}->{fout}')
        U.assert_close(got_t, ref_t, what=f'edge vs tile transposed {fin}->{fout}')
        assert torch.equal(ops.rgcn_conv(plan, x, w, root, bias), got), 'bitwise reproducible'


def rgcn_delete_case(num_edge_type, scale):
    import dataclasses
    from gnndelete_b200 import models as M
    from gnndelete_b200 import synthetic as S
    from oracle import unlearn as OU
    shape = dataclasses.replace(S.SHAPES['biokg'].scaled(scale), num_edge_type=num_edge_type)
    raw = S.make_graph(shape, seed=42)
    df = S.sample_df_mask(shape.num_edges, shape.num_deleted, seed=42)
    data = OU.build_unlearning_data(raw, df, num_edge_type=num_edge_type)
    om = U.oracle_model('rgcn', shape, data, dtype=torch.float64, num_nodes=shape.num_nodes,
                        num_edge_type=num_edge_type)
    gen = torch.Generator().manual_seed(7)
    pos_ei = data.edge_index[:, data.df_mask]
    pos_et = data.edge_type[data.df_mask]
    dec = pos_et < num_edge_type
    neg = OU.negative_sampling_kg(pos_ei[:, dec], pos_et[dec], generator=gen)
    loss1, loss2, parts = OU.kg_step_losses(om, data, neg, num_edge_type, alpha=0.5)
    (loss1 + loss2).backward()

    m = M.RGCNDelete(U.args_for(shape), shape.num_nodes, num_edge_type, data.sdf_node_1hop_mask,
                     data.sdf_node_2hop_mask)
    m.load_state_dict({k: v.float() for k, v in om.state_dict().items()}, strict=True)
    m = m.to(DEV)
    dd = data.clone().to(DEV)
    m1, m2 = OU.kg_non_df_masks(data)
    ei, et = dd.edge_index[:, dd.dr_mask], dd.edge_type[dd.dr_mask]
    z1, z2 = m(dd.x, ei, et, m1.to(DEV), m2.to(DEV), return_all_emb=True)
    U.assert_close(z1, parts['z1'], what='rgcn z1')
    U.assert_close(z2, parts['z2'], what='rgcn z2')
    # pre-activations within rounding distance of zero: fp32 and fp64 may disagree on the ReLU mask there, and the gradient
    # jumps by a finite amount.  Such elements must be at the noise level; the oracle's gradients are then taken with the
    # mask of the implementation under test (the forward values differ by < 1e-5 either way, checked above).
    zo1 = parts['z1'].detach()
    flips = (z1.detach().cpu() > 0) != (zo1 > 0)
    if bool(flips.any()):
        assert float(zo1[flips].abs().max()) <= 1e-5 * float(zo1.abs().max()), 'ReLU masks differ away from zero'
        om.zero_grad()
        om.relu_mask_override = (z1.detach().cpu() > 0)
        loss1, loss2, parts = OU.kg_step_losses(om, data, neg, num_edge_type, alpha=0.5)
        (loss1 + loss2).backward()
    with torch.no_grad():
        z1o, z2o = m.get_original_embeddings(dd.x, ei, et, return_all_emb=True)
    # node-embedding losses (gnndelete_nodeemb.py:770-798) written with torch ops on the CUDA outputs
    F = torch.nn.functional
    dei, ngd = parts['decoding_edge_index'].to(DEV), neg.to(DEV)
    m1d, m2d = m1.to(DEV), m2.to(DEV)
    l1 = 0.5 * F.mse_loss(torch.cat([z1[dei[0]], z1[dei[1]]]), torch.cat([z1o[ngd[0]], z1o[ngd[1]]])) + \
        0.5 * F.mse_loss(z1[m1d], z1o[m1d])
    l2 = 0.5 * F.mse_loss(torch.cat([z2[dei[0]], z2[dei[1]]]), torch.cat([z2o[ngd[0]], z2o[ngd[1]]])) + \
        0.5 * F.mse_loss(z2[m2d], z2o[m2d])
    U.assert_close(l1, loss1, what='rgcn loss1')
    U.assert_close(l2, loss2, what='rgcn loss2')
    (l1 + l2).backward()
    U.assert_close(m.deletion1.deletion_weight.grad, om.deletion1.deletion_weight.grad, what='rgcn dW_del1')
    U.assert_close(m.deletion2.deletion_weight.grad, om.deletion2.deletion_weight.grad, what='rgcn dW_del2')
    # DistMult decoder
    lg = m.decode(z2.detach(), dd.directed_df_edge_index, dd.directed_df_edge_type)
    ref = om.decode(parts['z2'].detach(), data.directed_df_edge_index, data.directed_df_edge_type)
    U.assert_close(lg, ref, what='distmult logits')
