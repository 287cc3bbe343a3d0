"""GPU parity of the fused epoch engine (eager and CUDA-graph) against the oracle over several epochs."""
import pytest
import torch

from tests import util as U

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _oracle_run(shape, data, neg, epochs, dtype):
    from oracle import unlearn as OU
    om = U.oracle_model('gcn', shape, data, dtype=dtype)
    d = data.clone()
    d.x = data.x.to(dtype)
    with torch.no_grad():
        zo = om.get_original_embeddings(d.x, d.train_pos_edge_index[:, d.dr_mask])
    opt = torch.optim.Adam([p for n, p in om.named_parameters() if 'del' in n], lr=1e-3)
    hist = []
    for _ in range(epochs):
        loss, lr, ll, _ = OU.edge_form_loss(om, d, neg, zo)
        loss.backward()
        opt.step(); opt.zero_grad()
        hist.append([float(loss), float(lr), float(ll)])
    return om, zo, torch.tensor(hist, dtype=torch.float64)


@pytest.mark.parametrize('graph,hoist,static', [(False, False, False), (True, False, False), (True, True, False),
                                                (True, False, True)])
def test_engine_loss_curve(lib, graph, hoist, static):
    from gnndelete_b200 import models as M
    from gnndelete_b200.engine import GCNDeleteEngine
    shape, raw, df, data, neg = U.make_case('cora', 0.05)
    epochs = 10
    om, zo, hist = _oracle_run(shape, data, neg, epochs, torch.float64)
    init = U.oracle_model('gcn', shape, data, dtype=torch.float32)       # same seed -> same initial weights
    m = M.GCNDelete(U.args_for(shape), data.sdf_node_1hop_mask, data.sdf_node_2hop_mask)
    m.load_state_dict(init.state_dict())
    m = m.to(DEV)
    eng = GCNDeleteEngine(m, data.clone().to(DEV), neg.to(DEV), z_ori=zo.float().to(DEV), hoist_layer1=hoist,
                          static_negatives=static)
    if graph:
        eng.capture()
    got = []
    for _ in range(epochs):
        got.append(eng.epoch().clone())
    got = torch.stack(got).cpu().double()
    # loss curve over 10 Adam steps; tolerance widened to 1e-4 because ten
    # optimizer steps compound the fp32 rounding of the gradients
    U.assert_close(got, hist, tol=1e-4, what='loss curve')
    U.assert_close(m.deletion1.deletion_weight, om.deletion1.deletion_weight, tol=1e-4, what='W_del1 after training')
    U.assert_close(m.deletion2.deletion_weight, om.deletion2.deletion_weight, tol=1e-4, what='W_del2 after training')


@pytest.mark.parametrize('static', [False, True])
def test_engine_new_negatives_every_epoch(lib, static):
    """The reference draws new negatives every epoch (gnndelete.py:221-225): after set_negatives the losses
    equal those of an engine built on the new negatives, in both incidence layouts and through the
    in-graph rebuild."""
    from gnndelete_b200 import models as M
    from gnndelete_b200 import synthetic as S
    from gnndelete_b200.engine import GCNDeleteEngine
    shape, raw, df, data, neg = U.make_case('cora', 0.05)
    om = U.oracle_model('gcn', shape, data, dtype=torch.float32)
    with torch.no_grad():
        zo = om.get_original_embeddings(data.x, data.train_pos_edge_index[:, data.dr_mask])
    neg2 = S.supplied_negatives(shape.num_nodes, neg.shape[1], seed=777)

    def fresh(ng, **kw):
        m = M.GCNDelete(U.args_for(shape), data.sdf_node_1hop_mask, data.sdf_node_2hop_mask)
        m.load_state_dict(om.state_dict())
        return GCNDeleteEngine(m.to(DEV), data.clone().to(DEV), ng.to(DEV), z_ori=zo.to(DEV), hoist_layer1=False, **kw)

    want = fresh(neg2).forward_backward().clone()
    eng = fresh(neg, static_negatives=static)
    eng.forward_backward()
    eng.set_negatives(neg2.to(DEV))
    U.assert_close(eng.forward_backward(), want, what='losses after set_negatives')
    g_want = fresh(neg2)
    g_want.forward_backward()
    U.assert_close(eng.params[0].grad, g_want.params[0].grad, what='dW_del1 after set_negatives')
    if not static:
        eng2 = fresh(neg)
        eng2.capture(dynamic_negatives=True)
        eng2.set_negatives(neg2.to(DEV))
        U.assert_close(eng2.epoch(), want, what='losses through the in-graph negative rebuild')
        # the double-buffered pipeline returns every step's losses, one step late, and matches the synchronous path
        from gnndelete_b200.engine import EpochPipeline
        eng3, eng4 = fresh(neg, deterministic=True), fresh(neg, deterministic=True)      # sorted per-step incidence: bitwise
        eng3.capture(dynamic_negatives=True)
        eng4.capture(dynamic_negatives=True)
        pipe = EpochPipeline(eng3)
        hosts = [neg.pin_memory(), neg2.pin_memory(), neg.pin_memory(), neg2.pin_memory()]
        got, last = [], None
        for h in hosts:
            k = pipe.submit(h)
            if last is not None:
                got.append(pipe.result(last).clone())
            last = k
        got.append(pipe.result(last).clone())
        for h, g_ in zip(hosts, got):
            eng4.set_negatives(h.to(DEV))
            assert torch.equal(eng4.epoch().cpu(), g_), 'pipelined step == synchronous step, bitwise'
        # default mode (negative-pair gradient added with float reductions): same steps within fp32 rounding
        eng5, eng6 = fresh(neg), fresh(neg)
        assert eng5.loss.neg_atomic
        eng5.capture(dynamic_negatives=True)
        eng6.capture(dynamic_negatives=True)
        pipe = EpochPipeline(eng5)
        for h in hosts:
            g_ = pipe.result(pipe.submit(h)).clone()
            eng6.set_negatives(h.to(DEV))
            U.assert_close(eng6.epoch().cpu(), g_, tol=1e-5, what='pipelined vs synchronous step (float reductions)')


def test_engine_chained_dx_dw_kernel(lib, monkeypatch):
    """GD_FUSED_DXDW=1: dX1 chained into dW_del1 through tensor memory -- same loss curve and weights as the oracle, and
    the first-step gradient within fp32 rounding of the two-kernel engine's."""
    from gnndelete_b200 import models as M
    from gnndelete_b200.engine import GCNDeleteEngine
    shape, raw, df, data, neg = U.make_case('cora', 0.05)
    epochs = 10
    om, zo, hist = _oracle_run(shape, data, neg, epochs, torch.float64)
    init = U.oracle_model('gcn', shape, data, dtype=torch.float32)

    def engine(flag):
        monkeypatch.setenv('GD_FUSED_DXDW', flag)
        m = M.GCNDelete(U.args_for(shape), data.sdf_node_1hop_mask, data.sdf_node_2hop_mask)
        m.load_state_dict(init.state_dict())
        m = m.to(DEV)
        return m, GCNDeleteEngine(m, data.clone().to(DEV), neg.to(DEV), z_ori=zo.float().to(DEV))

    m0, e0 = engine('0')
    m1, e1 = engine('1')
    assert e1.fused_dxdw and not e0.fused_dxdw
    e0.forward_backward(); e1.forward_backward()
    U.assert_close(e1.params[0].grad, e0.params[0].grad, tol=1e-5, what='dW_del1, chained vs two kernels')
    e1.capture()
    got = torch.stack([e1.epoch().clone() for _ in range(epochs)]).cpu().double()
    U.assert_close(got, hist, tol=1e-4, what='loss curve')
    U.assert_close(m1.deletion1.deletion_weight, om.deletion1.deletion_weight, tol=1e-4, what='W_del1 after training')


def test_engine_first_step_tight(lib):
    """One epoch at the 1e-5 bar: losses and both Del gradients."""
    import __graft_entry__ as G
    G.smoke()


def test_dense_block_ni_matches_fullbatch_oracle(lib):
    """train_fullbatch's dense NI (gnndelete.py:163-193, 239-241): losses and Del gradients against the
    oracle's N x N formulation; the CUDA path only ever touches the S2 x S2 block."""
    dense_ni_case(0.03)


@pytest.mark.parametrize('n_s', [37, 128, 700, 1500, 2600])
def test_dense_ni_tensor_core_kernel_matches_fp64(lib, n_s, monkeypatch):
    """gd_dense_ni_tc_fwd_bwd (tcgen05, 3xTF32) on its own: loss sum and dz against an fp64 evaluation of
    gnndelete.py:239-241 on the S x S block, and against the fp32 CUDA-core kernel.  Sizes cover a partial row block,
    one exact block, a split J sweep (jsplit > 1), ragged last blocks and (2600: six column blocks per CTA) the steady state
    of the block pipeline; excluded (Df) pairs in both orders."""
    from gnndelete_b200.losses import DenseNIPlan
    torch.manual_seed(n_s)
    n = n_s + 50
    mask = torch.zeros(n, dtype=torch.bool)
    mask[torch.randperm(n)[:n_s]] = True
    S = mask.nonzero().squeeze(1)
    z_ori = torch.randn(n, 64, dtype=torch.float64) * 0.4
    z = z_ori + 0.1 * torch.randn(n, 64, dtype=torch.float64)
    logits_ori = z_ori @ z_ori.t()
    df = torch.stack([S[torch.randint(0, n_s, (3 * n_s,))], S[torch.randint(0, n_s, (3 * n_s,))]])
    # fp64 reference on the block
    zs = z[S].clone().requires_grad_(True)
    keep = torch.ones(n_s, n_s, dtype=torch.bool).tril(-1)
    pos = torch.full((n,), -1, dtype=torch.long); pos[S] = torch.arange(n_s)
    pu, pv = pos[df[0]], pos[df[1]]
    keep[pu, pv] = False; keep[pv, pu] = False
    res = (torch.sigmoid(zs @ zs.t()) - torch.sigmoid(logits_ori[S][:, S]))[keep]
    loss_ref = (res ** 2).mean()
    (0.5 * loss_ref).backward()
    out = {}
    for mode in ('tc', 'simt'):
        monkeypatch.setenv('GD_DENSE_NI', mode)
        plan = DenseNIPlan(mask.to(DEV), df.to(DEV), logits_ori.float().to(DEV), n, 64, weight=0.5)
        assert plan.tensor_core == (mode == 'tc')
        assert plan.num_pairs == int(keep.sum())
        dz = torch.zeros(n, 64, device=DEV)
        loss = plan.forward_backward(z.float().to(DEV), dz)
        loss2 = plan.forward_backward(z.float().to(DEV), torch.zeros_like(dz))      # second call: same result (TMEM / barriers re-armed)
        assert torch.equal(loss, loss2)
        U.assert_close(loss.reshape(()), loss_ref.detach(), what=f'dense NI loss ({mode})')
        U.assert_close(dz[S.to(DEV)], zs.grad, what=f'dense NI dz ({mode})')
        assert float(dz[~mask.to(DEV)].abs().max()) == 0.0
        out[mode] = (loss, dz)
    U.assert_close(out['tc'][1], out['simt'][1], what='tensor-core vs CUDA-core dz')


def dense_ni_case(scale):
    from gnndelete_b200 import models as M
    from gnndelete_b200.engine import GCNDeleteEngine
    from oracle import unlearn as OU
    shape, raw, df, data, neg = U.make_case('cora', scale)
    om = U.oracle_model('gcn', shape, data, dtype=torch.float64)
    d64 = data.clone(); d64.x = data.x.double()
    with torch.no_grad():
        zo = om.get_original_embeddings(d64.x, d64.train_pos_edge_index[:, d64.dr_mask])
    logits_ori = zo @ zo.t()                                   # what base.py:288 stores in pred_proba.pt
    pm = OU.dense_pair_mask(d64, d64.sdf_node_2hop_mask)
    loss_o, lr_o, ll_o, _ = OU.fullbatch_loss(om, d64, neg, logits_ori, pm)
    loss_o.backward()
    m = M.GCNDelete(U.args_for(shape), data.sdf_node_1hop_mask, data.sdf_node_2hop_mask)
    m.load_state_dict({k: v.float() for k, v in om.state_dict().items()})
    m = m.to(DEV)
    eng = GCNDeleteEngine(m, data.clone().to(DEV), neg.to(DEV), logits_ori=logits_ori.float().to(DEV), hoist_layer1=False)
    assert eng.dense.num_pairs == int(pm.sum())
    losses = eng.forward_backward()
    U.assert_close(losses, torch.stack([loss_o, lr_o, ll_o]), what='dense NI losses')
    U.assert_close(m.deletion2.deletion_weight.grad, om.deletion2.deletion_weight.grad, what='dense NI dW2')
    U.assert_close(m.deletion1.deletion_weight.grad, om.deletion1.deletion_weight.grad, what='dense NI dW1')
    # CUDA-graph replay gives the same numbers
    eng.capture()
    l2 = eng.epoch().clone()
    U.assert_close(l2, torch.stack([loss_o, lr_o, ll_o]), what='dense NI losses (graph)')
