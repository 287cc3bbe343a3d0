"""CPU: host logic of the node-embedding loss path (`losses.row_mse_incidence`, the incidence
`gd_row_mse_fwd_bwd` walks) and the oracle's `nodeemb_epoch` restatement of
`GNNDeleteNodeembTrainer.train_fullbatch` (gnndelete_nodeemb.py:191-299)."""
import os

import numpy as np
import pytest
import torch

from tests import util as U

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'oracle_nodeemb_small.npz')


def walk_incidence(z, zo, rowptr, code, w, mix):
    """What the kernel computes, entry by entry in incidence order (float64)."""
    n = z.shape[0]
    s = [0.0, 0.0]
    dz = np.zeros_like(z)
    for r in range(n):
        for e in range(rowptr[r], rowptr[r + 1]):
            c = int(code[e])
            t = 1 if c < 0 else 0
            src = -1 - c if c < 0 else c
            d = z[r] - zo[src]
            s[t] += float((d * d).sum())
            dz[r] += 2 * mix[t] * w[t] * d
    l0, l1 = w[0] * s[0], w[1] * s[1]
    return np.array([mix[0] * l0 + mix[1] * l1, l0, l1]), dz


@pytest.mark.parametrize('reduction', ['mean', 'sum'])
def test_incidence_walk_equals_autograd_mse(reduction):
    from gnndelete_b200.losses import row_mse_incidence
    g = torch.Generator().manual_seed(5)
    n, f, m = 60, 12, 90
    z = torch.randn(n, f, generator=g, dtype=torch.float64, requires_grad=True)
    zo = torch.randn(n, f, generator=g, dtype=torch.float64)
    pos = torch.randint(0, 20, (2, m), generator=g)            # few distinct rows: long, repeated incidence rows
    neg = torch.randint(0, n, (2, m), generator=g)
    mask = torch.rand(n, generator=g) < 0.4
    mix = (0.3, 0.7)
    rowptr, code = row_mse_incidence(torch.cat([pos[0], pos[1]]), torch.cat([neg[0], neg[1]]), mask.nonzero().squeeze(1), n)
    assert rowptr.dtype == torch.int32 and code.dtype == torch.int32
    assert int(rowptr[0]) == 0 and int(rowptr[-1]) == code.numel() == 2 * m + int(mask.sum())
    assert bool((rowptr[1:] >= rowptr[:-1]).all())
    fct = torch.nn.MSELoss(reduction=reduction)
    loss_r = fct(torch.cat([z[pos[0]], z[pos[1]]]), torch.cat([zo[neg[0]], zo[neg[1]]]))
    loss_l = fct(z[mask], zo[mask])
    (mix[0] * loss_r + mix[1] * loss_l).backward()
    w = (1 / (2 * m * f), 1 / (int(mask.sum()) * f)) if reduction == 'mean' else (1.0, 1.0)
    losses, dz = walk_incidence(z.detach().numpy(), zo.numpy(), rowptr.numpy(), code.numpy(), w, mix)
    np.testing.assert_allclose(losses[1:], [float(loss_r), float(loss_l)], rtol=1e-12)
    np.testing.assert_allclose(dz, z.grad.numpy(), rtol=1e-10, atol=1e-14)


def test_incidence_is_stable_and_checked():
    from gnndelete_b200.losses import row_mse_incidence
    dst = torch.tensor([3, 1, 3, 3])
    src = torch.tensor([7, 8, 9, 0])
    rows = torch.tensor([3, 0])
    rowptr, code = row_mse_incidence(dst, src, rows, 10)
    assert rowptr.tolist() == [0, 1, 2, 2, 6, 6, 6, 6, 6, 6, 6]
    assert code.tolist() == [-1, 8, 7, 9, 0, -4]               # row 3: term-0 partners in input order, then its term-1 entry
    with pytest.raises(IndexError):
        row_mse_incidence(torch.tensor([10]), torch.tensor([0]), rows, 10)
    with pytest.raises(ValueError):
        row_mse_incidence(dst, src[:2], rows, 10)
    rowptr, code = row_mse_incidence(dst[:0], src[:0], rows[:0], 4)   # empty: all-zero rowptr
    assert rowptr.tolist() == [0] * 5 and code.numel() == 0


def _oracle_run(loss_type, loss_fct='mse_mean', epochs=3, alpha=0.4, dtype=torch.float64):
    from oracle import unlearn as OU
    shape, raw, df, data, neg = U.make_case('cora', 0.02)
    om = U.oracle_model('gcn', shape, data, dtype=dtype)
    # Reference defect: GCNDelete.forward leaves conv1 outside no_grad (deletion.py:62-63) and `conv.requires_grad =
    # False` is a no-op (:58-59), so deletion1's matmul saves W_del1 for the (unused) conv1 gradient and the
    # *layerwise schedules - optimizer[0].step() between the two backward passes - raise "modified by an inplace
    # operation" on any PyTorch >= 1.5 (see test_reference_gcn_layerwise_defect).  Freezing the conv parameters, as
    # the reference intended, changes no Del-weight arithmetic and lets the schedule run.
    for n, p in om.named_parameters():
        if 'del' not in n:
            p.requires_grad_(False)
    d = data.clone()
    d.x = data.x.to(dtype)
    with torch.no_grad():
        z1o, z2o = om.get_original_embeddings(d.x, d.train_pos_edge_index[:, d.dr_mask], return_all_emb=True)
    if 'layerwise' in loss_type:
        opt = [torch.optim.Adam(om.deletion1.parameters(), lr=1e-3), torch.optim.Adam(om.deletion2.parameters(), lr=1e-3)]
    else:
        opt = torch.optim.Adam([p for n, p in om.named_parameters() if 'del' in n], lr=1e-3)
    w0 = [om.deletion1.deletion_weight.detach().clone(), om.deletion2.deletion_weight.detach().clone()]
    hist = [torch.stack(OU.nodeemb_epoch(om, d, neg, z1o, z2o, opt, loss_type, alpha, loss_fct)) for _ in range(epochs)]
    return om, w0, torch.stack(hist)


def test_oracle_nodeemb_schedules():
    """The five loss_type branches differ in what they update and which gradients they clear."""
    om, w0, hist = _oracle_run('only1')
    assert not torch.equal(om.deletion1.deletion_weight, w0[0]) and torch.equal(om.deletion2.deletion_weight, w0[1])
    om, w0, hist = _oracle_run('only2_layerwise')
    assert torch.equal(om.deletion1.deletion_weight, w0[0]) and not torch.equal(om.deletion2.deletion_weight, w0[1])
    assert om.deletion1.deletion_weight.grad is not None           # loss2's deletion1 gradient stays in .grad (:271-277)
    om, w0, hist = _oracle_run('both_all')
    assert float(om.deletion1.deletion_weight.grad.abs().max()) > 0    # never zeroed (:219-229)
    om, w0, hist = _oracle_run('only2_all')
    assert not torch.equal(om.deletion1.deletion_weight, w0[0])    # one Adam over both Del weights: deletion1 moves too
    om_a, _, h_all = _oracle_run('both_all', epochs=1)
    om_l, _, h_lw = _oracle_run('both_layerwise', epochs=1)
    torch.testing.assert_close(h_all[0], h_lw[0])                  # same objective at the same weights
    assert bool(torch.isfinite(hist).all())


def test_reference_gcn_layerwise_defect():
    """As shipped (conv parameters left trainable), GCNDelete + both_layerwise cannot run: documents why the
    parity runs freeze the conv parameters."""
    from oracle import unlearn as OU
    shape, raw, df, data, neg = U.make_case('cora', 0.02)
    om = U.oracle_model('gcn', shape, data)
    with torch.no_grad():
        z1o, z2o = om.get_original_embeddings(data.x, data.train_pos_edge_index[:, data.dr_mask], return_all_emb=True)
    opt = [torch.optim.Adam(om.deletion1.parameters(), lr=1e-3), torch.optim.Adam(om.deletion2.parameters(), lr=1e-3)]
    with pytest.raises(RuntimeError, match='inplace operation'):
        OU.nodeemb_epoch(om, data, neg, z1o, z2o, opt, 'both_layerwise')


@pytest.mark.parametrize('loss_fct', ['mse_sum', 'kld_mean', 'kld_sum', 'cosine_mean', 'cosine_sum', 'linear_cka', 'rbf_cka'])
def test_oracle_nodeemb_loss_functions_run(loss_fct):
    om, w0, hist = _oracle_run('both_layerwise', loss_fct, epochs=2)
    assert bool(torch.isfinite(hist).all())
    assert not torch.equal(om.deletion2.deletion_weight, w0[1])


def test_oracle_nodeemb_against_golden():
    """The committed vectors (tests/golden/make_golden.py) pin the oracle's node-embedding epoch."""
    gold = np.load(GOLDEN)
    for lt in ('both_all', 'both_layerwise', 'only2_layerwise', 'only2_all', 'only1'):
        om, _, hist = _oracle_run(lt)
        np.testing.assert_allclose(hist.numpy(), gold[f'{lt}_hist'], rtol=1e-9)
        np.testing.assert_allclose(om.deletion1.deletion_weight.detach().numpy()[::4], gold[f'{lt}_W1'], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(om.deletion2.deletion_weight.detach().numpy()[::4], gold[f'{lt}_W2'], rtol=1e-9, atol=1e-12)
