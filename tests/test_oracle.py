"""CPU: pin the oracle as far as it can be pinned without PyG (PARITY UNPINNED, see oracle/__init__.py):
independent dense-matrix / networkx formulations of each restated operator, fp64 gradcheck of the
DeletionLayer, and the committed golden vectors."""
import os

import networkx as nx
import numpy as np
import pytest
import torch

from oracle import models as OM
from oracle import pyg_ops as P
from oracle import unlearn as OU
from tests import util as U

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'oracle_small.npz')


def _graph(n=40, e=120, seed=0):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, e), generator=g)
    ei = ei[:, ei[0] != ei[1]]
    ei = torch.unique(ei, dim=1)
    return n, ei, g


def test_gcn_conv_equals_dense_normalised_adjacency():
    n, ei, g = _graph()
    ei = torch.cat([ei, torch.tensor([[3], [3]])], 1)      # an explicit self loop must not be double counted
    x = torch.randn(n, 8, generator=g, dtype=torch.float64)
    w = torch.randn(5, 8, generator=g, dtype=torch.float64)
    b = torch.randn(5, generator=g, dtype=torch.float64)
    a = torch.zeros(n, n, dtype=torch.float64)
    keep = ei[0] != ei[1]
    a[ei[1][keep], ei[0][keep]] = 1
    a = a + torch.eye(n, dtype=torch.float64)
    d = a.sum(1).pow(-0.5)
    ref = (d.view(-1, 1) * a * d.view(1, -1)) @ (x @ w.t()) + b
    assert torch.allclose(P.gcn_conv(x, ei, w, b), ref, atol=1e-12)


def test_gat_conv_equals_dense_softmax():
    n, ei, g = _graph(seed=1)
    x = torch.randn(n, 8, generator=g, dtype=torch.float64)
    w = torch.randn(6, 8, generator=g, dtype=torch.float64)
    a_s = torch.randn(1, 1, 6, generator=g, dtype=torch.float64)
    a_d = torch.randn(1, 1, 6, generator=g, dtype=torch.float64)
    b = torch.randn(6, generator=g, dtype=torch.float64)
    h = x @ w.t()
    adj = torch.zeros(n, n, dtype=torch.bool)
    adj[ei[1], ei[0]] = True
    adj |= torch.eye(n, dtype=torch.bool)
    e = torch.nn.functional.leaky_relu((h * a_d.view(-1)).sum(-1).view(-1, 1) + (h * a_s.view(-1)).sum(-1).view(1, -1), 0.2)
    e = e.masked_fill(~adj, float('-inf'))
    ref = torch.softmax(e, dim=1) @ h + b
    assert torch.allclose(P.gat_conv(x, ei, w, a_s, a_d, b), ref, atol=1e-10)


def test_gin_conv_equals_dense():
    n, ei, g = _graph(seed=2)
    x = torch.randn(n, 8, generator=g, dtype=torch.float64)
    w = torch.randn(4, 8, generator=g, dtype=torch.float64)
    b = torch.randn(4, generator=g, dtype=torch.float64)
    a = torch.zeros(n, n, dtype=torch.float64)
    a.index_put_((ei[1], ei[0]), torch.ones(ei.shape[1], dtype=torch.float64), accumulate=True)
    assert torch.allclose(P.gin_conv(x, ei, w, b), (a @ x + x) @ w.t() + b, atol=1e-12)


@pytest.mark.parametrize('blocks', [None, 2])
def test_rgcn_conv_equals_dense(blocks):
    n, ei, g = _graph(seed=3)
    R, fin, fout = 4, 8, 6
    et = torch.randint(0, R, (ei.shape[1],), generator=g)
    x = torch.randn(n, fin, generator=g, dtype=torch.float64)
    if blocks is None:
        w = torch.randn(R, fin, fout, generator=g, dtype=torch.float64)
        dense = w
    else:
        w = torch.randn(R, blocks, fin // blocks, fout // blocks, generator=g, dtype=torch.float64)
        dense = torch.stack([torch.block_diag(*w[r]) for r in range(R)])
    root = torch.randn(fin, fout, generator=g, dtype=torch.float64)
    b = torch.randn(fout, generator=g, dtype=torch.float64)
    ref = x @ root + b
    for r in range(R):
        a = torch.zeros(n, n, dtype=torch.float64)
        sel = et == r
        a.index_put_((ei[1][sel], ei[0][sel]), torch.ones(int(sel.sum()), dtype=torch.float64), accumulate=True)
        ref = ref + (a / a.sum(1, keepdim=True).clamp(min=1)) @ x @ dense[r]
    assert torch.allclose(P.rgcn_conv(x, ei, et, w, root, b), ref, atol=1e-12)


def test_k_hop_subgraph_against_networkx_predecessors():
    """flow='source_to_target': a hop reaches the predecessors (sources of incoming edges)."""
    n, ei, g = _graph(n=60, e=150, seed=4)
    G = nx.DiGraph()
    G.add_nodes_from(range(n))
    G.add_edges_from(ei.t().tolist())
    seeds = [0, 7, 13]
    for hops in (1, 2, 3):
        frontier, subset = set(seeds), set(seeds)
        for _ in range(hops):
            frontier = {p for v in frontier for p in G.predecessors(v)}
            subset |= frontier
        s, sub_ei, inv, mask = P.k_hop_subgraph(torch.tensor(seeds), hops, ei, num_nodes=n)
        assert set(s.tolist()) == subset
        ref_mask = torch.tensor([(u in subset and v in subset) for u, v in ei.t().tolist()])
        assert torch.equal(mask, ref_mask)
        assert torch.equal(s[inv], torch.tensor(seeds))


def test_to_undirected_sorted_symmetric_and_carries_masks():
    n, ei, g = _graph(seed=5)
    lo = torch.stack([torch.minimum(ei[0], ei[1]), torch.maximum(ei[0], ei[1])])
    lo = torch.unique(lo, dim=1)
    a = (torch.rand(lo.shape[1], generator=g) < 0.3).int()
    sym, (a2,) = P.to_undirected(lo, [a])
    key = sym[0] * n + sym[1]
    assert torch.equal(key, torch.sort(key)[0]) and sym.shape[1] == 2 * lo.shape[1]
    assert P.is_undirected(sym, n) and not P.is_undirected(lo, n)
    look = {(int(u), int(v)): int(t) for (u, v), t in zip(lo.t().tolist(), a.tolist())}
    for (u, v), t in zip(sym.t().tolist(), a2.tolist()):
        assert t == look[(min(u, v), max(u, v))]


def test_deletion_layer_gradcheck_fp64():
    g = torch.Generator().manual_seed(6)
    mask = torch.rand(12, generator=g) < 0.5
    lay = OM.DeletionLayer(5, mask).double()
    with torch.no_grad():
        lay.deletion_weight.copy_(torch.randn(5, 5, generator=g, dtype=torch.float64))
    x = torch.randn(12, 5, generator=g, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(lambda xx, ww: torch.nn.functional.linear(
        torch.zeros(1, dtype=torch.float64), torch.zeros(1, 1, dtype=torch.float64)).sum() * 0 + _del(lay, xx, ww), (x, lay.deletion_weight))
    y = lay(x)
    assert torch.equal(y[~mask], x[~mask]) and y.data_ptr() != x.data_ptr()
    assert OM.DeletionLayer(5, None)(x) is x
    assert float(OM.DeletionLayer(4, mask).deletion_weight[0, 0]) == pytest.approx(1e-3)


def _del(lay, x, w):
    out = x.clone()
    out[lay.mask] = out[lay.mask] @ w
    return out


def test_mask_pipeline_invariants():
    """The asserts the reference itself carries (delete_gnn.py:147-148, 156, 182) hold on the oracle."""
    shape, raw, df, data, neg = U.make_case('cora', 0.05)
    assert not P.is_undirected(raw.train_pos_edge_index, raw.num_nodes)
    assert P.is_undirected(data.train_pos_edge_index, data.num_nodes)
    assert int(data.df_mask.sum()) == 2 * int(df.sum())
    two = data.train_pos_edge_index[:, data.sdf_mask]
    assert int(data.sdf_node_2hop_mask.sum()) == two.flatten().unique().numel()
    assert bool((data.df_mask <= data.sdf_mask).all())          # Df edges are inside S_Df
    assert bool((data.sdf_node_1hop_mask <= data.sdf_node_2hop_mask).all())


def test_dense_pair_mask_counts():
    shape, raw, df, data, neg = U.make_case('cora', 0.02)
    m = OU.dense_pair_mask(data, data.sdf_node_2hop_mask)
    k = int(data.sdf_node_2hop_mask.sum())
    assert int(m.sum()) == k * (k - 1) // 2 - int(data.df_mask.sum()) // 2


def test_golden_vectors():
    """tests/golden/oracle_small.npz was produced by tests/golden/make_golden.py from this oracle;
    it freezes the oracle's outputs so that later edits to the restatement are visible."""
    from tests.golden import make_golden
    want = np.load(GOLDEN)
    got = make_golden.compute()
    assert set(want.files) == set(got)
    for k in want.files:
        if want[k].dtype.kind in 'biu':
            assert np.array_equal(want[k], got[k]), k
        else:
            np.testing.assert_allclose(got[k], want[k], rtol=1e-9, atol=1e-12, err_msg=k)


def test_pyg_published_known_answers():
    """Known-answer vectors from PyTorch-Geometric's own documentation and unit tests - the ``k_hop_subgraph`` and
    ``to_undirected`` docstring examples, ``test/utils/test_subgraph.py``, ``test_undirected.py``, ``test_softmax.py``,
    ``test_loop.py`` (PyG 2.0-2.2) - written down from the published sources, each re-derived by hand from the
    documented algorithm; no PyG install is available here to execute them, so they narrow, but do not close, the
    "parity unpinned" gap (oracle/__init__.py).  PyG relabels nodes in its examples; the oracle keeps global ids."""
    from oracle import pyg_ops as P
    # k_hop_subgraph docstring: 2 hops around node 6 (flow source_to_target)
    ei = torch.tensor([[0, 1, 2, 3, 4, 5], [2, 2, 4, 4, 6, 6]])
    subset, sub_ei, inv, edge_mask = P.k_hop_subgraph(6, 2, ei)
    assert subset.tolist() == [2, 3, 4, 5, 6]
    assert edge_mask.tolist() == [False, False, True, True, True, True]
    relabel = {int(v): i for i, v in enumerate(subset)}
    assert [[relabel[int(v)] for v in r] for r in sub_ei] == [[0, 1, 2, 3], [2, 2, 4, 4]]
    assert inv.tolist() == [4]
    # test_subgraph.py: two seeds
    ei = torch.tensor([[1, 2, 4, 5], [0, 1, 5, 6]])
    subset, sub_ei, inv, edge_mask = P.k_hop_subgraph([0, 6], 2, ei)
    assert subset.tolist() == [0, 1, 2, 4, 5, 6]
    assert edge_mask.tolist() == [True, True, True, True] and inv.tolist() == [0, 5]
    # to_undirected docstring / test_undirected.py: duplicates merge, attributes add
    ei = torch.tensor([[0, 1, 1], [1, 0, 2]])
    sym, (w,) = P.to_undirected(ei, [torch.tensor([1., 1., 1.])])
    assert sym.tolist() == [[0, 1, 1, 2], [1, 0, 2, 1]] and w.tolist() == [2., 2., 1., 1.]
    assert P.is_undirected(sym) and not P.is_undirected(ei)
    # test_softmax.py
    out = P.segment_softmax(torch.tensor([1., 1., 1., 1.]), torch.tensor([0, 0, 1, 2]), 3)
    assert out.tolist() == [0.5, 0.5, 1.0, 1.0]
    # test_loop.py (add_remaining_self_loops): existing loops move to the end, one loop per node
    ei = torch.tensor([[0, 1, 0], [1, 0, 0]])
    assert P.add_remaining_self_loops(ei, 2).tolist() == [[0, 1, 0, 1], [1, 0, 0, 1]]
    # Kipf & Welling's renormalised adjacency D^-1/2 (A + I) D^-1/2 on the path 0 - 1 - 2, by hand
    path = torch.tensor([[0, 1, 1, 2], [1, 0, 2, 1]])
    a_hat = torch.tensor([[1 / 2, 1 / 6 ** 0.5, 0.], [1 / 6 ** 0.5, 1 / 3, 1 / 6 ** 0.5], [0., 1 / 6 ** 0.5, 1 / 2]],
                         dtype=torch.float64)
    got = P.gcn_conv(torch.eye(3, dtype=torch.float64), path, torch.eye(3, dtype=torch.float64), None)
    torch.testing.assert_close(got, a_hat)
