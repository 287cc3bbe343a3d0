"""GPU: the data-preparation rows either side of the path (prepare_dataset.py:31-136, 203-265; delete_gnn.py:76-110) -
edge split and Df candidate masks bit-exact against the oracle, and the d_{seed}.pkl / df_{seed}.pt round trip."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _symmetric_graph(n=400, m=3000, seed=3):
    from gnndelete_b200.data import GraphData
    g = torch.Generator().manual_seed(seed)
    e = torch.randint(0, n, (2, m), generator=g)
    e = e[:, e[0] != e[1]]
    lo, hi = torch.minimum(e[0], e[1]), torch.maximum(e[0], e[1])
    e = torch.unique(torch.stack([lo, hi]), dim=1)
    sym = torch.cat([e, e.flip(0)], 1)
    return GraphData(num_nodes=n, edge_index=sym[:, torch.randperm(sym.shape[1], generator=g)],
                     x=torch.randn(n, 16, generator=g)), g


def test_split_and_df_masks_bit_exact(lib, tmp_path):
    from gnndelete_b200 import prepare as PR
    from oracle import unlearn as OU
    data, g = _symmetric_graph()
    m = int((data.edge_index[0] < data.edge_index[1]).sum())
    perm = torch.randperm(m, generator=g)
    train, test, val = OU.split_edges(data, perm, val_ratio=0.05, test_ratio=0.05)
    want = OU.df_candidate_masks(train, test, data.num_nodes)
    d = PR.train_test_split_edges(data.clone().to(DEV), val_ratio=0.05, test_ratio=0.05, perm=perm.to(DEV))
    assert torch.equal(d.train_pos_edge_index.cpu(), train)
    assert torch.equal(d.test_pos_edge_index.cpu(), test) and torch.equal(d.val_pos_edge_index.cpu(), val)
    assert d.test_neg_edge_index.shape == test.shape and d.val_neg_edge_index.shape == val.shape
    assert bool((d.train_pos_edge_index[0] < d.train_pos_edge_index[1]).all())
    masks = PR.df_candidate_masks(d)
    assert torch.equal(masks['in'].cpu(), want['in']) and torch.equal(masks['out'].cpu(), want['out'])
    assert 0 < int(masks['in'].sum()) < masks['in'].numel()
    # on-disk formats and the Df draw of delete_gnn.py:88-110
    PR.save_prepared(str(tmp_path), 'Synth', 42, d, masks, meta={'num_features': 16})
    raw = torch.load(tmp_path / 'Synth' / 'df_42.pt')
    assert sorted(raw) == ['in', 'out'] and raw['in'].dtype == torch.bool
    meta, back, cand = PR.load_prepared(str(tmp_path), 'Synth', 42, df='out')
    assert meta['num_features'] == 16 and torch.equal(back.train_pos_edge_index, train) and torch.equal(cand, want['out'])
    with pytest.raises(KeyError):
        PR.load_prepared(str(tmp_path), 'Synth', 42, df='none')
    E = train.shape[1]
    df_mask = PR.sample_df(cand.to(DEV), 5.0, E)
    assert int(df_mask.sum()) == int(5.0 / 100 * E) and not bool((df_mask & ~cand.to(DEV)).any())
    assert int(PR.sample_df(cand.to(DEV), 100, E).sum()) == 100


def test_kg_split_keeps_the_reference_quirk(lib):
    from gnndelete_b200 import prepare as PR
    from gnndelete_b200.data import GraphData
    g = torch.Generator().manual_seed(0)
    n, m = 200, 1000
    data = GraphData(num_nodes=n, edge_index=torch.randint(0, n, (2, m), generator=g), x=torch.arange(n),
                     edge_type=torch.randint(0, 7, (m,), generator=g)).to(DEV)
    perm = torch.randperm(m, generator=g).to(DEV)
    d = PR.train_test_split_edges(data, kg=True, perm=perm)                       # types stay in file order (:78, :100)
    n_v, n_t = 50, 100
    assert torch.equal(d.train_edge_type, data.edge_type[n_v + n_t:]) and torch.equal(d.test_edge_type, data.edge_type[:n_t])
    assert torch.equal(d.train_pos_edge_index, data.edge_index[:, perm][:, n_v + n_t:])
    assert torch.equal(d.test_neg_edge_index[1], d.test_pos_edge_index[1])       # negative_sampling_kg keeps tails
    f = PR.train_test_split_edges(data, kg=True, perm=perm, permute_edge_type=True)
    assert torch.equal(f.train_edge_type, data.edge_type[perm][n_v + n_t:])
