"""CPU: host logic of the data-preparation rows (`gnndelete_b200/prepare.py`: plain tensor ops and file formats) against
the oracle's restatement of prepare_dataset.py:31-136; the Df candidate masks need the CUDA k-hop kernel and are
covered by tests/test_gpu_prepare.py."""
import pytest
import torch

from tests.test_gpu_prepare import _symmetric_graph


def test_split_matches_oracle_and_partitions_the_edges():
    from gnndelete_b200 import prepare as PR
    from oracle import unlearn as OU
    data, g = _symmetric_graph()
    m = int((data.edge_index[0] < data.edge_index[1]).sum())
    perm = torch.randperm(m, generator=g)
    train, test, val = OU.split_edges(data, perm)
    d = PR.train_test_split_edges(data.clone(), perm=perm)
    assert torch.equal(d.train_pos_edge_index, train) and torch.equal(d.test_pos_edge_index, test)
    assert torch.equal(d.val_pos_edge_index, val)
    assert test.shape[1] == m // 10 and val.shape[1] == m // 20 and train.shape[1] + test.shape[1] + val.shape[1] == m
    keys = torch.cat([train, test, val], 1)
    assert torch.unique(keys[0] * data.num_nodes + keys[1]).numel() == m          # a partition: no edge twice
    with pytest.raises(ValueError):
        PR.train_test_split_edges(data.clone(), perm=perm[:-1])


def test_low_degree_edges_go_to_the_eval_splits_first():
    """prepare_dataset.py:52-61 ('ogbl'): edges with 2-hop degree < 50 are permuted in front, so test / val take them."""
    from gnndelete_b200 import prepare as PR
    data, g = _symmetric_graph()
    m = int((data.edge_index[0] < data.edge_index[1]).sum())
    deg = torch.full((m,), 100)
    deg[: m // 4] = 10
    d = PR.train_test_split_edges(data.clone(), two_hop_degree=deg, generator=g)
    row, col = data.edge_index
    keep = row < col
    low = set((row[keep][: m // 4] * data.num_nodes + col[keep][: m // 4]).tolist())
    ev = torch.cat([d.test_pos_edge_index, d.val_pos_edge_index], 1)
    assert all(int(k) in low for k in (ev[0] * data.num_nodes + ev[1]))


def test_prepared_files_round_trip(tmp_path):
    from gnndelete_b200 import prepare as PR
    from oracle import unlearn as OU
    data, g = _symmetric_graph()
    d = PR.train_test_split_edges(data.clone(), generator=g)
    masks = OU.df_candidate_masks(d.train_pos_edge_index, d.test_pos_edge_index, d.num_nodes)
    PR.save_prepared(str(tmp_path), 'Synth', 7, d, masks)
    meta, back, cand = PR.load_prepared(str(tmp_path), 'Synth', 7, df='in')
    assert meta['name'] == 'Synth' and torch.equal(cand, masks['in'])
    for k in ('train_pos_edge_index', 'test_pos_edge_index', 'val_neg_edge_index', 'x'):
        assert torch.equal(back[k], d[k])
    E = d.train_pos_edge_index.shape[1]
    df = PR.sample_df(cand, 2.5, E, generator=g)
    assert int(df.sum()) == int(2.5 / 100 * E) and not bool((df & ~cand).any())
