"""CPU: the GraphSAINT random-walk sampler of the mini-batch loops (gnndelete.py:333-337), plain tensor ops: walks follow
edges, batches are induced subgraphs in PyG's ordering, attributes are sliced by the rule of graph_saint.py's collate."""
import torch

from gnndelete_b200.data import GraphData
from gnndelete_b200.sampler import GraphSAINTRandomWalkSampler


def _data(n=300, m=1500, seed=0, isolated=20):
    g = torch.Generator().manual_seed(seed)
    e = torch.randint(0, n - isolated, (2, m), generator=g)            # the last `isolated` nodes have no edges
    e = torch.unique(e[:, e[0] != e[1]], dim=1)
    ei = torch.cat([e, e.flip(0)], 1)
    ei = torch.unique(ei, dim=1)[:, torch.randperm(torch.unique(ei, dim=1).shape[1], generator=g)]
    E = ei.shape[1]
    d = GraphData(num_nodes=n, edge_index=ei, train_pos_edge_index=ei, x=torch.randn(n, 8, generator=g),
                  node_id=torch.arange(n), sdf_mask=torch.rand(E, generator=g) < 0.5, df_mask=torch.rand(E, generator=g) < 0.1,
                  sdf_node_2hop_mask=torch.rand(n, generator=g) < 0.3, name='toy', scalar=torch.tensor(3.0))
    return d, g


def test_random_walk_follows_edges_and_stays_on_isolated_nodes():
    d, g = _data()
    s = GraphSAINTRandomWalkSampler(d, batch_size=64, walk_length=2, num_steps=1, generator=g)
    start = torch.cat([torch.arange(0, 280, 5), torch.arange(280, 300)])
    walk = s.random_walk(start)
    assert walk.shape == (start.numel(), 3) and torch.equal(walk[:, 0], start)
    edges = set(map(tuple, d.edge_index.t().tolist()))
    has_out = torch.bincount(d.edge_index[0], minlength=300) > 0
    for a, b, c in walk.tolist():
        for u, v in ((a, b), (b, c)):
            assert (u, v) in edges if has_out[u] else u == v
    # uniform over the neighbours: every neighbour of a well-connected node is eventually visited
    hub = int(torch.bincount(d.edge_index[0]).argmax())
    nbrs = set(d.edge_index[1][d.edge_index[0] == hub].tolist())
    seen = set(s.random_walk(torch.full((4000,), hub))[:, 1].tolist())
    assert seen == nbrs


def test_batches_are_induced_subgraphs_with_sliced_attributes():
    d, g = _data()
    s = GraphSAINTRandomWalkSampler(d, batch_size=40, walk_length=2, num_steps=5, generator=g)
    assert len(s) == 5
    batches = list(s)
    assert len(batches) == 5
    n, E = d.num_nodes, d.edge_index.shape[1]
    for b in batches:
        node_idx = b.node_id                                         # global ids of the batch's nodes (sliced arange)
        assert torch.equal(node_idx, torch.unique(node_idx)) and b.num_nodes == node_idx.numel() <= 40 * 3
        inside = torch.zeros(n, dtype=torch.bool)
        inside[node_idx] = True
        keep = inside[d.edge_index[0]] & inside[d.edge_index[1]]
        want = d.edge_index[:, keep]
        order = torch.argsort(want[0] * n + want[1])                 # PyG: (row, col) order of the sorted adjacency
        want, want_ids = want[:, order], keep.nonzero().squeeze(1)[order]
        assert torch.equal(node_idx[b.edge_index], want)             # relabelled edges map back to the induced edges
        assert torch.equal(b.train_pos_edge_index, d.train_pos_edge_index)      # [2, E]: first dim is 2 -> passed through
        assert torch.equal(b.sdf_mask, d.sdf_mask[want_ids]) and torch.equal(b.df_mask, d.df_mask[want_ids])
        assert torch.equal(b.x, d.x[node_idx]) and torch.equal(b.sdf_node_2hop_mask, d.sdf_node_2hop_mask[node_idx])
        assert b.name == 'toy' and float(b.scalar) == 3.0
    assert len({tuple(b.node_id.tolist()) for b in batches}) > 1     # different batches


def test_sampler_is_reproducible_with_a_generator():
    d, _ = _data()
    a = [b.node_id for b in GraphSAINTRandomWalkSampler(d, 30, 2, 3, generator=torch.Generator().manual_seed(5))]
    b = [b.node_id for b in GraphSAINTRandomWalkSampler(d, 30, 2, 3, generator=torch.Generator().manual_seed(5))]
    assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_train_minibatch_host_loop_with_the_oracle_model(tmp_path):
    """`GNNDeleteTrainer.train_minibatch` (gnndelete.py:311-450) is device-agnostic host logic around the model: run it
    on the CPU with the ORACLE's GCNDelete standing in for the CUDA model (the product models refuse CPU tensors)."""
    import types
    from gnndelete_b200.trainer import GNNDeleteTrainer
    from tests import util as U
    shape, raw, df, data, neg = U.make_case('cora', 0.05)
    om = U.oracle_model('gcn', shape, data)
    w0 = [om.deletion1.deletion_weight.detach().clone(), om.deletion2.deletion_weight.detach().clone()]
    args = types.SimpleNamespace(unlearning_model='gnndelete', gnn='gcn', dataset='ogbl-collab', epochs=2, valid_freq=100,
                                 lr=1e-3, checkpoint_dir=str(tmp_path), random_seed=1, saint_minibatch=True, batch_size=64,
                                 num_steps=3, device='cpu')
    opt = torch.optim.Adam([p for n, p in om.named_parameters() if 'del' in n], lr=args.lr)
    tr = GNNDeleteTrainer(args)
    tr.train(om, data.clone(), opt, args)
    logs = [l for l in tr.trainer_log['log'] if 'train_loss' in l]
    assert len(logs) == 2 and all(l['train_loss'] == l['train_loss'] and l['train_loss'] > 0 for l in logs)
    assert not torch.equal(om.deletion1.deletion_weight, w0[0]) and not torch.equal(om.deletion2.deletion_weight, w0[1])
    assert bool(torch.isfinite(om.deletion2.deletion_weight).all())
    import os
    assert os.path.exists(os.path.join(args.checkpoint_dir, 'model_final.pt'))
    # ablation variants (:391-398) and the global-id NI target
    for unl, extra in (('gnndelete_ablation_random', {}), ('gnndelete_ablation_locality', {}), ('gnndelete', {'saint_global_z_ori': True})):
        a2 = types.SimpleNamespace(**{**vars(args), 'unlearning_model': unl, 'epochs': 1, **extra})
        t2 = GNNDeleteTrainer(a2)
        t2.train(om, data.clone(), opt, a2)
        row = [l for l in t2.trainer_log['log'] if 'train_loss' in l][0]
        assert row['train_loss'] == row['train_loss']
        if unl.endswith('random'):
            assert row['loss_l'] == 0.0
        if unl.endswith('locality'):
            assert row['loss_r'] == 0.0
