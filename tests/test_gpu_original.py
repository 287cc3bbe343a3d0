"""GPU: training of the original model and the retrain baseline through the drop-in trainers
(`framework.get_trainer` with `--unlearning_model original | retrain`; reference base.py:75-142, retrain.py:38-131)
against the oracle run with the same negatives and optimizer - every conv parameter gets its gradient from the
CUDA kernels (weight-gradient GEMMs, aggregation transpose, GAT score gradients, pair-decode incidence gather)."""
import os
import types

import pytest
import torch

from tests import util as U

pytestmark = pytest.mark.gpu
DEV = 'cuda'
EPOCHS = 3


def _args(tmp, **kw):
    base = dict(unlearning_model='original', gnn='gcn', dataset='Cora', in_dim=128, hidden_dim=128, out_dim=64,
                epochs=EPOCHS, valid_freq=100, lr=0.05, alpha=0.5, checkpoint_dir=str(tmp), random_seed=42,
                num_edge_type=None, eval_on_cpu=False)
    base.update(kw)
    return types.SimpleNamespace(**base)


def _negatives(n, count):
    g = torch.Generator().manual_seed(7)
    return [torch.randint(0, n, (2, count), generator=g) for _ in range(EPOCHS)]


@pytest.mark.parametrize('gnn,mode', [('gcn', 'original'), ('gat', 'original'), ('gin', 'original'), ('gcn', 'retrain')])
def test_original_and_retrain_training_vs_oracle(lib, tmp_path, gnn, mode):
    import framework
    from oracle import unlearn as OU
    shape, raw, df, data, _ = U.make_case('pubmed' if gnn == 'gat' else 'cora', 0.05, in_dim=128)
    retrain = mode == 'retrain'
    count = int(data.dr_mask.sum()) if retrain else data.train_pos_edge_index.shape[1]
    negs = _negatives(shape.num_nodes, count)
    # learning rates at which three steps move every parameter by 1-20 % (GIN sums un-normalised neighbourhoods:
    # its logits and gradients are ~1e3 times larger)
    args = _args(tmp_path, gnn=gnn, unlearning_model=mode, lr=1e-5 if gnn == 'gin' else 0.05)
    # oracle, fp64.  Plain SGD: the update is linear in the gradient, so the comparison measures the gradients
    # (Adam's first steps are lr * sign(g), which turns fp32 noise on near-zero entries into full-size steps)
    om = U.oracle_model(gnn, shape, data, dtype=torch.float64, delete=False)
    init = {k: v.float().clone() for k, v in om.state_dict().items()}
    d64 = data.clone(); d64.x = data.x.double()
    opt = torch.optim.SGD(om.parameters(), lr=args.lr)
    want = torch.stack([OU.link_train_epoch(om, d64, negs[e], opt, retrain=retrain) for e in range(EPOCHS)])

    model = framework.get_model(args, num_nodes=data.num_nodes, num_edge_type=None)
    assert type(model).__name__ == gnn.upper()
    model.load_state_dict(init)
    model = model.to(DEV)
    optimizer = torch.optim.SGD(model.parameters(), lr=args.lr)          # train_gnn.py:83 builds Adam over all parameters
    trainer = framework.get_trainer(args)
    assert type(trainer).__name__ == {'original': 'Trainer', 'retrain': 'RetrainTrainer'}[mode]
    trainer.negative_sampler = lambda data_, ei, cnt, epoch: negs[epoch].to(DEV)
    trainer.train(model, data.clone(), optimizer, args)
    hist = torch.tensor([l['train_loss'] for l in trainer.trainer_log['log'] if 'train_loss' in l])
    U.assert_close(hist, want, tol=1e-4, what='BCE loss curve')
    got, ref = model.state_dict(), om.state_dict()
    for k in ref:
        if ref[k].dtype.is_floating_point and ref[k].numel() > 1:
            U.assert_close(got[k], ref[k], tol=1e-4, what=f'{gnn} {k} after {EPOCHS} SGD steps')
    assert os.path.exists(os.path.join(args.checkpoint_dir, 'model_final.pt'))


def test_original_training_adam_with_validation(lib, tmp_path):
    """train_gnn.py's configuration (Adam over all parameters) with the eval / best-checkpoint leg: loss falls,
    checkpoints and node embeddings are written, the checkpoint loads into the Delete model (delete_gnn.py:206-207)."""
    import framework
    shape, raw, df, data, _ = U.make_case('cora', 0.05)
    args = _args(tmp_path, epochs=12, valid_freq=4, lr=0.01)
    torch.manual_seed(0)
    model = framework.get_model(args, num_nodes=data.num_nodes, num_edge_type=None).to(DEV)
    optimizer = torch.optim.Adam(model.parameters(), lr=args.lr)
    trainer = framework.get_trainer(args)
    d = data.clone()
    d.dtrain_mask = torch.ones(d.train_pos_edge_index.shape[1], dtype=torch.bool)      # train_gnn.py:45
    trainer.train(model, d, optimizer, args)
    losses = [l['train_loss'] for l in trainer.trainer_log['log'] if 'train_loss' in l]
    assert len(losses) == 12 and losses[-1] < losses[0]
    for f in ('model_best.pt', 'model_final.pt', 'node_embeddings.pt'):
        assert os.path.exists(os.path.join(args.checkpoint_dir, f)), f
    ckpt = torch.load(os.path.join(args.checkpoint_dir, 'model_best.pt'))
    dargs = _args(tmp_path, unlearning_model='gnndelete')
    dm = framework.get_model(dargs, data.sdf_node_1hop_mask, data.sdf_node_2hop_mask, num_nodes=data.num_nodes)
    missing = dm.load_state_dict(ckpt['model_state'], strict=False)
    assert sorted(missing.missing_keys) == ['deletion1.deletion_weight', 'deletion2.deletion_weight']
    assert not missing.unexpected_keys
