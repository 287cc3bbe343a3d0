"""The three GraphSAINT mini-batch loops on the GPU (opt-in `args.saint_minibatch`): each loop is run
twice from the same seeds on cuda:0 - once with the product's CUDA model, once with the ORACLE's model moved to the GPU
(same sampler stream, so the same batches and negatives) - and the logged losses / trained weights must agree at 1e-4
(several optimizer steps compound fp32 rounding).
  GNNDeleteTrainer.train_minibatch          gnndelete.py:311-450
  GNNDeleteNodeembTrainer.train_minibatch   gnndelete_nodeemb.py:352-494
  Trainer.train_minibatch                   base.py:144-227"""
import types

import pytest
import torch

from tests import util as U

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _args(tmp_path, sub, **kw):
    d = dict(unlearning_model='gnndelete', gnn='gcn', dataset='ogbl-collab', epochs=2, valid_freq=100, lr=1e-3, alpha=0.5,
             checkpoint_dir=str(tmp_path / sub), random_seed=1, saint_minibatch=True, batch_size=96, num_steps=3, device=DEV,
             in_dim=128, hidden_dim=128, out_dim=64, loss_fct='mse_mean', loss_type='both_layerwise', eval_on_cpu=False,
             num_edge_type=None)
    d.update(kw)
    return types.SimpleNamespace(**d)


def _pair(gnn, shape, data, delete=True):
    """(product model on the GPU, oracle model on the GPU) with identical parameters."""
    import framework
    om = U.oracle_model(gnn, shape, data, delete=delete)
    a = types.SimpleNamespace(unlearning_model='gnndelete' if delete else 'original', gnn=gnn, in_dim=shape.in_dim,
                              hidden_dim=shape.hidden_dim, out_dim=shape.out_dim)
    m = framework.get_model(a, data.sdf_node_1hop_mask if delete else None, data.sdf_node_2hop_mask if delete else None,
                            num_nodes=data.num_nodes, num_edge_type=None)
    m.load_state_dict({k: v.clone() for k, v in om.state_dict().items()}, strict=True)
    return m.to(DEV), om.to(DEV)


def _losses(trainer, key='train_loss'):
    return torch.tensor([l[key] for l in trainer.trainer_log['log'] if key in l])


def test_gnndelete_train_minibatch_on_gpu(lib, tmp_path):
    from gnndelete_b200.trainer import GNNDeleteTrainer
    shape, raw, df, data, neg = U.make_case('cora', 0.05)
    m, om = _pair('gcn', shape, data)
    runs = []
    for sub, model in (('cuda', m), ('oracle', om)):
        args = _args(tmp_path, sub)
        opt = torch.optim.Adam([p for n, p in model.named_parameters() if 'del' in n], lr=args.lr)
        tr = GNNDeleteTrainer(args)
        tr.train(model, data.clone().to(DEV), opt, args)
        runs.append(tr)
    la, lb = _losses(runs[0]), _losses(runs[1])
    assert la.numel() == 2 and bool(torch.isfinite(la).all()) and bool((la > 0).all())
    U.assert_close(la, lb, tol=1e-4, what='mini-batch loss curve, CUDA model vs oracle model')
    U.assert_close(m.deletion1.deletion_weight, om.deletion1.deletion_weight, tol=1e-4, what='W_del1')
    U.assert_close(m.deletion2.deletion_weight, om.deletion2.deletion_weight, tol=1e-4, what='W_del2')


def test_nodeemb_train_minibatch_on_gpu(lib, tmp_path):
    import framework
    shape, raw, df, data, neg = U.make_case('cora', 0.05)
    m, om = _pair('gcn', shape, data)
    runs = []
    for sub, model in (('cuda', m), ('oracle', om)):
        for n, p in model.named_parameters():
            if 'del' not in n:
                p.requires_grad_(False)
        args = _args(tmp_path, sub, unlearning_model='gnndelete_nodeemb', valid_freq=2)
        optimizer = [torch.optim.Adam(model.deletion1.parameters(), lr=1e-3), torch.optim.Adam(model.deletion2.parameters(), lr=1e-3)]
        tr = framework.get_trainer(args)
        tr.train(model, data.clone().to(DEV), optimizer, args)
        runs.append(tr)
    la, lb = _losses(runs[0]), _losses(runs[1])
    assert la.numel() == 2 and bool(torch.isfinite(la).all()) and bool((la > 0).all())
    U.assert_close(la, lb, tol=1e-4, what='node-embedding mini-batch loss curve')
    U.assert_close(m.deletion1.deletion_weight, om.deletion1.deletion_weight, tol=1e-4, what='W_del1')
    U.assert_close(m.deletion2.deletion_weight, om.deletion2.deletion_weight, tol=1e-4, what='W_del2')
    assert len([l for l in runs[0].trainer_log['log'] if 'val_dt_auc' in l]) == 1


@pytest.mark.parametrize('mode', ['original', 'retrain'])
def test_original_train_minibatch_on_gpu(lib, tmp_path, mode):
    import framework
    shape, raw, df, data, neg = U.make_case('cora', 0.05)
    m, om = _pair('gcn', shape, data, delete=False)
    runs = []
    for sub, model in (('cuda', m), ('oracle', om)):
        args = _args(tmp_path, sub, unlearning_model=mode, epochs=3, valid_freq=3, lr=0.01, batch_size=128, num_steps=4)
        tr = framework.get_trainer(args)
        tr.train(model, data.clone().to(DEV), torch.optim.Adam(model.parameters(), lr=args.lr), args)
        runs.append(tr)
    la, lb = _losses(runs[0]), _losses(runs[1])
    assert la.numel() == 3 and bool(torch.isfinite(la).all()) and bool((la > 0).all())
    U.assert_close(la, lb, tol=1e-4, what=f'{mode} mini-batch loss curve')
    for (k, v), (_, w) in zip(m.state_dict().items(), om.state_dict().items()):
        U.assert_close(v, w, tol=1e-3, what=f'{mode} {k} after 12 Adam steps at lr 0.01')
