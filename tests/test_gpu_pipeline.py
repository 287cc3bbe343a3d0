"""GPU: the reference's two scripts end to end on a synthetic Cora-shaped graph, through the drop-in package only:
train_gnn.py (train the original model, test, save pred_proba.pt) -> delete_gnn.py (sample Df, build the deletion
masks, load the checkpoint into GCNDelete, unlearn with GNNDeleteTrainer on the dense-block NI objective of
train_fullbatch, test, save the log)."""
import json
import os
import types

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _args(ckpt_dir, **kw):
    base = dict(unlearning_model='original', gnn='gcn', dataset='Cora', in_dim=128, hidden_dim=128, out_dim=64,
                epochs=20, valid_freq=10, lr=0.01, alpha=0.5, checkpoint_dir=str(ckpt_dir), random_seed=42,
                num_edge_type=None, eval_on_cpu=False, loss_fct='mse_mean', loss_type='both_layerwise', df_size=5.0)
    base.update(kw)
    return types.SimpleNamespace(**base)


def test_train_then_unlearn_pipeline(lib, tmp_path):
    from framework import get_model, get_trainer                                  # delete_gnn.py:14, train_gnn.py:13
    from gnndelete_b200 import masks as MK
    from gnndelete_b200 import synthetic as S
    shape = S.SHAPES['cora'].scaled(0.05)
    raw = S.make_graph(shape, seed=42).to(DEV)                                    # stands in for d_{seed}.pkl
    n = raw.num_nodes

    # ---- train_gnn.py:45-91
    ori_dir = tmp_path / 'original'
    args = _args(ori_dir)
    data = raw.clone()
    data.train_pos_edge_index = MK.to_undirected(raw.train_pos_edge_index, num_nodes=n)
    data.dtrain_mask = torch.ones(data.train_pos_edge_index.shape[1], dtype=torch.bool, device=DEV)
    torch.manual_seed(0)
    model = get_model(args, num_nodes=n, num_edge_type=None).to(DEV)
    optimizer = torch.optim.Adam(model.parameters(), lr=args.lr)
    trainer = get_trainer(args)
    trainer.train(model, data, optimizer, args)
    trainer.test(model, data)
    trainer.save_log()
    losses = [l['train_loss'] for l in trainer.trainer_log['log'] if 'train_loss' in l]
    assert losses[-1] < losses[0]
    assert 0.0 <= trainer.trainer_log['dt_auc'] <= 1.0
    for f in ('model_best.pt', 'model_final.pt', 'pred_proba.pt', 'trainer_log.json', 'training_args.json'):
        assert os.path.exists(ori_dir / f), f

    # ---- delete_gnn.py:88-260
    del_dir = tmp_path / 'gnndelete'
    args = _args(del_dir, unlearning_model='gnndelete', epochs=30, valid_freq=15, lr=1e-3)
    df_mask = S.sample_df_mask(raw.train_pos_edge_index.shape[1], shape.num_deleted, seed=42, device=DEV)   # :95-110
    data = MK.build_unlearning_data(raw, df_mask)                                 # :113-189
    assert int(data.df_mask.sum()) == 2 * shape.num_deleted
    model = get_model(args, data.sdf_node_1hop_mask, data.sdf_node_2hop_mask, num_nodes=n, num_edge_type=None)   # :196
    logits_ori = torch.load(ori_dir / 'pred_proba.pt')                            # :200-203
    assert logits_ori.shape == (n, n)
    ckpt = torch.load(ori_dir / 'model_best.pt', map_location=DEV)
    model.load_state_dict(ckpt['model_state'], strict=False)                      # :206-207
    model = model.to(DEV)
    params = [{'params': [p for name, p in model.named_parameters() if 'del' in name], 'weight_decay': 0.0}]
    optimizer = torch.optim.Adam(params, lr=args.lr)                              # :229-241
    trainer = get_trainer(args)
    before = trainer.eval(model, data, 'test')
    trainer.train(model, data, optimizer, args, logits_ori)                       # :260 (dense-block NI: not 'ogbl')
    trainer.test(model, data)                                                     # :281-283
    trainer.save_log()
    log = json.load(open(del_dir / 'trainer_log.json'))
    steps = [l for l in log['log'] if 'train_loss' in l]
    assert len(steps) == args.epochs and all(l['train_loss'] == l['train_loss'] for l in steps)
    assert 0.0 <= log['df_auc'] <= 1.0 and 0.0 <= log['dt_auc'] <= 1.0
    # unlearning keeps the conv weights frozen and moves only the two Del operators
    final = torch.load(del_dir / 'model_final.pt')['model_state']
    for k, v in ckpt['model_state'].items():
        assert torch.equal(final[k].cpu(), v.cpu()), k
    assert not torch.equal(final['deletion1.deletion_weight'], torch.ones(128, 128) / 1000)
    assert os.path.exists(del_dir / 'pred_proba.pt') and before[1] == before[1]
