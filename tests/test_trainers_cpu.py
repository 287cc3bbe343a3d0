"""CPU: the trainers' host loops (schedules, optimizer handling, logging, checkpoints) are device agnostic; only the
models need CUDA.  Run them here with the ORACLE's models standing in for the CUDA models (`args.device = 'cpu'`) and
compare with the oracle's own restatement of the reference loops.  The fused-kernel paths (`loss_fct = mse_*`, the GCN
engine) are GPU-only and covered by tests/test_gpu_*.py."""
import os
import types

import pytest
import torch

from tests import util as U
from tests.test_nodeemb_cpu import _oracle_run


def _args(tmp, **kw):
    base = dict(unlearning_model='gnndelete_nodeemb', gnn='gcn', dataset='Cora', in_dim=128, hidden_dim=128, out_dim=64,
                epochs=3, valid_freq=100, lr=1e-3, alpha=0.4, checkpoint_dir=str(tmp), random_seed=42, num_edge_type=None,
                eval_on_cpu=False, loss_fct='cosine_mean', loss_type='both_layerwise', device='cpu')
    base.update(kw)
    return types.SimpleNamespace(**base)


@pytest.mark.parametrize('loss_type', ['both_all', 'both_layerwise', 'only2_layerwise', 'only2_all', 'only1'])
def test_nodeemb_trainer_schedules_match_the_oracle(tmp_path, loss_type):
    """Every `loss_type` branch of gnndelete_nodeemb.py:219-299 through `GNNDeleteNodeembTrainer` (generic loss path)."""
    import framework
    shape, raw, df, data, neg = U.make_case('cora', 0.02)
    args = _args(tmp_path, loss_type=loss_type)
    model = U.oracle_model('gcn', shape, data, dtype=torch.float64)
    for n, p in model.named_parameters():                        # see tests/test_nodeemb_cpu.py::_oracle_run
        if 'del' not in n:
            p.requires_grad_(False)
    if 'layerwise' in loss_type:
        optimizer = [torch.optim.Adam(model.deletion1.parameters(), lr=args.lr),
                     torch.optim.Adam(model.deletion2.parameters(), lr=args.lr)]
    else:
        optimizer = torch.optim.Adam([p for n, p in model.named_parameters() if 'del' in n], lr=args.lr)
    trainer = framework.get_trainer(args)
    assert type(trainer).__name__ == 'GNNDeleteNodeembTrainer'
    d = data.clone()
    d.x = data.x.double()
    d.neg_edge_index = neg
    trainer.train(model, d, optimizer, args)
    hist = torch.tensor([[l['train_loss'], l['loss_r'], l['loss_l']] for l in trainer.trainer_log['log'] if 'train_loss' in l],
                        dtype=torch.float64)
    om, _, want = _oracle_run(loss_type, 'cosine_mean')
    torch.testing.assert_close(hist, want, rtol=1e-9, atol=1e-12)
    torch.testing.assert_close(model.deletion1.deletion_weight, om.deletion1.deletion_weight, rtol=1e-9, atol=1e-12)
    torch.testing.assert_close(model.deletion2.deletion_weight, om.deletion2.deletion_weight, rtol=1e-9, atol=1e-12)
    assert os.path.exists(os.path.join(args.checkpoint_dir, 'model_final.pt'))
    with pytest.raises(ValueError):                              # optimizer form must match the loss type (delete_gnn.py:221-226)
        wrong = optimizer[0] if isinstance(optimizer, list) else [optimizer, optimizer]
        trainer.train(model, d, wrong, args)


def test_nodeemb_trainer_rejects_unknown_names(tmp_path):
    import framework
    shape, raw, df, data, neg = U.make_case('cora', 0.02)
    model = U.oracle_model('gcn', shape, data)
    opt = torch.optim.Adam(model.deletion1.parameters())
    with pytest.raises(NotImplementedError):
        framework.get_trainer(_args(tmp_path, loss_type='nope')).train(model, data, opt, _args(tmp_path, loss_type='nope'))
    with pytest.raises(NotImplementedError):
        a = _args(tmp_path, loss_type='only1', loss_fct='nope')
        framework.get_trainer(a).train(model, data, opt, a)
    with pytest.raises(NotImplementedError):
        framework.get_trainer(_args(tmp_path, unlearning_model='graph_eraser'))


@pytest.mark.parametrize('mode', ['original', 'retrain'])
def test_original_and_retrain_loops_match_the_oracle(tmp_path, mode):
    """`Trainer.train_fullbatch` (base.py:75-142) / `RetrainTrainer` (retrain.py:38-131): negatives hook, BCE step, eval
    every `valid_freq`, best / final checkpoints, log keys."""
    import framework
    from oracle import unlearn as OU
    shape, raw, df, data, _ = U.make_case('cora', 0.05)
    retrain = mode == 'retrain'
    count = int(data.dr_mask.sum()) if retrain else data.train_pos_edge_index.shape[1]
    g = torch.Generator().manual_seed(7)
    negs = [torch.randint(0, shape.num_nodes, (2, count), generator=g) for _ in range(4)]
    args = _args(tmp_path, unlearning_model=mode, epochs=4, valid_freq=2, lr=0.01)
    # float64 on both sides: the CPU scatter-adds are multi-threaded, and Adam turns fp32 summation-order noise on
    # near-zero gradient entries into visible weight differences
    ref = U.oracle_model('gcn', shape, data, dtype=torch.float64, delete=False)
    model = U.oracle_model('gcn', shape, data, dtype=torch.float64, delete=False)
    data.x = data.x.double()
    opt_ref = torch.optim.Adam(ref.parameters(), lr=args.lr)
    want = torch.stack([OU.link_train_epoch(ref, data, negs[e], opt_ref, retrain=retrain) for e in range(4)])
    trainer = framework.get_trainer(args)
    trainer.negative_sampler = lambda data_, ei, cnt, epoch: negs[epoch]
    seen = []
    trainer._train_edges, inner = (lambda d: (seen.append(1), inner(d))[1]), trainer._train_edges
    optimizer = torch.optim.Adam(model.parameters(), lr=args.lr)
    trainer.train(model, data.clone(), optimizer, args)
    hist = torch.tensor([l['train_loss'] for l in trainer.trainer_log['log'] if 'train_loss' in l], dtype=torch.float64)
    torch.testing.assert_close(hist, want, rtol=1e-6, atol=1e-9)      # the trainer's link labels are float32 (base.py:45-50)
    for k, v in ref.state_dict().items():
        torch.testing.assert_close(model.state_dict()[k], v, rtol=1e-5, atol=1e-7)
    vals = [l for l in trainer.trainer_log['log'] if 'val_dt_auc' in l]
    assert len(vals) == 2 and len(seen) == 4
    assert all(0.0 <= v['val_dt_auc'] <= 1.0 and 0.0 <= v['val_dt_aup'] <= 1.0 for v in vals)
    if retrain:
        assert all(0.0 <= v['val_df_auc'] <= 1.0 for v in vals) and 'best_metric' in trainer.trainer_log
    else:
        assert all(v['val_df_auc'] != v['val_df_auc'] for v in vals) and 'best_valid_loss' in trainer.trainer_log   # nan: no Df for 'original'
        assert os.path.exists(os.path.join(args.checkpoint_dir, 'node_embeddings.pt'))
    for f in ('model_best.pt', 'model_final.pt', 'training_args.json'):
        assert os.path.exists(os.path.join(args.checkpoint_dir, f)), f
    trainer.test(model, data.clone())
    trainer.save_log()
    assert os.path.exists(os.path.join(args.checkpoint_dir, 'pred_proba.pt'))
    assert tuple(torch.load(os.path.join(args.checkpoint_dir, 'pred_proba.pt')).shape) == (shape.num_nodes, shape.num_nodes)


def _sk_df_metrics(df_logit, dr_logit, samples):
    from sklearn.metrics import average_precision_score, roc_auc_score
    lab = [0] * len(df_logit) + [1] * len(df_logit)
    aucs = [roc_auc_score(lab, df_logit + dr_logit[i].tolist()) for i in samples]
    aups = [average_precision_score(lab, df_logit + dr_logit[i].tolist()) for i in samples]
    return sum(aucs) / len(aucs), sum(aups) / len(aups)


def test_eval_matches_the_reference_formulas(tmp_path):
    """`Trainer.eval` (base.py:229-305) with the oracle's GCNDelete: BCE on the sigmoid outputs (sic), Dt AUC / AP and the
    resampled Df-vs-Dr AUC / AP against sklearn, which is what the reference calls."""
    from sklearn.metrics import average_precision_score, roc_auc_score
    from gnndelete_b200.trainer import Trainer
    shape, raw, df, data, neg = U.make_case('cora', 0.05)
    om = U.oracle_model('gcn', shape, data)
    tr = Trainer(_args(tmp_path, unlearning_model='gnndelete'))
    loss, dt_auc, dt_aup, df_auc, df_aup, df_logit, _, log = tr.eval(om, data, 'val', num_df_resamples=6)
    with torch.no_grad():
        z = om(data.x, data.train_pos_edge_index[:, data.dr_mask])
        prob = om.decode(z, data.val_pos_edge_index, data.val_neg_edge_index).sigmoid()
        dr_logit = om.decode(z, data.train_pos_edge_index[:, data.dr_mask]).sigmoid()
    label = torch.cat([torch.ones(data.val_pos_edge_index.shape[1]), torch.zeros(data.val_neg_edge_index.shape[1])])
    assert loss == pytest.approx(torch.nn.functional.binary_cross_entropy_with_logits(prob, label).item(), rel=1e-6)
    assert dt_auc == pytest.approx(roc_auc_score(label, prob), abs=1e-9)
    assert dt_aup == pytest.approx(average_precision_score(label, prob), abs=1e-9)
    assert len(df_logit) == data.directed_df_edge_index.shape[1] and tr.df_pos_edge.shape == (6, len(df_logit))
    want_auc, want_aup = _sk_df_metrics(df_logit, dr_logit, tr.df_pos_edge)
    assert df_auc == pytest.approx(want_auc, abs=1e-9) and df_aup == pytest.approx(want_aup, abs=1e-9)
    assert log['val_df_logit_mean'] == pytest.approx(sum(df_logit) / len(df_logit), rel=1e-5)
    # 'original': no Df leg (base.py:251-252)
    tr2 = Trainer(_args(tmp_path, unlearning_model='original'))
    out = tr2.eval(om, data, 'test')
    assert out[3] != out[3] and out[5] == []


def test_kg_eval_matches_the_reference_formulas(tmp_path):
    """`KGTrainer.eval` (base.py:494-566) with the oracle's RGCNDelete: DistMult logits, Dt metrics on the RAW logits, the
    Df leg on forward-direction retained triples."""
    import dataclasses
    from sklearn.metrics import average_precision_score, roc_auc_score
    from gnndelete_b200 import synthetic as S
    from gnndelete_b200.trainer import KGTrainer
    from oracle import unlearn as OU
    net = 9
    shape = dataclasses.replace(S.SHAPES['biokg'].scaled(0.002), num_edge_type=net)
    raw = S.make_graph(shape, seed=42)
    df = S.sample_df_mask(shape.num_edges, shape.num_deleted, seed=42)
    data = OU.build_unlearning_data(raw, df, num_edge_type=net)
    om = U.oracle_model('rgcn', shape, data, num_nodes=shape.num_nodes, num_edge_type=net)
    tr = KGTrainer(_args(tmp_path, unlearning_model='gnndelete_nodeemb', gnn='rgcn', dataset='ogbl-biokg', num_edge_type=net))
    loss, dt_auc, dt_aup, df_auc, df_aup, df_logit, pair, log = tr.eval(om, data, 'val', num_df_resamples=5)
    with torch.no_grad():
        z = om(data.x, data.edge_index[:, data.dr_mask], data.edge_type[data.dr_mask])
        ei = torch.cat([data.val_pos_edge_index, data.val_neg_edge_index], -1)
        logits = om.decode(z, ei, torch.cat([data.val_edge_type, data.val_edge_type]))
        half = data.dr_mask[:data.dr_mask.shape[0] // 2]
        dr_logit = om.decode(z, data.train_pos_edge_index[:, half], data.train_edge_type[half]).sigmoid()
    label = torch.cat([torch.ones(data.val_pos_edge_index.shape[1]), torch.zeros(data.val_neg_edge_index.shape[1])])
    assert loss == pytest.approx(torch.nn.functional.binary_cross_entropy_with_logits(logits, label).item(), rel=1e-6)
    assert dt_auc == pytest.approx(roc_auc_score(label, logits), abs=1e-9)
    assert dt_aup == pytest.approx(average_precision_score(label, logits), abs=1e-9)
    want_auc, want_aup = _sk_df_metrics(df_logit, dr_logit, tr.df_pos_edge)
    assert df_auc == pytest.approx(want_auc, abs=1e-9) and df_aup == pytest.approx(want_aup, abs=1e-9) and pair is None
    with pytest.raises(NotImplementedError):
        tr.train(om, data, None, None)


@pytest.mark.parametrize('mode', ['original', 'retrain'])
def test_original_minibatch_loop(tmp_path, mode):
    """`Trainer.train_minibatch` (base.py:144-227) over GraphSAINT batches, opt-in: runs, learns, evaluates, checkpoints."""
    import framework
    shape, raw, df, data, _ = U.make_case('cora', 0.05)
    args = _args(tmp_path, unlearning_model=mode, epochs=3, valid_freq=3, lr=0.01, saint_minibatch=True, batch_size=128,
                 num_steps=4, dataset='ogbl-collab')
    model = U.oracle_model('gcn', shape, data, delete=False)
    w0 = {k: v.clone() for k, v in model.state_dict().items()}
    trainer = framework.get_trainer(args)
    trainer.train(model, data.clone(), torch.optim.Adam(model.parameters(), lr=args.lr), args)
    steps = [l for l in trainer.trainer_log['log'] if 'train_loss' in l]
    assert len(steps) == 3 and all(l['train_loss'] == l['train_loss'] and l['train_loss'] > 0 for l in steps)
    assert any(not torch.equal(model.state_dict()[k], v) for k, v in w0.items())
    assert len([l for l in trainer.trainer_log['log'] if 'val_dt_auc' in l]) == 1
    for f in ('model_best.pt', 'model_final.pt'):
        assert os.path.exists(os.path.join(args.checkpoint_dir, f)), f


def test_nodeemb_minibatch_loop(tmp_path):
    """`GNNDeleteNodeembTrainer.train_minibatch` (gnndelete_nodeemb.py:352-494) over GraphSAINT batches, opt-in."""
    import framework
    shape, raw, df, data, neg = U.make_case('cora', 0.05)
    args = _args(tmp_path, epochs=2, valid_freq=2, saint_minibatch=True, batch_size=96, num_steps=3, dataset='ogbl-collab',
                 unlearning_model='gnndelete_nodeemb', alpha=0.5)
    model = U.oracle_model('gcn', shape, data)
    for n, p in model.named_parameters():
        if 'del' not in n:
            p.requires_grad_(False)
    w0 = [model.deletion1.deletion_weight.detach().clone(), model.deletion2.deletion_weight.detach().clone()]
    optimizer = [torch.optim.Adam(model.deletion1.parameters(), lr=1e-3), torch.optim.Adam(model.deletion2.parameters(), lr=1e-3)]
    trainer = framework.get_trainer(args)
    trainer.train(model, data.clone(), optimizer, args)
    rows = [l for l in trainer.trainer_log['log'] if 'train_loss' in l]
    assert len(rows) == 2 and all(r['train_loss'] == r['train_loss'] and r['train_loss'] > 0 for r in rows)
    assert all(r['train_loss'] == pytest.approx(0.5 * r['train_loss_r'] + 0.5 * r['train_loss_l'], rel=1e-5) for r in rows)
    assert not torch.equal(model.deletion1.deletion_weight, w0[0]) and not torch.equal(model.deletion2.deletion_weight, w0[1])
    assert len([l for l in trainer.trainer_log['log'] if 'val_dt_auc' in l]) == 1
    with pytest.raises(ValueError):
        trainer.train(model, data.clone(), optimizer[0], args)
