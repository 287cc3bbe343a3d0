"""GPU: the drop-in trainers driven through the reference's own call sequence
(`framework.get_model` / `get_trainer` / `trainer.train`, delete_gnn.py:196-260) against the
oracle run with the same schedule."""
import dataclasses
import os
import types

import pytest
import torch

from tests import util as U

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _args(tmp, **kw):
    base = dict(unlearning_model='gnndelete', gnn='gcn', dataset='Cora', in_dim=128, hidden_dim=128, out_dim=64,
                epochs=6, valid_freq=3, lr=1e-3, alpha=0.5, checkpoint_dir=str(tmp), random_seed=42,
                num_edge_type=None, eval_on_cpu=False, loss_fct='mse_mean', loss_type='both_layerwise')
    base.update(kw)
    return types.SimpleNamespace(**base)


def test_gnndelete_trainer_dropin(lib, tmp_path):
    import framework
    from oracle import unlearn as OU
    shape, raw, df, data, neg = U.make_case('cora', 0.05)
    args = _args(tmp_path)
    # oracle: same weights, same supplied negatives, 6 Adam steps
    om = U.oracle_model('gcn', shape, data, dtype=torch.float64)
    init = {k: v.float().clone() for k, v in om.state_dict().items()}
    d64 = data.clone(); d64.x = data.x.double()
    with torch.no_grad():
        zo = om.get_original_embeddings(d64.x, d64.train_pos_edge_index[:, d64.dr_mask])
    opt = torch.optim.Adam([p for n, p in om.named_parameters() if 'del' in n], lr=args.lr)
    for _ in range(args.epochs):
        loss, _, _, _ = OU.edge_form_loss(om, d64, neg, zo)
        loss.backward(); opt.step(); opt.zero_grad()

    model = framework.get_model(args, data.sdf_node_1hop_mask, data.sdf_node_2hop_mask, num_nodes=data.num_nodes,
                                num_edge_type=None)
    assert type(model).__name__ == 'GCNDelete'
    model.load_state_dict(init, strict=False)                    # delete_gnn.py:206-207
    params = [{'params': [p for n, p in model.named_parameters() if 'del' in n], 'weight_decay': 0.0}]
    optimizer = torch.optim.Adam(params, lr=args.lr)             # delete_gnn.py:229-241
    trainer = framework.get_trainer(args)
    d = data.clone()
    d.neg_edge_index = neg.to(DEV)                               # supplied negatives for parity
    trainer.train(model, d, optimizer, args)
    U.assert_close(model.deletion1.deletion_weight, om.deletion1.deletion_weight, tol=1e-4, what='W_del1')
    U.assert_close(model.deletion2.deletion_weight, om.deletion2.deletion_weight, tol=1e-4, what='W_del2')
    for f in ('training_args.json', 'model_final.pt', 'model_best.pt'):
        assert os.path.exists(os.path.join(args.checkpoint_dir, f)), f
    ckpt = torch.load(os.path.join(args.checkpoint_dir, 'model_final.pt'))
    assert set(ckpt['model_state']) == set(init)
    assert len(ckpt['optimizer_state']['state']) == 2
    logs = [l for l in trainer.trainer_log['log'] if 'train_loss' in l]
    assert len(logs) == args.epochs
    vals = [l for l in trainer.trainer_log['log'] if 'val_dt_auc' in l]
    assert len(vals) == 2 and 0.0 <= vals[0]['val_dt_auc'] <= 1.0
    # test() + save_log() as delete_gnn.py:281-283
    trainer.test(model, d.to(DEV))
    trainer.save_log()
    assert os.path.exists(os.path.join(args.checkpoint_dir, 'trainer_log.json'))


@pytest.mark.parametrize('gnn', ['gat', 'gin'])
def test_gnndelete_trainer_gat_gin(lib, tmp_path, gnn):
    """BASELINE config 2 route (`--gnn gat`) and GIN: the drop-in trainer against the oracle run with the same
    schedule (6 Adam steps, supplied negatives)."""
    import framework
    from oracle import unlearn as OU
    shape, raw, df, data, neg = U.make_case('pubmed' if gnn == 'gat' else 'cora', 0.1, in_dim=128)
    args = _args(tmp_path, gnn=gnn, dataset='PubMed')
    om = U.oracle_model(gnn, shape, data, dtype=torch.float64)
    init = {k: v.float().clone() for k, v in om.state_dict().items()}
    d64 = data.clone(); d64.x = data.x.double()
    with torch.no_grad():
        zo = om.get_original_embeddings(d64.x, d64.train_pos_edge_index[:, d64.dr_mask])
    opt = torch.optim.Adam([p for n, p in om.named_parameters() if 'del' in n], lr=args.lr)
    hist = []
    for _ in range(args.epochs):
        loss, _, _, _ = OU.edge_form_loss(om, d64, neg, zo)
        loss.backward(); opt.step(); opt.zero_grad()
        hist.append(float(loss))
    model = framework.get_model(args, data.sdf_node_1hop_mask, data.sdf_node_2hop_mask, num_nodes=data.num_nodes,
                                num_edge_type=None)
    assert type(model).__name__ == {'gat': 'GATDelete', 'gin': 'GINDelete'}[gnn]
    model.load_state_dict(init, strict=False)
    optimizer = torch.optim.Adam([p for n, p in model.named_parameters() if 'del' in n], lr=args.lr)
    trainer = framework.get_trainer(args)
    d = data.clone()
    d.neg_edge_index = neg.to(DEV)
    trainer.train(model, d, optimizer, args)
    U.assert_close(model.deletion1.deletion_weight, om.deletion1.deletion_weight, tol=1e-4, what='W_del1')
    U.assert_close(model.deletion2.deletion_weight, om.deletion2.deletion_weight, tol=1e-4, what='W_del2')
    assert trainer.trainer_log['captured_step'], trainer.trainer_log.get('capture_error')   # the step ran as one CUDA graph
    logs = [l['train_loss'] for l in trainer.trainer_log['log'] if 'train_loss' in l]
    U.assert_close(torch.tensor(logs), torch.tensor(hist), tol=1e-4, what='loss curve')
    assert os.path.exists(os.path.join(args.checkpoint_dir, 'model_final.pt'))


def test_eval_matches_sklearn(lib, tmp_path):
    """Trainer.eval's device AUC/AP against sklearn on the same logits (base.py:247-248)."""
    from sklearn.metrics import average_precision_score, roc_auc_score
    from gnndelete_b200 import models as M
    from gnndelete_b200.trainer import Trainer
    shape, raw, df, data, neg = U.make_case('cora', 0.05)
    args = _args(tmp_path)
    m = M.GCNDelete(U.args_for(shape), data.sdf_node_1hop_mask, data.sdf_node_2hop_mask).to(DEV)
    U.randomize(m)
    d = data.clone().to(DEV)
    tr = Trainer(args)
    loss, dt_auc, dt_aup, df_auc, df_aup, df_logit, _, log = tr.eval(m, d, 'val', num_df_resamples=5)
    with torch.no_grad():
        z = m(d.x, d.train_pos_edge_index[:, d.dr_mask])
        logits = m.decode(z, d.val_pos_edge_index, d.val_neg_edge_index).sigmoid().cpu()
    label = torch.cat([torch.ones(d.val_pos_edge_index.shape[1]), torch.zeros(d.val_neg_edge_index.shape[1])])
    assert dt_auc == pytest.approx(roc_auc_score(label, logits), abs=1e-9)
    assert dt_aup == pytest.approx(average_precision_score(label, logits), abs=1e-9)
    assert 0 <= df_auc <= 1 and len(df_logit) == d.directed_df_edge_index.shape[1]
    # Df-vs-resampled-Dr AUC / AP (base.py:264-280): the batched device version against sklearn, sample by sample
    dr_logit = m.decode(z, d.train_pos_edge_index[:, d.dr_mask]).sigmoid().cpu()
    lab = [0] * len(df_logit) + [1] * len(df_logit)
    aucs = [roc_auc_score(lab, df_logit + dr_logit[i.cpu()].tolist()) for i in tr.df_pos_edge]
    aups = [average_precision_score(lab, df_logit + dr_logit[i.cpu()].tolist()) for i in tr.df_pos_edge]
    assert len(aucs) == 5
    assert df_auc == pytest.approx(sum(aucs) / 5, abs=1e-9) and df_aup == pytest.approx(sum(aups) / 5, abs=1e-9)


def test_kg_trainer_matches_reference_schedule(lib, tmp_path):
    """Two backward passes + two Adam steps per step, with the deletion1 gradient of loss2
    carried into the next step (gnndelete_nodeemb.py:788-796)."""
    import framework
    from gnndelete_b200 import synthetic as S
    from oracle import unlearn as OU
    net = 51
    shape = dataclasses.replace(S.SHAPES['biokg'].scaled(0.002), num_edge_type=net)
    raw = S.make_graph(shape, seed=42)
    df = S.sample_df_mask(shape.num_edges, shape.num_deleted, seed=42)
    data = OU.build_unlearning_data(raw, df, num_edge_type=net)
    args = _args(tmp_path, gnn='rgcn', unlearning_model='gnndelete_nodeemb', dataset='ogbl-biokg', epochs=4,
                 num_edge_type=net)
    om = U.oracle_model('rgcn', shape, data, dtype=torch.float64, num_nodes=shape.num_nodes, num_edge_type=net)
    init = {k: v.float().clone() for k, v in om.state_dict().items()}
    pos_ei, pos_et = data.edge_index[:, data.df_mask], data.edge_type[data.df_mask]
    dec = pos_et < net
    neg = OU.negative_sampling_kg(pos_ei[:, dec], pos_et[dec], generator=torch.Generator().manual_seed(3))
    o1 = torch.optim.Adam(om.deletion1.parameters(), lr=args.lr)
    o2 = torch.optim.Adam(om.deletion2.parameters(), lr=args.lr)
    for _ in range(args.epochs):
        loss1, loss2, _ = OU.kg_step_losses(om, data, neg, net, alpha=args.alpha)
        loss1.backward(retain_graph=True); o1.step(); o1.zero_grad()
        loss2.backward(retain_graph=True); o2.step(); o2.zero_grad()

    model = framework.get_model(args, data.sdf_node_1hop_mask, data.sdf_node_2hop_mask, num_nodes=shape.num_nodes,
                                num_edge_type=net)
    assert type(model).__name__ == 'RGCNDelete'
    model.load_state_dict(init, strict=False)
    model = model.to(DEV)
    optimizer = [torch.optim.Adam(model.deletion1.parameters(), lr=args.lr),
                 torch.optim.Adam(model.deletion2.parameters(), lr=args.lr)]
    trainer = framework.get_trainer(args)
    assert type(trainer).__name__ == 'KGGNNDeleteNodeembTrainer'
    d = data.clone()
    d.neg_edge_index = neg.to(DEV)
    trainer.train(model, d, optimizer, args)
    U.assert_close(model.deletion1.deletion_weight, om.deletion1.deletion_weight, tol=1e-4, what='kg W_del1')
    U.assert_close(model.deletion2.deletion_weight, om.deletion2.deletion_weight, tol=1e-4, what='kg W_del2')


def test_kg_eval_matches_sklearn_and_oracle(lib, tmp_path):
    """KGTrainer.eval (base.py:494-566): DistMult logits against the oracle's decode, Dt AUC / AP on the raw logits and
    the Df-vs-resampled-Dr AUC / AP against sklearn; then the trainer's validation / best-checkpoint leg."""
    import framework
    from sklearn.metrics import average_precision_score, roc_auc_score
    from gnndelete_b200 import synthetic as S
    from oracle import unlearn as OU
    net = 51
    shape = dataclasses.replace(S.SHAPES['biokg'].scaled(0.002), num_edge_type=net)
    raw = S.make_graph(shape, seed=42)
    df = S.sample_df_mask(shape.num_edges, shape.num_deleted, seed=42)
    data = OU.build_unlearning_data(raw, df, num_edge_type=net)
    args = _args(tmp_path, gnn='rgcn', unlearning_model='gnndelete_nodeemb', dataset='ogbl-biokg', epochs=2, valid_freq=1,
                 num_edge_type=net)
    om = U.oracle_model('rgcn', shape, data, dtype=torch.float64, num_nodes=shape.num_nodes, num_edge_type=net)
    model = framework.get_model(args, data.sdf_node_1hop_mask, data.sdf_node_2hop_mask, num_nodes=shape.num_nodes,
                                num_edge_type=net)
    model.load_state_dict({k: v.float().clone() for k, v in om.state_dict().items()}, strict=False)
    model = model.to(DEV)
    trainer = framework.get_trainer(args)
    d = data.clone().to(DEV)
    loss, dt_auc, dt_aup, df_auc, df_aup, df_logit, _, log = trainer.eval(model, d, 'val', num_df_resamples=4)
    with torch.no_grad():
        zo = om(data.x, data.edge_index[:, data.dr_mask], data.edge_type[data.dr_mask])
        ei = torch.cat([data.val_pos_edge_index, data.val_neg_edge_index], -1)
        et = torch.cat([data.val_edge_type, data.val_edge_type], -1)
        want = om.decode(zo, ei, et)
        z = model(d.x, d.edge_index[:, d.dr_mask].contiguous(), d.edge_type[d.dr_mask].contiguous())
        got = model.decode(z, ei.to(DEV), et.to(DEV))
        half = d.dr_mask[:d.dr_mask.shape[0] // 2]
        dr_logit = model.decode(z, d.train_pos_edge_index[:, half].contiguous(), d.train_edge_type[half].contiguous()).sigmoid().cpu()
    U.assert_close(got, want, what='KG val logits')
    label = torch.cat([torch.ones(data.val_pos_edge_index.shape[1]), torch.zeros(data.val_neg_edge_index.shape[1])])
    assert dt_auc == pytest.approx(roc_auc_score(label, got.cpu()), abs=1e-9)
    assert dt_aup == pytest.approx(average_precision_score(label, got.cpu()), abs=1e-9)
    assert len(df_logit) == data.directed_df_edge_index.shape[1] and trainer.df_pos_edge.shape == (4, len(df_logit))
    lab = [0] * len(df_logit) + [1] * len(df_logit)
    aucs = [roc_auc_score(lab, df_logit + dr_logit[i.cpu()].tolist()) for i in trainer.df_pos_edge]
    aups = [average_precision_score(lab, df_logit + dr_logit[i.cpu()].tolist()) for i in trainer.df_pos_edge]
    assert df_auc == pytest.approx(sum(aucs) / 4, abs=1e-9) and df_aup == pytest.approx(sum(aups) / 4, abs=1e-9)
    # training with validation every epoch
    optimizer = [torch.optim.Adam(model.deletion1.parameters(), lr=args.lr),
                 torch.optim.Adam(model.deletion2.parameters(), lr=args.lr)]
    trainer.train(model, data.clone(), optimizer, args)
    vals = [l for l in trainer.trainer_log['log'] if 'val_dt_auc' in l]
    assert len(vals) == 2 and os.path.exists(os.path.join(args.checkpoint_dir, 'model_best.pt'))
    trainer.test(model, d)
    assert 0.0 <= trainer.trainer_log['dt_auc'] <= 1.0 and trainer.logit_all_pair is None      # 'ogbl': no all-pair logits


def test_negative_sampling_kg_permutes_heads_within_relation(lib):
    from gnndelete_b200.kg import negative_sampling_kg
    g = torch.Generator().manual_seed(0)
    ei = torch.randint(0, 50, (2, 400), generator=g).to(DEV)
    et = torch.randint(0, 7, (400,), generator=g).to(DEV)
    out = negative_sampling_kg(ei, et)
    assert torch.equal(out[1], ei[1])
    for r in range(7):
        m = et == r
        assert torch.equal(out[0, m].sort()[0], ei[0, m].sort()[0])
    assert not torch.equal(out[0], ei[0])
