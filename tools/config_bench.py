"""Epochs/s of BASELINE configs 1, 2 and 4 through the drop-in trainers (whole graph per step, synthetic shapes),
next to the headline config 3 that bench.py times.  One JSON line per config.
usage: python tools/config_bench.py [--epochs 30] [--scale 1.0]"""
import argparse, json, os, sys, tempfile, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

ap = argparse.ArgumentParser()
ap.add_argument('--epochs', type=int, default=30)
ap.add_argument('--scale', type=float, default=1.0)
ap.add_argument('--configs', default='cora,cora-dense,pubmed,biokg',
                help="'cora-dense' = config 1 with train_fullbatch's dense-block NI loss (logits_ori given), 'cora' = edge-form NI")
a = ap.parse_args()
import framework
from gnndelete_b200 import masks as MK, synthetic as S
from gnndelete_b200.engine import GCNDeleteEngine
dev = 'cuda'


def targs(tmp, gnn, shape, **kw):
    d = dict(unlearning_model='gnndelete', gnn=gnn, dataset='synthetic', in_dim=shape.in_dim, hidden_dim=shape.hidden_dim,
             out_dim=shape.out_dim, epochs=a.epochs, valid_freq=10 ** 9, lr=1e-3, alpha=0.5, checkpoint_dir=tmp, random_seed=42,
             num_edge_type=shape.num_edge_type or None, eval_on_cpu=False, loss_fct='mse_mean', loss_type='both_layerwise')
    d.update(kw)
    return types.SimpleNamespace(**d)


for name in a.configs.split(','):
    dense_ni = name.endswith('-dense')
    shape = S.SHAPES[name.replace('-dense', '')].scaled(a.scale)
    kg = shape.num_edge_type > 0
    raw = S.make_graph(shape, seed=42, device='cpu').to(dev)
    df = S.sample_df_mask(shape.num_edges, shape.num_deleted, seed=42, device='cpu').to(dev)
    data = MK.build_unlearning_data(raw, df, num_edge_type=shape.num_edge_type if kg else None)
    tmp = tempfile.mkdtemp()
    args = targs(tmp, shape.gnn, shape)
    torch.manual_seed(42)
    model = framework.get_model(args, data.sdf_node_1hop_mask, data.sdf_node_2hop_mask, num_nodes=data.num_nodes,
                                num_edge_type=shape.num_edge_type if kg else None).to(dev)
    dels = [p for n, p in model.named_parameters() if 'del' in n]
    if kg:
        optimizer = [torch.optim.Adam([model.deletion1.deletion_weight], lr=1e-3), torch.optim.Adam([model.deletion2.deletion_weight], lr=1e-3)]
        from gnndelete_b200.kg import negative_sampling_kg          # supplied negatives: seed 43
        pe, pt = data.edge_index[:, data.df_mask], data.edge_type[data.df_mask]
        keep = pt < shape.num_edge_type
        data.neg_edge_index = negative_sampling_kg(pe[:, keep], pt[keep], torch.Generator(device=dev).manual_seed(43))
    else:
        optimizer = torch.optim.Adam(dels, lr=1e-3)
        data.neg_edge_index = S.supplied_negatives(shape.num_nodes, int(data.df_mask.sum()), seed=43, device='cpu').to(dev)
    trainer = framework.get_trainer(args)
    warm = targs(tmp, shape.gnn, shape, epochs=3)
    logits_ori = None
    if dense_ni:                                                # what base.py:288 stores in pred_proba.pt
        with torch.no_grad():
            zo_ = model.get_original_embeddings(data.x, data.train_pos_edge_index[:, data.dr_mask])
            logits_ori = zo_ @ zo_.t()
    extra = (logits_ori,) if dense_ni else ()
    trainer.train(model, data, optimizer, warm, *extra)         # plans, workspaces, module loading
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    trainer.train(model, data, optimizer, args, *extra)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    line = {'config': name, 'ni_loss': 'dense block (train_fullbatch)' if dense_ni else 'edge form', 'gnn': shape.gnn, 'nodes': shape.num_nodes, 'directed_edges': shape.num_edges,
            'deleted': shape.num_deleted, 'dims': [shape.in_dim, shape.hidden_dim, shape.out_dim], 'epochs': a.epochs,
            'epochs_per_s_through_trainer': a.epochs / dt, 'ms_per_epoch': 1e3 * dt / a.epochs,
            'note': 'wall clock around trainer.train (includes plan build of the second call, logging, checkpoint write)'}
    tt = sorted(l['train_time'] for l in trainer.trainer_log['log'] if 'train_time' in l)
    if tt:
        line['epochs_per_s_steady_state'] = 1.0 / tt[len(tt) // 4]       # trainer's own per-epoch time (logged every log_every epochs), lower quartile
    for k in ('captured_step', 'capture_error'):
        if k in trainer.trainer_log:
            line[k] = trainer.trainer_log[k]
    if shape.gnn == 'gcn' and not dense_ni:                      # the fused engine alone, CUDA-graph replay
        with torch.no_grad():
            zo = model.get_original_embeddings(data.x, data.train_pos_edge_index[:, data.dr_mask])
        eng = GCNDeleteEngine(model, data, data.neg_edge_index, z_ori=zo, hoist_layer1=False, static_negatives=True)
        eng.capture(warmup=2)
        for _ in range(5): eng.epoch()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200): eng.epoch()
        e1.record(); torch.cuda.synchronize()
        line['epochs_per_s_engine_graph'] = 200 / (e0.elapsed_time(e1) / 1e3)
    print(json.dumps(line), flush=True)
