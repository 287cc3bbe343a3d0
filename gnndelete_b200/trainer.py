"""Host-side mirror of the reference trainers for the unlearning hot path
(``framework/trainer/base.py`` ``Trainer`` and ``framework/trainer/gnndelete.py``
``GNNDeleteTrainer``): same class names, constructor, ``train`` / ``eval`` / ``test`` /
``save_log`` signatures, log keys and checkpoint files.  The epoch body runs on the
fused CUDA engine (``engine.py``); wandb / tqdm logging of the reference is replaced by
the ``trainer_log`` dict only (no third-party logging dependency).
"""
from __future__ import annotations

import json
import os
import time

import torch

from . import metrics
from . import ops
from .engine import GATDeleteEngine, GCNDeleteEngine


def get_loss_fct(name):
    """``gnndelete.py:25-35``: the two train loops hard-code MSE; anything else raises."""
    if name in ('mse', 'mse_mean'):
        return torch.nn.MSELoss()
    raise NotImplementedError(name)


class Trainer:
    """``framework/trainer/base.py:24-35, 229-391`` (the parts GNNDelete uses)."""

    def __init__(self, args):
        self.args = args
        self.trainer_log = {'unlearning_model': args.unlearning_model, 'dataset': args.dataset, 'log': []}
        self.logit_all_pair = None
        self.df_pos_edge = []
        os.makedirs(args.checkpoint_dir, exist_ok=True)
        with open(os.path.join(args.checkpoint_dir, 'training_args.json'), 'w') as f:
            json.dump({k: v for k, v in vars(args).items() if _jsonable(v)}, f)

    # ---- training of the ORIGINAL model (base.py:66-142): produces the checkpoint the unlearning path starts from
    def negative_sampler(self, data, edge_index, count, epoch):
        """Negatives of one epoch.  PyG's ``negative_sampling`` (base.py:84-87) is randomised rejection sampling and
        not reproducible across implementations: uniform random pairs here; parity runs replace
        this hook (or supply ``data.train_neg_edge_index`` = a fixed ``[2, count]`` tensor)."""
        fixed = getattr(data, 'train_neg_edge_index', None)
        if fixed is not None:
            return fixed
        if getattr(self, '_neg_gen', None) is None:
            self._neg_gen = torch.Generator(device=edge_index.device).manual_seed(getattr(self.args, 'random_seed', 42))
        return torch.randint(0, data.num_nodes, (2, int(count)), generator=self._neg_gen, device=edge_index.device)

    def _train_edges(self, data):
        """Message-passing / supervision edge set and the number of negatives (base.py:84-92)."""
        ei = data.train_pos_edge_index
        count = int(data.dtrain_mask.sum()) if hasattr(data, 'dtrain_mask') else ei.shape[1]
        return ei, count

    best_by = 'valid_loss'        # base.py:120 keeps the lowest validation loss; RetrainTrainer the best dt_auc + df_auc

    def train(self, model, data, optimizer, args, logits_ori=None, attack_model_all=None, attack_model_sub=None):
        """``Trainer.train`` (base.py:66-73).  Both reference loops (full-batch and the GraphSAINT mini-batch one
        for 'ogbl' datasets) run the step below; here the whole graph is always the batch."""
        if getattr(args, 'saint_minibatch', False):
            return self.train_minibatch(model, data, optimizer, args)
        return self.train_fullbatch(model, data, optimizer, args)

    def train_minibatch(self, model, data, optimizer, args):
        """``Trainer.train_minibatch`` (base.py:144-227), opt-in through ``args.saint_minibatch``: GraphSAINT random-walk
        batches (``batch_size=args.batch_size, walk_length=2, num_steps=args.num_steps``) of the training graph, per batch
        as many uniform negatives as the batch has edges (``negative_sampling``'s default count, :160-162), the BCE step,
        evaluation on the whole graph every ``valid_freq`` epochs.  A batch without edges is skipped (BCE of nothing is
        nan in the reference).  Host loop tested on the CPU with the oracle's models; not yet run on a B200."""
        from .data import GraphData
        from .sampler import GraphSAINTRandomWalkSampler
        dev = torch.device(getattr(args, 'device', 'cuda'))
        model = model.to(dev)
        data = data.to(dev)
        ei, _ = self._train_edges(data)
        graph = GraphData(num_nodes=int(data.num_nodes), edge_index=ei, x=data.x)         # all the step reads (:148, :158-160)
        gen = torch.Generator(device=dev).manual_seed(getattr(args, 'random_seed', 42))
        loader = GraphSAINTRandomWalkSampler(graph, batch_size=args.batch_size, walk_length=2, num_steps=args.num_steps,
                                             generator=gen)
        t_start = time.time()
        best_valid_loss, best_epoch = 1000000, 0
        for epoch in range(args.epochs):
            model.train()
            total, steps = torch.zeros((), device=dev), 0
            for batch in loader:
                bei = batch.edge_index
                if bei.shape[1] == 0:
                    continue
                z = model(batch.x, bei)
                neg_edge_index = torch.randint(0, z.size(0), (2, bei.shape[1]), generator=gen, device=dev)
                logits = model.decode(z, bei, neg_edge_index)
                label = self.get_link_labels(bei, neg_edge_index)
                loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, label)
                loss.backward()
                optimizer.step()
                optimizer.zero_grad()
                total += loss.detach()
                steps += 1
            self.trainer_log['log'].append({'epoch': epoch, 'train_loss': (total / max(steps, 1)).item()})
            if (epoch + 1) % args.valid_freq == 0:
                valid_loss, dt_auc, dt_aup, df_auc, df_aup, _, _, valid_log = self.eval(model, data, 'val')
                valid_log['epoch'] = epoch
                self.trainer_log['log'].append(valid_log)
                if valid_loss < best_valid_loss:
                    best_valid_loss, best_epoch = valid_loss, epoch
                    torch.save({'model_state': model.state_dict(), 'optimizer_state': optimizer.state_dict()},
                               os.path.join(args.checkpoint_dir, 'model_best.pt'))
        self.trainer_log['training_time'] = time.time() - t_start
        torch.save({'model_state': model.state_dict(), 'optimizer_state': optimizer.state_dict()},
                   os.path.join(args.checkpoint_dir, 'model_final.pt'))
        self.trainer_log['best_epoch'], self.trainer_log['best_valid_loss'] = best_epoch, best_valid_loss
        return model

    def train_fullbatch(self, model, data, optimizer, args, logits_ori=None, attack_model_all=None, attack_model_sub=None):
        """``Trainer.train_fullbatch`` (base.py:75-142) / ``RetrainTrainer.train_fullbatch`` (retrain.py:38-131):
        per epoch new negatives, ``z = model(x, edges)``, ``BCEWithLogits(decode(z, edges, neg), labels)``, backward
        to EVERY parameter, ``optimizer.step()``.  Forward and backward run on the same kernels as the unlearning
        path (aggregation, tcgen05 GEMMs incl. their weight gradients, pair decode + incidence gather)."""
        dev = torch.device(getattr(args, 'device', 'cuda'))       # host loop is device agnostic; the CUDA models are not
        model = model.to(dev)
        data = data.to(dev)
        t_start = time.time()
        best_valid_loss, best_metric, best_epoch = 1000000, 0, 0
        ring = []
        for epoch in range(args.epochs):
            model.train()
            ei, count = self._train_edges(data)
            neg_edge_index = self.negative_sampler(data, ei, count, epoch)
            z = model(data.x, ei)
            logits = model.decode(z, ei, neg_edge_index)
            label = self.get_link_labels(ei, neg_edge_index)
            loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, label)
            loss.backward()
            optimizer.step()
            optimizer.zero_grad()
            ring.append(loss.detach())
            last = epoch + 1 == args.epochs
            if (epoch + 1) % self.log_every == 0 or last or (epoch + 1) % args.valid_freq == 0:
                for i, v in enumerate(torch.stack(ring).cpu().tolist()):
                    self.trainer_log['log'].append({'epoch': epoch + 1 - len(ring) + i, 'train_loss': v})
                ring = []
            if (epoch + 1) % args.valid_freq == 0:
                valid_loss, dt_auc, dt_aup, df_auc, df_aup, _, _, valid_log = self.eval(model, data, 'val')
                valid_log['epoch'] = epoch
                self.trainer_log['log'].append(valid_log)
                if self.best_by == 'valid_loss':
                    better = valid_loss < best_valid_loss
                else:
                    better = dt_auc + (df_auc if df_auc == df_auc else 0.0) > best_metric
                if better:
                    best_valid_loss, best_metric, best_epoch = valid_loss, dt_auc + (df_auc if df_auc == df_auc else 0.0), epoch
                    torch.save({'model_state': model.state_dict(), 'optimizer_state': optimizer.state_dict()},
                               os.path.join(args.checkpoint_dir, 'model_best.pt'))
                    if self.best_by == 'valid_loss':
                        torch.save(z.detach(), os.path.join(args.checkpoint_dir, 'node_embeddings.pt'))     # base.py:131
        self.trainer_log['training_time'] = time.time() - t_start
        torch.save({'model_state': model.state_dict(), 'optimizer_state': optimizer.state_dict()},
                   os.path.join(args.checkpoint_dir, 'model_final.pt'))
        self.trainer_log['best_epoch'] = best_epoch
        self.trainer_log['best_valid_loss' if self.best_by == 'valid_loss' else 'best_metric'] = \
            best_valid_loss if self.best_by == 'valid_loss' else best_metric
        return model

    log_every = 100

    @torch.no_grad()
    def get_link_labels(self, pos_edge_index, neg_edge_index):
        e = pos_edge_index.size(1) + neg_edge_index.size(1)
        labels = torch.zeros(e, dtype=torch.float, device=pos_edge_index.device)
        labels[:pos_edge_index.size(1)] = 1.
        return labels

    @torch.no_grad()
    def eval(self, model, data, stage='val', pred_all=False, num_df_resamples=500):
        """``base.py:229-305``: forward on the ``dr_mask`` edge set, decode val/test edges,
        BCE + AUC/AP on Dt, AUC/AP of Df against ``num_df_resamples`` random Dr samples.
        AUC / AP are computed on the device (``metrics.py``, rank based, same tie handling
        as sklearn) instead of 500 sklearn calls on Python lists."""
        model.eval()
        pos, neg = data[f'{stage}_pos_edge_index'], data[f'{stage}_neg_edge_index']
        mask = data.dtrain_mask if hasattr(data, 'dtrain_mask') else data.dr_mask
        z = model(data.x, data.train_pos_edge_index[:, mask])
        logits = model.decode(z, pos, neg).sigmoid()
        label = self.get_link_labels(pos, neg)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, label).item()
        dt_auc = metrics.roc_auc(label, logits)
        dt_aup = metrics.average_precision(label, logits)
        if self.args.unlearning_model in ['original']:
            df_logit = torch.empty(0, device=z.device)
        else:
            df_logit = model.decode(z, data.directed_df_edge_index).sigmoid()
        n_df = df_logit.numel()
        if n_df > 0:
            dr_edges = data.train_pos_edge_index[:, data.dr_mask]
            if len(self.df_pos_edge) == 0:
                g = torch.Generator(device='cpu').manual_seed(getattr(self.args, 'random_seed', 42))
                for _ in range(num_df_resamples):
                    self.df_pos_edge.append(torch.randperm(dr_edges.shape[1], generator=g)[:n_df].to(z.device))
            dr_logit = model.decode(z, dr_edges).sigmoid()
            # the 500 (Df vs random-Dr-sample) problems of base.py:264-280 in two batched passes
            idx = self.df_pos_edge if torch.is_tensor(self.df_pos_edge) else torch.stack(self.df_pos_edge)
            self.df_pos_edge = idx
            aucs, aups = metrics.resampled_auc_ap(df_logit, dr_logit, idx)
            df_auc = aucs.mean().item()
            df_aup = aups.mean().item()
        else:
            df_auc = df_aup = float('nan')
        logit_all_pair = (z @ z.t()).cpu() if pred_all else None
        log = {
            f'{stage}_loss': loss, f'{stage}_dt_auc': dt_auc, f'{stage}_dt_aup': dt_aup,
            f'{stage}_df_auc': df_auc, f'{stage}_df_aup': df_aup,
            f'{stage}_df_logit_mean': df_logit.mean().item() if n_df else float('nan'),
            f'{stage}_df_logit_std': df_logit.std(unbiased=False).item() if n_df else float('nan'),
        }
        return loss, dt_auc, dt_aup, df_auc, df_aup, df_logit.tolist(), logit_all_pair, log

    @torch.no_grad()
    def test(self, model, data, model_retrain=None, attack_model_all=None, attack_model_sub=None, ckpt='best'):
        """``base.py:307-375`` without the membership-inference / retrain-comparison legs
        (out of scope)."""
        if ckpt == 'best':
            path = os.path.join(self.args.checkpoint_dir, 'model_best.pt')
            if os.path.exists(path):
                model.load_state_dict(torch.load(path, map_location=data.x.device)['model_state'])
            else:
                # the reference raises here (base.py:313-315); no best checkpoint means the validation metric never
                # improved (e.g. a NaN df_auc): evaluate the final weights, but say so in the log
                self.trainer_log['best_checkpoint_missing'] = True
        pred_all = 'ogbl' not in self.args.dataset
        loss, dt_auc, dt_aup, df_auc, df_aup, df_logit, logit_all_pair, test_log = self.eval(model, data, 'test', pred_all)
        self.trainer_log['dt_loss'] = loss
        self.trainer_log['dt_auc'] = dt_auc
        self.trainer_log['dt_aup'] = dt_aup
        self.trainer_log['df_logit'] = df_logit
        self.logit_all_pair = logit_all_pair
        self.trainer_log['df_auc'] = df_auc
        self.trainer_log['df_aup'] = df_aup
        self.trainer_log['auc_sum'] = dt_auc + df_auc                     # base.py:327-330
        self.trainer_log['aup_sum'] = dt_aup + df_aup
        self.trainer_log['auc_gap'] = abs(dt_auc - df_auc)
        self.trainer_log['aup_gap'] = abs(dt_aup - df_aup)
        return loss, dt_auc, dt_aup, df_auc, df_aup, df_logit, logit_all_pair, test_log

    def save_log(self):
        with open(os.path.join(self.args.checkpoint_dir, 'trainer_log.json'), 'w') as f:
            json.dump(self.trainer_log, f)
        torch.save(self.logit_all_pair, os.path.join(self.args.checkpoint_dir, 'pred_proba.pt'))


def _jsonable(v):
    try:
        json.dumps(v)
        return True
    except TypeError:
        return False


class RetrainTrainer(Trainer):
    """``framework/trainer/retrain.py:28-131``: the retrain-from-scratch baseline - the same BCE link-prediction loop
    on the RETAINED edges (``train_pos_edge_index[:, dr_mask]``, as many negatives as retained edges, :56-63), best
    checkpoint by ``dt_auc + df_auc`` (:106-116)."""

    best_by = 'metric'

    def _train_edges(self, data):
        ei = data.train_pos_edge_index[:, data.dr_mask].contiguous()
        return ei, int(data.dr_mask.sum())


class GNNDeleteTrainer(Trainer):
    """``framework/trainer/gnndelete.py:37-450``.

    Both reference loops run one optimisation step per (sub)graph with
    ``loss = 0.5 * MSE(df_logits, neg_logits) + 0.5 * NI``.  Here the whole graph is the
    batch (180 GB of HBM make GraphSAINT sampling unnecessary) and NI is
    the edge form of ``train_minibatch`` (:379-386).  ``z_ori`` — which the reference
    fails to produce (``data.dtrain_mask`` is never set) — is the base
    model's embedding on the ``dr_mask`` edge set unless supplied as ``data.z_ori``.
    """

    log_every = 100      # epochs between device->host loss reads (the reference syncs 3x per epoch)

    def train(self, model, data, optimizer, args, logits_ori=None, attack_model_all=None, attack_model_sub=None):
        if not hasattr(model, 'deletion1'):
            raise NotImplementedError('GNNDeleteTrainer trains the *Delete models (deletion1 / deletion2)')
        if getattr(args, 'saint_minibatch', False):
            return self.train_minibatch(model, data, optimizer, args)
        # reference dispatch (gnndelete.py:39-44): 'ogbl' datasets take the mini-batch loop (edge-form NI),
        # everything else the full-batch loop whose NI term is the dense S_Df x S_Df block against the
        # original model's pair logits (`logits_ori`, pred_proba.pt) - for every architecture.  Without
        # `logits_ori` the edge form is used for every graph.
        dense = logits_ori is not None and 'ogbl' not in self.args.dataset
        if type(model).__name__ not in ('GCNDelete', 'GATDelete'):
            # GINDelete: the same step body through the autograd modules (every layer is still one of the CUDA
            # kernels; only the orchestration differs from the fused GCN / GAT engines)
            return self.train_autograd(model, data, optimizer, args, logits_ori if dense else None)
        return self.train_edge_form(model, data, optimizer, args, logits_ori if dense else None)

    def _loss_mix(self):
        """Weight of ``loss_e`` in the objective (gnndelete.py:390-398): the ablation variants keep one term only;
        otherwise the hard-coded 0.5 / 0.5 mix (``args.alpha`` is ignored there)."""
        unl = getattr(self.args, 'unlearning_model', '') or ''
        if 'ablation_random' in unl:
            return 1.0
        if 'ablation_locality' in unl:
            return 0.0
        return 0.5

    def train_minibatch(self, model, data, optimizer, args, logits_ori=None, attack_model_all=None, attack_model_sub=None):
        """``train_minibatch`` (gnndelete.py:311-450) with its GraphSAINT random-walk batches - opt-in through
        ``args.saint_minibatch`` (the default for every dataset is the whole-graph step above, which needs no sampling on
        a 180 GB device).  Per batch (:347-409): forward on the batch's ``sdf`` edges with the batch's node masks passed
        positionally, as many uniform negatives as the batch holds Df entries, ``loss_e = MSE(df logits, negative
        logits)``, edge-form ``loss_l`` on the batch's ``sdf`` edges with ``u < v``, ``0.5 / 0.5`` mix (or the ablation
        variants, :391-398), backward, ``optimizer.step()``.

        Two things differ from the reference, both documented defects there: ``z_ori`` is the base model's embedding on
        the ``dr_mask`` edges (the reference's ``get_embedding`` reads an unset ``data.dtrain_mask``);
        a batch without Df entries (or without NI pairs) contributes 0 for that term instead of ``MSE(empty) = nan``.
        Kept as in the reference unless ``args.saint_global_z_ori`` is set: the NI target indexes the FULL-graph
        ``z_ori`` with the batch-LOCAL node ids (:379-383).  Composed of calls the GPU tests cover one by one (model
        forward with positional masks, ``decode``, MSE, backward); the loop itself has not been run on a B200 yet."""
        from .sampler import GraphSAINTRandomWalkSampler
        F = torch.nn.functional
        dev = torch.device(getattr(args, 'device', 'cuda'))     # the loop is device agnostic; the CUDA models are not
        model = model.to(dev)
        data = data.to(dev)
        ei = data.train_pos_edge_index
        with torch.no_grad():
            z_ori = getattr(data, 'z_ori', None)
            if z_ori is None:
                z_ori = model.get_original_embeddings(data.x, ei[:, data.dr_mask].contiguous())
        data.edge_index = ei                                                             # :331-332
        data.node_id = torch.arange(data.x.shape[0], device=dev)
        gen = torch.Generator(device=dev).manual_seed(getattr(args, 'random_seed', 42))
        loader = GraphSAINTRandomWalkSampler(data, batch_size=args.batch_size, walk_length=2, num_steps=args.num_steps,
                                             generator=gen)
        unl = self.args.unlearning_model
        global_ids = bool(getattr(args, 'saint_global_z_ori', False))
        best_metric = 0
        for epoch in range(args.epochs):
            model.train()
            sums, steps, t0 = torch.zeros(3, device=dev), 0, time.time()
            for batch in loader:
                bei = batch.edge_index
                sdf_edges = bei[:, batch.sdf_mask].contiguous()
                z = model(batch.x, sdf_edges, batch.sdf_node_1hop_mask, batch.sdf_node_2hop_mask)       # :352
                neg_size = int(batch.df_mask.sum())                                                      # :356-360
                zero = z.sum() * 0.0
                if neg_size > 0:
                    neg = torch.randint(0, z.size(0), (2, neg_size), generator=gen, device=dev)
                    df_logits = model.decode(z, bei[:, batch.df_mask].contiguous(), neg)                # :362-363
                    loss_e = F.mse_loss(df_logits[:neg_size], df_logits[neg_size:])
                else:
                    loss_e = zero
                lower = sdf_edges[0] < sdf_edges[1]                                                      # :379-381
                row, col = sdf_edges[0][lower], sdf_edges[1][lower]
                if row.numel() > 0:
                    src_r, src_c = (batch.node_id[row], batch.node_id[col]) if global_ids else (row, col)
                    target = (z_ori[src_r] * z_ori[src_c]).sum(dim=-1)                                   # :383
                    loss_l = F.mse_loss(model.decode(z, torch.stack([row, col])), target)              # :384-386
                else:
                    loss_l = zero
                if 'ablation_random' in unl:                                                             # :390-398
                    loss_l, loss = zero.detach(), loss_e
                elif 'ablation_locality' in unl:
                    loss_e, loss = zero.detach(), loss_l
                else:
                    loss = 0.5 * loss_e + 0.5 * loss_l
                loss.backward()
                optimizer.step()
                optimizer.zero_grad()
                sums += torch.stack([loss.detach(), loss_e.detach(), loss_l.detach()])
                steps += 1
            v = (sums / max(steps, 1)).tolist()
            self.trainer_log['log'].append({'Epoch': epoch, 'train_loss': v[0], 'loss_r': v[1], 'loss_l': v[2],
                                            'train_time': (time.time() - t0) / max(steps, 1)})
            if (epoch + 1) % args.valid_freq == 0:                                                       # :416-443
                valid_loss, dt_auc, dt_aup, df_auc, df_aup, _, _, valid_log = self.eval(model, data, 'val')
                valid_log['epoch'] = epoch
                self.trainer_log['log'].append(valid_log)
                if dt_auc + df_auc > best_metric:
                    best_metric = dt_auc + df_auc
                    torch.save({'model_state': model.state_dict(), 'optimizer_state': optimizer.state_dict()},
                               os.path.join(args.checkpoint_dir, 'model_best.pt'))
        torch.save({'model_state': {k: v.to('cpu') for k, v in model.state_dict().items()},
                    'optimizer_state': optimizer.state_dict()}, os.path.join(args.checkpoint_dir, 'model_final.pt'))
        return model

    def _negatives(self, data, count, generator):
        """Uniform random pairs (``negative_sampling`` is randomised rejection sampling in
        PyG and not reproducible across implementations); a caller that
        needs parity supplies ``data.neg_edge_index``."""
        return torch.randint(0, data.num_nodes, (2, count), generator=generator, device=data.x.device)

    def start(self, model, data, optimizer, args, logits_ori=None):
        """Set up the GCNDelete unlearning run and return its :class:`EdgeFormSession` - the object :meth:`train`
        loops over (one ``session.step()`` per epoch).  Public so that a caller can drive the epochs itself, e.g.
        feed each epoch's negatives from host memory (``bench.py``'s end-to-end leg does)."""
        dev = self._cuda_device(args)
        model = model.to(dev)
        data = data.to(dev)
        n_df = int(data.df_mask.sum())
        gen = torch.Generator(device=dev).manual_seed(getattr(args, 'random_seed', 42))
        fixed_neg = getattr(data, 'neg_edge_index', None)
        neg = fixed_neg if fixed_neg is not None else self._negatives(data, n_df, gen)
        with torch.no_grad():
            z_ori = getattr(data, 'z_ori', None)
            if z_ori is None and logits_ori is None:
                z_ori = model.get_original_embeddings(data.x, data.train_pos_edge_index[:, data.dr_mask])
        # the engine trains exactly the two Del weights with ONE Adam hyper-parameter set (delete_gnn.py:215-241):
        # refuse an optimizer that holds anything else instead of silently ignoring it
        want = {id(model.deletion1.deletion_weight), id(model.deletion2.deletion_weight)}
        have = {id(p) for g in optimizer.param_groups for p in g['params']}
        if have != want:
            raise ValueError('GNNDeleteTrainer expects an optimizer over deletion1/deletion2.deletion_weight only')
        group = optimizer.param_groups[0]
        for g in optimizer.param_groups[1:]:
            if (g['lr'], g['betas'], g['eps']) != (group['lr'], group['betas'], group['eps']):
                raise ValueError('per-group Adam hyper-parameters are not supported by the fused GCNDelete epoch')
        engine_cls = GATDeleteEngine if type(model).__name__ == 'GATDelete' else GCNDeleteEngine
        eng = engine_cls(model, data, neg, z_ori=z_ori, lr=group['lr'], betas=group['betas'], eps=group['eps'],
                              hoist_layer1=getattr(args, 'hoist_layer1', True), logits_ori=logits_ori,
                              static_negatives=fixed_neg is not None, alpha=self._loss_mix())
        self.engine = eng
        if getattr(args, 'capture_step', True):
            # one cudaGraphLaunch per epoch (edge-form and dense-block NI alike); resampled negatives are written into the
            # graph's staging buffer
            try:
                eng.capture(warmup=2, dynamic_negatives=fixed_neg is None)
            except Exception as exc:                            # capture is an optimisation: fall back to eager epochs
                self.trainer_log['capture_error'] = repr(exc)
                eng.graph = None
                torch.cuda.synchronize()
        self.trainer_log['captured_step'] = eng.graph is not None
        return EdgeFormSession(self, model, data, eng, gen, n_df, resample=fixed_neg is None)

    def train_edge_form(self, model, data, optimizer, args, logits_ori=None):
        sess = self.start(model, data, optimizer, args, logits_ori)
        model, data, eng = sess.model, sess.data, sess.engine
        best_metric = 0
        ring = []
        t0 = time.time()
        for epoch in range(args.epochs):
            model.train()
            losses = sess.step()
            ring.append(losses.clone())
            last = epoch + 1 == args.epochs
            if (epoch + 1) % self.log_every == 0 or last or (epoch + 1) % args.valid_freq == 0:
                vals = torch.stack(ring).cpu()
                dt = (time.time() - t0) / len(ring)
                for i, v in enumerate(vals.tolist()):
                    self.trainer_log['log'].append({'Epoch': epoch + 1 - len(ring) + i, 'train_loss': v[0],
                                                    'loss_r': v[1], 'loss_l': v[2], 'train_time': dt})
                ring, t0 = [], time.time()
            if (epoch + 1) % args.valid_freq == 0:
                valid_loss, dt_auc, dt_aup, df_auc, df_aup, _, _, valid_log = self.eval(model, data, 'val')
                valid_log['epoch'] = epoch
                self.trainer_log['log'].append(valid_log)
                if dt_auc + df_auc > best_metric:
                    best_metric = dt_auc + df_auc
                    torch.save({'model_state': model.state_dict(), 'optimizer_state': self._optimizer_state(optimizer, eng)},
                               os.path.join(args.checkpoint_dir, 'model_best.pt'))
        torch.save({'model_state': {k: v.to('cpu') for k, v in model.state_dict().items()},
                    'optimizer_state': self._optimizer_state(optimizer, eng)},
                   os.path.join(args.checkpoint_dir, 'model_final.pt'))
        return model

    @staticmethod
    def _cuda_device(args):
        dev = torch.device(getattr(args, 'device', None) or 'cuda')
        if dev.type != 'cuda':
            raise RuntimeError('gnndelete_b200 has no CPU path: args.device must be a CUDA device')
        return dev

    def train_autograd(self, model, data, optimizer, args, logits_ori=None):
        """``train_minibatch``'s step body (gnndelete.py:347-409) on the whole graph for any *Delete model:
        forward with positional masks (:352), fused decode + DEC + edge-form NI (`EdgeLossFn`), backward
        through the conv / Del autograd Functions, ``optimizer.step()`` (the caller's optimizer)."""
        from .losses import DenseNIPlan, EdgeLossPlan, dense_ni_loss, edge_loss
        dev = self._cuda_device(args)
        model = model.to(dev)
        data = data.to(dev)
        ei = data.train_pos_edge_index
        n_df = int(data.df_mask.sum())
        gen = torch.Generator(device=dev).manual_seed(getattr(args, 'random_seed', 42))
        fixed_neg = getattr(data, 'neg_edge_index', None)
        neg = fixed_neg if fixed_neg is not None else self._negatives(data, n_df, gen)
        ei_sdf = ei[:, data.sdf_mask].contiguous()
        with torch.no_grad():
            z_ori = getattr(data, 'z_ori', None)
            if z_ori is None:
                z_ori = model.get_original_embeddings(data.x, ei[:, data.dr_mask].contiguous())
        ni = ei_sdf[:, ei_sdf[0] < ei_sdf[1]]                                   # gnndelete.py:379-381
        alpha = self._loss_mix()
        dense = None
        if logits_ori is not None:
            # train_fullbatch's NI term (gnndelete.py:163-193, 239-241): the dense S2 x S2 sigmoid block against the
            # original model's pair logits, for every architecture - the edge-form pairs are dropped
            dense = DenseNIPlan(data.sdf_node_2hop_mask, ei[:, data.df_mask], logits_ori, data.num_nodes,
                                model.deletion2.dim, weight=1.0 - alpha)
            ni = ni[:, :0]
        plan = EdgeLossPlan(ei[:, data.df_mask], neg, ni, data.num_nodes, z_ori=z_ori.contiguous(),
                            alpha=alpha, static_negatives=fixed_neg is not None)
        # The step (forward, fused loss, backward through the autograd Functions, Adam on the Del weights) is
        # captured into ONE CUDA graph: on PubMed-sized graphs the ~40 launches of a step are host-launch bound
        # when issued from Python.  Adam runs through gd_adam_step (device-side step counter) on the optimizer's
        # hyper-parameters; its moments are mirrored into ``optimizer.state`` when checkpoints are written.
        params = [p for g in optimizer.param_groups for p in g['params']]
        group = optimizer.param_groups[0]
        state = [dict(m=torch.zeros_like(p), v=torch.zeros_like(p), step=torch.zeros(1, dtype=torch.float32, device=dev))
                 for p in params]
        out_static = torch.zeros(3, dtype=torch.float32, device=dev)

        def step():
            z = model(data.x, ei_sdf, data.sdf_node_1hop_mask, data.sdf_node_2hop_mask)       # :352
            loss, loss_r, loss_l = edge_loss(z, plan)
            if dense is not None:                         # loss = alpha * loss_r  (no NI pairs)  + (1 - alpha) * dense NI
                wl, loss_l = dense_ni_loss(z, dense)
                loss = loss + wl
            loss.backward()
            for p, st in zip(params, state):
                ops.adam_step(p.data, p.grad, st['m'], st['v'], st['step'], group['lr'], group['betas'][0],
                              group['betas'][1], group['eps'])
            out_static.copy_(torch.stack([loss.detach(), loss_r, loss_l]))

        graph = None
        if getattr(args, 'capture_step', True):
            snap = [p.detach().clone() for p in params]
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(2):                          # warm-up: plans, workspaces, module loading
                        for p in params:
                            p.grad = None
                        step()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                for p in params:
                    p.grad = None
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    step()
                graph = g
            except Exception as exc:                            # capture is an optimisation: fall back to eager steps
                self.trainer_log['capture_error'] = repr(exc)
                graph = None
                torch.cuda.synchronize()
            with torch.no_grad():                               # undo the warm-up / capture steps
                for p, s0 in zip(params, snap):
                    p.copy_(s0)
                for st in state:
                    for v in st.values():
                        v.zero_()
        self.trainer_log['captured_step'] = graph is not None
        best_metric, ring, t0 = 0, [], time.time()
        for epoch in range(args.epochs):
            model.train()
            if fixed_neg is None and epoch > 0:
                plan.update_negatives(self._negatives(data, n_df, gen))
            if graph is not None:
                graph.replay()
            else:
                for p in params:
                    p.grad = None
                step()
            loss, loss_r, loss_l = out_static.clone().unbind(0)
            ring.append(torch.stack([loss.detach(), loss_r, loss_l]))
            last = epoch + 1 == args.epochs
            if (epoch + 1) % self.log_every == 0 or last or (epoch + 1) % args.valid_freq == 0:
                vals = torch.stack(ring).cpu()
                dt = (time.time() - t0) / len(ring)
                for i, v in enumerate(vals.tolist()):
                    self.trainer_log['log'].append({'Epoch': epoch + 1 - len(ring) + i, 'train_loss': v[0],
                                                    'loss_r': v[1], 'loss_l': v[2], 'train_time': dt})
                ring, t0 = [], time.time()
            if (epoch + 1) % args.valid_freq == 0:
                valid_loss, dt_auc, dt_aup, df_auc, df_aup, _, _, valid_log = self.eval(model, data, 'val')
                valid_log['epoch'] = epoch
                self.trainer_log['log'].append(valid_log)
                if dt_auc + df_auc > best_metric:
                    best_metric = dt_auc + df_auc
                    torch.save({'model_state': model.state_dict(), 'optimizer_state': self._mirror_adam(optimizer, params, state)},
                               os.path.join(args.checkpoint_dir, 'model_best.pt'))
        torch.save({'model_state': {k: v.to('cpu') for k, v in model.state_dict().items()},
                    'optimizer_state': self._mirror_adam(optimizer, params, state)},
                   os.path.join(args.checkpoint_dir, 'model_final.pt'))
        return model

    @staticmethod
    def _mirror_adam(optimizer, params, state):
        for p, st in zip(params, state):
            optimizer.state[p] = {'step': st['step'].detach().cpu().reshape(()).clone(), 'exp_avg': st['m'],
                                  'exp_avg_sq': st['v']}
        return optimizer.state_dict()

    @staticmethod
    def _optimizer_state(optimizer, eng):
        """Mirror the engine's Adam moments into the caller's ``torch.optim.Adam`` so the
        saved ``optimizer_state`` stays loadable by the reference."""
        params = [p for g in optimizer.param_groups for p in g['params']]
        for p, st in zip(eng.params, eng.state):
            for q in params:
                if q is p:
                    optimizer.state[q] = {'step': st['step'].detach().cpu().reshape(()).clone(),
                                          'exp_avg': st['m'], 'exp_avg_sq': st['v']}
        return optimizer.state_dict()


class EdgeFormSession:
    """One GCNDelete unlearning run in progress (``GNNDeleteTrainer.start``): the fused epoch engine, captured into
    a CUDA graph when possible, plus the negative-sampling state.  ``step()`` = one epoch of ``train`` (gnndelete.py:
    215-259 with the edge-form NI of :379-386): new negatives -> forward -> losses -> backward -> Adam."""

    def __init__(self, trainer, model, data, engine, generator, n_df, resample):
        self.trainer, self.model, self.data, self.engine = trainer, model, data, engine
        self.gen, self.n_df, self.resample = generator, n_df, resample
        self.epochs_done = 0
        self._pipe = None

    def step(self, negatives=None):
        """One epoch; returns the persistent device tensor (loss, loss_r, loss_l) - no host sync.  ``negatives``:
        this epoch's ``[2, n_df]`` negative edges (host or device; a pinned host tensor is copied asynchronously);
        default: resampled on the device like the reference does every epoch (gnndelete.py:221-225), or the fixed
        ``data.neg_edge_index`` when one was supplied."""
        if negatives is not None:
            self.engine.set_negatives(negatives)
        elif self.resample and self.epochs_done > 0:
            self.engine.set_negatives(self.trainer._negatives(self.data, self.n_df, self.gen))
        self.epochs_done += 1
        return self.engine.epoch()

    # pipelined host feeding: epoch k's negatives upload and epoch k-1's losses download while a neighbour computes
    def submit(self, neg_host):
        """Queue one epoch on pinned host negatives; returns its index for :meth:`result`."""
        if self._pipe is None:
            from .engine import EpochPipeline
            self._pipe = EpochPipeline(self.engine)
        self.epochs_done += 1
        return self._pipe.submit(neg_host)

    def result(self, k):
        """Host (loss, loss_r, loss_l) of epoch ``k`` (one of the last two submitted); blocks until it has arrived."""
        return self._pipe.result(k)


class KGTrainer(Trainer):
    """Evaluation side of ``framework/trainer/base.py:394-692`` (``KGTrainer``) for the RGCN models: DistMult decode
    of the val / test triples on the ``dr_mask`` message-passing edges, BCE + AUC / AP on Dt - on the raw logits,
    the KG variant applies no sigmoid there (:508-514) - and AUC / AP of the deleted triples against
    ``num_df_resamples`` random samples of retained forward triples (:517-540).  KG training of the original model
    (:395-492, GraphSAINT mini-batches, relation-weight gradients) is outside the accelerated path."""

    def train(self, model, data, optimizer, args, logits_ori=None, attack_model_all=None, attack_model_sub=None):
        raise NotImplementedError('training the original RGCN model (KGTrainer.train, base.py:395-492) is outside the '
                                  'accelerated hot path')

    @torch.no_grad()
    def eval(self, model, data, stage='val', pred_all=False, num_df_resamples=500):
        model.eval()
        pos, neg = data[f'{stage}_pos_edge_index'], data[f'{stage}_neg_edge_index']
        et = data[f'{stage}_edge_type']
        z = model(data.x, data.edge_index[:, data.dr_mask].contiguous(), data.edge_type[data.dr_mask].contiguous())
        logits = model.decode(z, torch.cat([pos, neg], dim=-1), torch.cat([et, et], dim=-1))
        label = self.get_link_labels(pos, neg)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, label).item()
        dt_auc = metrics.roc_auc(label, logits)
        dt_aup = metrics.average_precision(label, logits)
        if self.args.unlearning_model in ['original']:
            df_logit = torch.empty(0, device=z.device)
        else:
            df_logit = model.decode(z, data.directed_df_edge_index, data.directed_df_edge_type).sigmoid()
        n_df = df_logit.numel()
        if n_df > 0:
            dr_mask = data.dr_mask[:data.dr_mask.shape[0] // 2]                      # forward-direction triples
            dr_logit = model.decode(z, data.train_pos_edge_index[:, dr_mask].contiguous(),
                                    data.train_edge_type[dr_mask].contiguous()).sigmoid()
            if len(self.df_pos_edge) == 0:       # the reference redraws the 500 samples on every call; cached here
                g = torch.Generator(device='cpu').manual_seed(getattr(self.args, 'random_seed', 42))
                self.df_pos_edge = torch.stack([torch.randperm(dr_logit.numel(), generator=g)[:n_df]
                                                for _ in range(num_df_resamples)]).to(z.device)
            aucs, aups = metrics.resampled_auc_ap(df_logit, dr_logit, self.df_pos_edge)
            df_auc, df_aup = aucs.mean().item(), aups.mean().item()
        else:
            df_auc = df_aup = float('nan')
        logit_all_pair = (z @ z.t()).cpu() if pred_all else None
        log = {
            f'{stage}_loss': loss, f'{stage}_dt_auc': dt_auc, f'{stage}_dt_aup': dt_aup,
            f'{stage}_df_auc': df_auc, f'{stage}_df_aup': df_aup,
            f'{stage}_df_logit_mean': df_logit.mean().item() if n_df else float('nan'),
            f'{stage}_df_logit_std': df_logit.std(unbiased=False).item() if n_df else float('nan'),
        }
        return loss, dt_auc, dt_aup, df_auc, df_aup, df_logit.tolist(), logit_all_pair, log

    @torch.no_grad()
    def test(self, model, data, model_retrain=None, attack_model_all=None, attack_model_sub=None, ckpt='ckpt'):
        """``base.py:570-640``; the default ``ckpt='ckpt'`` never reloads the best checkpoint (the reference compares
        ``ckpt is 'best'``) - kept."""
        return super().test(model, data, ckpt=ckpt)


class KGGNNDeleteNodeembTrainer(KGTrainer):
    """``framework/trainer/gnndelete_nodeemb.py:659-846`` (``KGGNNDeleteNodeembTrainer``),
    the route ``--gnn rgcn --unlearning_model gnndelete[_nodeemb]`` dispatches to.

    Per step (:749-798): forward on the ``dr_mask`` edges with the ``*_non_df`` node masks,
    original embeddings under ``no_grad``, ``negative_sampling_kg`` on the forward-direction
    Df triples, per-layer node-embedding MSEs, then TWO backward passes and TWO Adam steps
    (``optimizer`` is the ``[optimizer1, optimizer2]`` pair of delete_gnn.py:221-226).
    The schedule is reproduced literally — including that ``loss2.backward()`` leaves its
    deletion1 gradient in ``.grad`` until the next step's ``optimizer[0].step()``.  The whole
    graph is the batch (no GraphSAINT sampling); the frozen original embeddings are
    computed once instead of once per step."""

    log_every = 10

    def start(self, model, data, optimizer, args):
        """Set up the KG unlearning run and return its :class:`KGNodeembSession` (one ``session.step()`` per epoch)."""
        from .kg import negative_sampling_kg
        from .losses import RowMSEPlan
        if not isinstance(optimizer, (list, tuple)) or len(optimizer) != 2:
            raise ValueError('expects the [optimizer1, optimizer2] pair built for *layerwise loss types '
                             '(delete_gnn.py:221-226)')
        dev = torch.device(getattr(args, 'device', None) or 'cuda')
        model = model.to(dev)
        data = data.to(dev)
        alpha = args.alpha
        non_df = torch.ones(data.x.shape[0], dtype=torch.bool, device=dev)          # :723-727
        non_df[data.directed_df_edge_index.flatten().unique()] = False
        m1 = data.sdf_node_1hop_mask & non_df
        m2 = data.sdf_node_2hop_mask & non_df
        data.sdf_node_1hop_mask_non_df_mask, data.sdf_node_2hop_mask_non_df_mask = m1, m2
        edge_index = data.edge_index[:, data.dr_mask].contiguous()                  # :749-750
        edge_type = data.edge_type[data.dr_mask].contiguous()
        pos_ei = data.edge_index[:, data.df_mask]                                    # :758-763
        pos_et = data.edge_type[data.df_mask]
        dec = pos_et < args.num_edge_type
        dec_ei, dec_et = pos_ei[:, dec], pos_et[dec]
        with torch.no_grad():                                                        # :754-755
            z1o, z2o = model.get_original_embeddings(data.x, edge_index, edge_type, return_all_emb=True)
        gen = torch.Generator(device=dev).manual_seed(getattr(args, 'random_seed', 42))
        fixed_neg = getattr(data, 'neg_edge_index', None)
        # per layer, the four row gathers + two MSEs + their scatter gradients (:770-781) are one kernel pass
        neg = fixed_neg if fixed_neg is not None else negative_sampling_kg(dec_ei, dec_et, gen)
        plan1 = RowMSEPlan(dec_ei, neg, m1, z1o, mix=(alpha, 1 - alpha))
        plan2 = RowMSEPlan(dec_ei, neg, m2, z2o, mix=(alpha, 1 - alpha))
        sess = KGNodeembSession(self, model, data, optimizer, edge_index, edge_type, m1, m2, plan1, plan2, dec_ei, dec_et,
                                gen, resample=fixed_neg is None)
        if getattr(args, 'capture_step', True) and fixed_neg is not None:
            sess.capture()
        self.trainer_log['captured_step'] = sess.graph is not None
        return sess

    def train(self, model, data, optimizer, args, logits_ori=None, attack_model_all=None, attack_model_sub=None):
        sess = self.start(model, data, optimizer, args)
        model, data = sess.model, sess.data
        ring, best_metric = [], 0
        for epoch in range(args.epochs):
            model.train()
            ring.append(sess.step().clone())
            if (epoch + 1) % self.log_every == 0 or epoch + 1 == args.epochs:
                for i, v in enumerate(torch.stack(ring).cpu().tolist()):
                    self.trainer_log['log'].append({'Epoch': epoch + 1 - len(ring) + i, 'train_loss': v[0],
                                                    'loss_r': v[1], 'loss_l': v[2]})
                ring = []
            if (epoch + 1) % args.valid_freq == 0:                                   # :815-841
                valid_loss, dt_auc, dt_aup, df_auc, df_aup, _, _, valid_log = self.eval(model, data, 'val')
                valid_log['epoch'] = epoch
                self.trainer_log['log'].append(valid_log)
                if dt_auc + df_auc > best_metric:
                    best_metric = dt_auc + df_auc
                    sess.mirror_optimizers()
                    torch.save({'model_state': model.state_dict()}, os.path.join(args.checkpoint_dir, 'model_best.pt'))
        sess.mirror_optimizers()
        torch.save({'model_state': {k: v.to('cpu') for k, v in model.state_dict().items()}},
                   os.path.join(args.checkpoint_dir, 'model_final.pt'))
        return model


class KGNodeembSession:
    """One KG unlearning run in progress (``KGGNNDeleteNodeembTrainer.start``).  ``step()`` is the reference's step
    body (gnndelete_nodeemb.py:744-798) with its literal schedule: ``loss1.backward`` -> Adam on deletion1 ->
    ``loss2.backward`` -> Adam on deletion2, where ``loss2.backward`` also leaves a deletion1 gradient that stays in
    ``.grad`` until the next step's first Adam update.  Adam runs through ``gd_adam_step`` on the hyper-parameters of the
    caller's two optimizers (device-side step counters, so the whole step is capturable into ONE CUDA graph when the
    negatives are fixed); the moments are mirrored into the optimizers when checkpoints are written."""

    def __init__(self, trainer, model, data, optimizers, edge_index, edge_type, m1, m2, plan1, plan2, dec_ei, dec_et,
                 generator, resample):
        self.trainer, self.model, self.data, self.optimizers = trainer, model, data, list(optimizers)
        self.edge_index, self.edge_type, self.m1, self.m2 = edge_index, edge_type, m1, m2
        self.plan1, self.plan2, self.dec_ei, self.dec_et = plan1, plan2, dec_ei, dec_et
        self.gen, self.resample = generator, resample
        dev = data.x.device
        self.groups = []
        for opt in self.optimizers:
            g = opt.param_groups[0]
            ps = [p for grp in opt.param_groups for p in grp['params']]
            st = [dict(m=torch.zeros_like(p), v=torch.zeros_like(p), step=torch.zeros(1, dtype=torch.float32, device=dev)) for p in ps]
            for p in ps:
                p.grad = torch.zeros_like(p)          # kept as tensors (zeroed, never None): static addresses for the graph
            self.groups.append((g, ps, st))
        self.out = torch.zeros(3, dtype=torch.float32, device=dev)
        self.graph = None
        self.epochs_done = 0

    def _adam(self, k):
        g, ps, st = self.groups[k]
        for p, s_ in zip(ps, st):
            ops.adam_step(p.data, p.grad, s_['m'], s_['v'], s_['step'], g['lr'], g['betas'][0], g['betas'][1], g['eps'])
            p.grad.zero_()                            # optimizer.zero_grad() of the schedule

    def _body(self):
        from .losses import row_mse
        z1, z2 = self.model(self.data.x, self.edge_index, self.edge_type, self.m1, self.m2, return_all_emb=True)
        loss1, loss_r1, loss_l1 = row_mse(z1, self.plan1)                        # :788-796
        loss2, loss_r2, loss_l2 = row_mse(z2, self.plan2)
        loss1.backward(retain_graph=True)
        self._adam(0)
        loss2.backward(retain_graph=True)
        self._adam(1)
        self.out.copy_(torch.stack([(loss1 + loss2).detach(), loss_r1 + loss_r2, loss_l1 + loss_l2]))

    def capture(self, warmup=2):
        """Capture one step into a CUDA graph (fixed negatives only).  Warm-up steps run first and are undone."""
        params = [p for _, ps, _ in self.groups for p in ps]
        snap_p = [p.detach().clone() for p in params]
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    self._body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._body()
            self.graph = g
        except Exception as exc:                      # capture is an optimisation: fall back to eager steps
            self.trainer.trainer_log['capture_error'] = repr(exc)
            self.graph = None
            torch.cuda.synchronize()
        with torch.no_grad():                         # undo the warm-up / capture steps
            for p, s0 in zip(params, snap_p):
                p.copy_(s0)
                p.grad.zero_()
            for _, _, st in self.groups:
                for s_ in st:
                    for v in s_.values():
                        v.zero_()

    def step(self, negatives=None):
        """One step; returns the persistent device tensor (loss1 + loss2, loss_r1 + loss_r2, loss_l1 + loss_l2).
        ``negatives``: this step's corrupted triples ``[2, n]`` (host or device); rebuilds the loss incidence, so it is
        only accepted by a session that was not captured."""
        if negatives is not None:
            if self.graph is not None:
                raise RuntimeError('per-step negatives need an uncaptured session (args.capture_step = False)')
            neg = negatives.to(self.dec_ei.device, non_blocking=True)
            self.plan1.set_pairs(self.dec_ei, neg)
            self.plan2.set_pairs(self.dec_ei, neg)
        elif self.resample and self.epochs_done > 0:                             # :764-768, new heads every step
            from .kg import negative_sampling_kg
            neg = negative_sampling_kg(self.dec_ei, self.dec_et, self.gen)
            self.plan1.set_pairs(self.dec_ei, neg)
            self.plan2.set_pairs(self.dec_ei, neg)
        self.epochs_done += 1
        if self.graph is not None:
            self.graph.replay()
        else:
            self._body()
        return self.out

    def mirror_optimizers(self):
        for opt, (_, ps, st) in zip(self.optimizers, self.groups):
            for p, s_ in zip(ps, st):
                opt.state[p] = {'step': s_['step'].detach().cpu().reshape(()).clone(), 'exp_avg': s_['m'], 'exp_avg_sq': s_['v']}


# ---- node-embedding loss functions of gnndelete_nodeemb.py:19-92 (the non-MSE members run as device tensor ops
# ---- under autograd; the two MSE members go through the fused gd_row_mse_fwd_bwd kernel instead)
def _bounded_kld(reduction):
    def fct(logits, truth):
        kld = torch.nn.functional.kl_div(torch.log_softmax(logits, -1), torch.softmax(truth, -1), reduction=reduction)
        return 1 - torch.exp(-kld)
    return fct


def _cosine_distance(reduce):
    def fct(logits, truth):
        d = 1 - torch.nn.functional.cosine_similarity(logits, truth)
        return d.mean() if reduce == 'mean' else d.sum()
    return fct


def _centered(k):
    # H K H with H = I - 11^T / n, without forming H
    return k - k.mean(0, keepdim=True) - k.mean(1, keepdim=True) + k.mean()


def _gram_linear(x):
    return x @ x.t()


def _gram_rbf(x, sigma=None):
    g = x @ x.t()
    d = torch.diag(g)
    k = d.unsqueeze(0) + d.unsqueeze(1) - 2 * g
    if sigma is None:
        sigma = torch.sqrt(torch.median(k[k != 0]).detach())      # the reference takes math.sqrt of it: a constant
    return torch.exp(-0.5 * k / (sigma * sigma))


def _cka(gram):
    def hsic(a, b):
        return (_centered(gram(a)) * _centered(gram(b))).sum()

    def fct(x, y):
        return hsic(x, y) / (torch.sqrt(hsic(x, x)) * torch.sqrt(hsic(y, y)))
    return fct


def get_nodeemb_loss_fct(name):
    """``gnndelete_nodeemb.py:69-92``.  Returns ``('mse', reduction)`` for the members the fused kernel covers,
    else a callable ``(logits, truth) -> scalar`` (``rbf_cka`` raises ``NameError`` in the reference - ``math`` is
    never imported, SURVEY.md §10 #13 - and is implemented as written otherwise)."""
    if name in ('mse_mean', 'mse_sum'):
        return ('mse', name[4:])
    table = {
        'kld_mean': _bounded_kld('batchmean'), 'kld_sum': _bounded_kld('sum'),
        'cosine_mean': _cosine_distance('mean'), 'cosine_sum': _cosine_distance('sum'),
        'linear_cka': _cka(_gram_linear), 'rbf_cka': _cka(_gram_rbf),
    }
    if name not in table:
        raise NotImplementedError(name)
    return table[name]


class GNNDeleteNodeembTrainer(Trainer):
    """``framework/trainer/gnndelete_nodeemb.py:94-349`` (``GNNDeleteNodeembTrainer``): the layer-wise
    Deleted-Edge-Consistency / Neighbourhood-Influence objective on NODE EMBEDDINGS, the route
    ``--unlearning_model gnndelete_nodeemb`` takes and the default ``loss_type`` of ``training_args.py:63-66``.

    Per epoch (:191-299): ``z1, z2 = model(x, ei[:, sdf_mask], return_all_emb=True)`` with the stored masks;
    ``loss_r{l} = fct(cat(z{l}[Df heads], z{l}[Df tails]), cat(z{l}_ori[neg heads], z{l}_ori[neg tails]))``;
    ``loss_l{l} = fct(z{l}[S{l} minus Df nodes], z{l}_ori[same rows])``; then the backward / optimizer schedule of
    ``args.loss_type`` reproduced literally (``both_all`` never clears gradients; the ``*layerwise`` types take
    the ``[optimizer1, optimizer2]`` pair of delete_gnn.py:221-226).  With ``loss_fct`` ``mse_mean`` / ``mse_sum``
    the four gathers + MSEs + their scatter gradients of a layer are ONE kernel pass (``losses.RowMSEPlan``).
    The reference's 'ogbl' mini-batch variant (:352-520, GraphSAINT sampling) is replaced by the same whole-graph
    step, as in ``GNNDeleteTrainer``.  Negatives are drawn once before the loop (:186-189); a caller that needs
    parity supplies ``data.neg_edge_index``."""

    log_every = 100

    def train(self, model, data, optimizer, args, logits_ori=None, attack_model_all=None, attack_model_sub=None):
        if getattr(args, 'saint_minibatch', False):
            return self.train_minibatch(model, data, optimizer, args)
        return self.train_fullbatch(model, data, optimizer, args, logits_ori, attack_model_all, attack_model_sub)

    def train_minibatch(self, model, data, optimizer, args, logits_ori=None, attack_model_all=None, attack_model_sub=None):
        """``GNNDeleteNodeembTrainer.train_minibatch`` (gnndelete_nodeemb.py:352-494), opt-in through
        ``args.saint_minibatch``: GraphSAINT random-walk batches; per batch the original embeddings of the BATCH graph
        (all of its edges, :397-399), the Del forward on the batch's ``sdf`` edges with the batch's masks, uniform
        negatives, the four node-embedding MSEs and - always - the ``both_layerwise`` schedule with the
        ``[optimizer1, optimizer2]`` pair (:431-441).  A term over an empty selection contributes 0 (nan in the
        reference).  Device tensor ops under autograd; host loop tested on the CPU with the oracle's models, not yet
        run on a B200."""
        from .sampler import GraphSAINTRandomWalkSampler
        if not isinstance(optimizer, (list, tuple)) or len(optimizer) != 2:
            raise ValueError('the mini-batch loop steps the [optimizer1, optimizer2] pair (gnndelete_nodeemb.py:431-441)')
        F = torch.nn.functional
        alpha = self.args.alpha
        dev = torch.device(getattr(args, 'device', 'cuda'))
        model = model.to(dev)
        data = data.to(dev)
        non_df = torch.ones(data.x.shape[0], dtype=torch.bool, device=dev)                  # :369-373
        non_df[data.directed_df_edge_index.flatten().unique()] = False
        data.sdf_node_1hop_mask_non_df_mask = data.sdf_node_1hop_mask & non_df
        data.sdf_node_2hop_mask_non_df_mask = data.sdf_node_2hop_mask & non_df
        data.edge_index = data.train_pos_edge_index                                          # :376-377
        data.node_id = torch.arange(data.x.shape[0], device=dev)
        gen = torch.Generator(device=dev).manual_seed(getattr(args, 'random_seed', 42))
        loader = GraphSAINTRandomWalkSampler(data, batch_size=args.batch_size, walk_length=2, num_steps=args.num_steps,
                                             generator=gen)

        def mse(a, b):
            return F.mse_loss(a, b) if a.numel() else a.sum() * 0.0

        best_metric = 0
        for epoch in range(args.epochs):
            model.train()
            sums, steps, t0 = torch.zeros(3, device=dev), 0, time.time()
            for batch in loader:
                bei = batch.edge_index
                with torch.no_grad():                                                        # :397-399
                    z1_ori, z2_ori = model.get_original_embeddings(batch.x, bei, return_all_emb=True)
                z1, z2 = model(batch.x, bei[:, batch.sdf_mask].contiguous(), batch.sdf_node_1hop_mask,
                               batch.sdf_node_2hop_mask, return_all_emb=True)               # :403
                pos_edge = bei[:, batch.df_mask]                                             # :406-411
                neg_edge = torch.randint(0, batch.x.shape[0], (2, pos_edge.shape[1]), generator=gen, device=dev)
                m1, m2 = batch.sdf_node_1hop_mask_non_df_mask, batch.sdf_node_2hop_mask_non_df_mask
                loss_r1 = mse(torch.cat([z1[pos_edge[0]], z1[pos_edge[1]]]), torch.cat([z1_ori[neg_edge[0]], z1_ori[neg_edge[1]]]))
                loss_r2 = mse(torch.cat([z2[pos_edge[0]], z2[pos_edge[1]]]), torch.cat([z2_ori[neg_edge[0]], z2_ori[neg_edge[1]]]))
                loss_l1 = mse(z1[m1], z1_ori[m1])                                            # :424-425
                loss_l2 = mse(z2[m2], z2_ori[m2])
                loss1 = alpha * loss_r1 + (1 - alpha) * loss_l1                              # :431-441
                loss1.backward(retain_graph=True)
                optimizer[0].step()
                optimizer[0].zero_grad()
                loss2 = alpha * loss_r2 + (1 - alpha) * loss_l2
                loss2.backward(retain_graph=True)
                optimizer[1].step()
                optimizer[1].zero_grad()
                sums += torch.stack([(loss1 + loss2).detach(), (loss_r1 + loss_r2).detach(), (loss_l1 + loss_l2).detach()])
                steps += 1
                del z1, z2, loss1, loss2
            v = (sums / max(steps, 1)).tolist()
            self.trainer_log['log'].append({'epoch': epoch, 'train_loss': v[0], 'train_loss_r': v[1], 'train_loss_l': v[2],
                                            'train_time': (time.time() - t0) / max(steps, 1)})
            if (epoch + 1) % args.valid_freq == 0:                                          # :464-487
                valid_loss, dt_auc, dt_aup, df_auc, df_aup, _, _, valid_log = self.eval(model, data, 'val')
                valid_log['epoch'] = epoch
                self.trainer_log['log'].append(valid_log)
                if dt_auc + df_auc > best_metric:
                    best_metric = dt_auc + df_auc
                    torch.save({'model_state': model.state_dict()}, os.path.join(args.checkpoint_dir, 'model_best.pt'))
        torch.save({'model_state': {k: v.to('cpu') for k, v in model.state_dict().items()}},
                   os.path.join(args.checkpoint_dir, 'model_final.pt'))
        return model

    def train_fullbatch(self, model, data, optimizer, args, logits_ori=None, attack_model_all=None, attack_model_sub=None):
        from .losses import RowMSEPlan, row_mse
        loss_type = self.args.loss_type
        if loss_type not in ('both_all', 'both_layerwise', 'only2_layerwise', 'only2_all', 'only1'):
            raise NotImplementedError(loss_type)
        layerwise = 'layerwise' in loss_type
        if layerwise != isinstance(optimizer, (list, tuple)) or (layerwise and len(optimizer) != 2):
            raise ValueError("'*layerwise' loss types take the [optimizer1, optimizer2] pair, the others one optimizer "
                             '(delete_gnn.py:221-226)')
        fct = get_nodeemb_loss_fct(self.args.loss_fct)
        alpha = self.args.alpha
        dev = torch.device(getattr(args, 'device', 'cuda'))       # host loop is device agnostic; the CUDA models are not
        model = model.to(dev)
        data = data.to(dev)
        ei = data.train_pos_edge_index
        non_df = torch.ones(data.x.shape[0], dtype=torch.bool, device=dev)             # :171-175
        non_df[data.directed_df_edge_index.flatten().unique()] = False
        m1 = data.sdf_node_1hop_mask & non_df
        m2 = data.sdf_node_2hop_mask & non_df
        data.sdf_node_1hop_mask_non_df_mask, data.sdf_node_2hop_mask_non_df_mask = m1, m2
        with torch.no_grad():                                                           # :179-180
            z1_ori, z2_ori = model.get_original_embeddings(data.x, ei[:, data.dr_mask].contiguous(), return_all_emb=True)
        pos_edge = ei[:, data.df_mask]                                                  # :200
        neg_edge = getattr(data, 'neg_edge_index', None)                                # :186-189
        if neg_edge is None:
            gen = torch.Generator(device=dev).manual_seed(getattr(args, 'random_seed', 42))
            neg_edge = torch.randint(0, data.num_nodes, (2, pos_edge.shape[1]), generator=gen, device=dev)
        ei_sdf = ei[:, data.sdf_mask].contiguous()
        # weights of (loss_r, loss_l) in the objective that is differentiated, per layer
        mix = (alpha, 1.0) if loss_type in ('only2_all', 'only1') else (alpha, 1 - alpha)
        if isinstance(fct, tuple):
            plan1 = RowMSEPlan(pos_edge, neg_edge, m1, z1_ori, mix=mix, reduction=fct[1])
            plan2 = RowMSEPlan(pos_edge, neg_edge, m2, z2_ori, mix=mix, reduction=fct[1])

            def layer_losses(z, z_ori, mask, plan):
                return row_mse(z, plan)
        else:
            plan1 = plan2 = None

            def layer_losses(z, z_ori, mask, plan):
                embed = torch.cat([z[pos_edge[0]], z[pos_edge[1]]], dim=0)             # :203-207
                embed_ori = torch.cat([z_ori[neg_edge[0]], z_ori[neg_edge[1]]], dim=0)
                loss_r = fct(embed, embed_ori)                                          # :209-210
                loss_l = fct(z[mask], z_ori[mask])                                      # :213-214
                return mix[0] * loss_r + mix[1] * loss_l, loss_r.detach(), loss_l.detach()

        best_metric, ring, t0 = 0, [], time.time()
        for epoch in range(args.epochs):
            model.train()
            z1, z2 = model(data.x, ei_sdf, return_all_emb=True)                         # :195
            obj1, loss_r1, loss_l1 = layer_losses(z1, z1_ori, m1, plan1)
            obj2, loss_r2, loss_l2 = layer_losses(z2, z2_ori, m2, plan2)
            if loss_type == 'both_all':                                                 # :219-229 (no zero_grad)
                loss_l, loss_r = loss_l1 + loss_l2, loss_r1 + loss_r2
                loss = obj1 + obj2
                loss.backward()
                optimizer.step()
            elif loss_type == 'both_layerwise':                                         # :231-246
                loss_l, loss_r = loss_l1 + loss_l2, loss_r1 + loss_r2
                obj1.backward(retain_graph=True)
                optimizer[0].step()
                optimizer[0].zero_grad()
                obj2.backward(retain_graph=True)
                optimizer[1].step()
                optimizer[1].zero_grad()
                loss = obj1 + obj2
            elif loss_type == 'only2_layerwise':                                        # :264-279
                loss_l, loss_r = loss_l1 + loss_l2, loss_r1 + loss_r2
                optimizer[0].zero_grad()
                obj2.backward()
                optimizer[1].step()
                optimizer[1].zero_grad()
                loss = obj2
            elif loss_type == 'only2_all':                                              # :281-289
                loss_l, loss_r, loss = loss_l2, loss_r2, obj2
                loss.backward()
                optimizer.step()
                optimizer.zero_grad()
            else:                                                                       # 'only1', :291-299
                loss_l, loss_r, loss = loss_l1, loss_r1, obj1
                loss.backward()
                optimizer.step()
                optimizer.zero_grad()
            ring.append(torch.stack([loss.detach(), loss_r, loss_l]))
            del z1, z2, obj1, obj2, loss
            last = epoch + 1 == args.epochs
            if (epoch + 1) % self.log_every == 0 or last or (epoch + 1) % args.valid_freq == 0:
                vals = torch.stack(ring).cpu()
                dt = (time.time() - t0) / len(ring)
                for i, v in enumerate(vals.tolist()):
                    self.trainer_log['log'].append({'Epoch': epoch + 1 - len(ring) + i, 'train_loss': v[0],
                                                    'loss_r': v[1], 'loss_l': v[2], 'train_time': dt})
                ring, t0 = [], time.time()
            if (epoch + 1) % args.valid_freq == 0:                                      # :310-338
                valid_loss, dt_auc, dt_aup, df_auc, df_aup, _, _, valid_log = self.eval(model, data, 'val')
                valid_log['epoch'] = epoch
                self.trainer_log['log'].append(valid_log)
                if dt_auc + df_auc > best_metric:
                    best_metric = dt_auc + df_auc
                    torch.save({'model_state': model.state_dict()}, os.path.join(args.checkpoint_dir, 'model_best.pt'))
        torch.save({'model_state': {k: v.to('cpu') for k, v in model.state_dict().items()}},     # :341-346
                   os.path.join(args.checkpoint_dir, 'model_final.pt'))
        return model
