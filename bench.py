#!/usr/bin/env python
"""Headline benchmark: GNNDelete edge-unlearning (Del-training) epochs/s on the
OGB-Collab-shaped synthetic graph (BASELINE.json configs[2]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload NAME]

One step = one epoch: GCNDelete forward on the fixed sdf-masked edge
set (BOTH conv layers recomputed — hoisting the frozen layer-1 conv is reported
separately as ``value_hoisted``) -> decode on (Df, supplied negatives) ->
0.5*MSE(pos,neg) + 0.5*NI(edge form) -> backward to deletion{1,2}.deletion_weight -> Adam.

N > 1: the Collab-shaped graph fits one GPU ("small graphs stay on one GPU"), so for the headline ``value`` the ranks
run independent replicas (different seeds = different deletion requests), no data-path collective ("scaling": "weak").
The line ADDITIONALLY carries a ``partitioned`` block: BASELINE config 5 (10 M nodes / 200 M edges) row-partitioned
over the same N ranks (gnndelete_b200/dist.py, NCCL halo exchange), parity-checked against the single-GPU engine on a
1/16-scale graph before it is timed, with the one-GPU point of the same engine measured in the same invocation on
rank 0, the efficiency, the halo bytes and the device time inside each collective.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = 'Del-training epochs/s on OGB-Collab shape'
UNIT = 'epochs/s'
ARITH = ('fp32 storage and accumulation; dense contractions on tcgen05 tensor cores as 3xTF32 (hi/lo operand split, '
         'fp32-level accuracy, 1e-5 vs the fp64 oracle); aggregation / loss kernels plain fp32')
NCU_TRAFFIC_JSON = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')


# --------------------------------------------------------------------------- utils
def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def measured_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch group from the last committed `ncu --set full` capture
    (written by tools/ncu_traffic.py next to the commit it was taken at); None when there is no capture."""
    try:
        with open(NCU_TRAFFIC_JSON) as f:
            t = json.load(f)
        return t['kernels'][key]['dram_bytes'], {'commit': t.get('commit'), 'report': t.get('report')}
    except (OSError, KeyError, ValueError):
        return None, None


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                 '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        out = self.proc.communicate()[0]
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def dist_setup(n_gpus):
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.cuda.set_device(local)
        from gnndelete_b200.dist import nccl_options
        dist.init_process_group('nccl', device_id=torch.device('cuda', local), pg_options=nccl_options())
    return rank, local, world


def barrier_sync(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(value, world, dev):
    if world == 1:
        return value
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ------------------------------------------------------------------------ workload
def build_case(shape, seed, dev, gen_device='cpu', with_eval_edges=True, same_on_all_ranks=False):
    """Synthetic inputs + masks (CUDA mask pipeline) + random-init model + z_ori.  ``same_on_all_ranks``: the inputs
    generated on rank 0 are broadcast - torch's device-side random permutations / samplers are not bitwise
    reproducible from one process to the next at the 10 M-node size, and the ranks of a partition must hold the SAME
    graph."""
    from gnndelete_b200 import synthetic as S
    from gnndelete_b200 import masks as MK
    from gnndelete_b200 import models as M
    raw = S.make_graph(shape, seed=seed, device=gen_device, with_eval_edges=with_eval_edges).to(dev)
    df = S.sample_df_mask(shape.num_edges, shape.num_deleted, seed=seed, device=gen_device).to(dev)
    if same_on_all_ranks:
        import torch.distributed as dist
        dfb = df.to(torch.uint8)
        for t in (raw.train_pos_edge_index, raw.x, dfb):
            dist.broadcast(t, 0)
        df = dfb.bool()
    data = MK.build_unlearning_data(raw, df)
    neg = S.supplied_negatives(shape.num_nodes, int(data.df_mask.sum()), seed=seed + 1, device=gen_device).to(dev)
    if same_on_all_ranks:
        dist.broadcast(neg, 0)
    args = types.SimpleNamespace(in_dim=shape.in_dim, hidden_dim=shape.hidden_dim, out_dim=shape.out_dim)
    torch.manual_seed(seed)
    model = M.GCNDelete(args, data.sdf_node_1hop_mask, data.sdf_node_2hop_mask).to(dev)
    with torch.no_grad():
        z_ori = model.get_original_embeddings(data.x, data.train_pos_edge_index[:, data.dr_mask])
    return data, neg, model, z_ori


# SURVEY.md §8(d) algorithmic (compulsory) bytes: every operand once, int32 ids, fp32 elements
def spmm_algo_bytes(n, nnz, feat, src_elt=4):
    """rowptr + col + deg^-1/2 + read N*F + write N*F."""
    return 4 * (n + 1) + 4 * nnz + 4 * n + src_elt * n * feat + 4 * n * feat


def gemm_algo_bytes(m, k, n, gathered):
    return (4 * m if gathered else 0) + 4 * m * (k + n) + 4 * k * n


def loss_algo_bytes(n_df, n_ni, n, out):
    """decode + DEC + NI(edge form) fwd+bwd with precomputed target logits: pair ids, targets, z read once, dz written."""
    p = 2 * n_df + n_ni
    return 8 * p + 4 * n_ni + 4 * n * out + 4 * n * out


def khop_algo_bytes(e, n):
    """2 calls (2 hops + 1 hop) over E directed edges incl. the reference's bool-byte outputs."""
    return (2 + 1) * 8 * e + 2 * 8 * e + 2 * e // 8 + 5 * n // 8 + 2 * e + 2 * n


def time_kernels(eng, data, steps):
    """Eager epochs with CUDA events around each launch group (on the launch stream); ms averages."""
    from gnndelete_b200 import masks as MK
    from gnndelete_b200 import ops
    st = torch.cuda.current_stream()
    rec = {}

    def timed(name, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st); fn(); b.record(st)
        rec.setdefault(name, []).append((a, b))

    m, p = eng.model, eng.plan
    W1, W2 = m.conv1.lin.weight.detach(), m.conv2.lin.weight.detach()
    for _ in range(steps):
        timed('gemm_xw1', lambda: ops.gemm_rows(eng.x, W1, True, out=eng.h0, out_scale=p.dinv))
        timed('spmm_l1_f128', lambda: ops.spmm(p.fwd, eng.h0, out=eng.a1, row_scale=p.dinv, bias=m.conv1.bias.detach()))
        w1 = m.deletion1.deletion_weight.detach(); w2 = m.deletion2.deletion_weight.detach()
        timed('gemm_del1', lambda: ops.gemm_rows(eng.a1, w1, False, out=eng.x1, rows=eng.rows1, relu_mask_out=eng.x1_bits))
        ops.copy_rows(eng.a1, eng.x1, eng.comp1)
        timed('gemm_xw2', lambda: ops.gemm_rows(eng.x1, W2, True, out=eng.h1, out_scale=p.dinv, relu_in=True))
        timed('spmm_l2_f64', lambda: ops.spmm(p.fwd, eng.h1, out=eng.a2, row_scale=p.dinv, bias=m.conv2.bias.detach()))
        timed('gemm_del2', lambda: ops.gemm_rows(eng.a2, w2, False, out=eng.z, rows=eng.rows2))
        ops.copy_rows(eng.a2, eng.z, eng.comp2)
        timed('loss_fwd_bwd', lambda: (eng.loss.forward(eng.z, dz_out=eng.dz), eng.loss.backward(eng.z, out=eng.dz)))
        timed('gemm_dw2', lambda: ops.gemm_tn_rows(eng.a2, eng.dz, rows=eng.rows2, out=eng.params[1].grad))
        timed('gemm_da2', lambda: ops.gemm_rows(eng.dz, w2, True, out=eng.da2, rows=eng.rows2, out_scale=p.dinv))
        ops.copy_rows(eng.dz, eng.da2, eng.comp2, row_scale=p.dinv)
        timed('spmm_bwd_f64', lambda: ops.spmm(p.bwd, eng.da2, out=eng.dh1))
        if eng.fused_dxdw:      # dX1 chained into dW_del1 through tensor memory: one kernel (+ the reduction of the partials)
            timed('gemm_dxdw1', lambda: ops.gemm_dxdw(eng.dh1, W2, False, eng.a1, rows=eng.rows1, in_scale=p.dinv,
                                                      gate_bits=eng.x1_bits, out=eng.params[0].grad))
        else:
            timed('gemm_dx1', lambda: ops.gemm_rows(eng.dh1, W2, False, out=eng.dx1, rows=eng.rows1, out_scale=p.dinv,
                                                    gate=None if eng.bitmask else eng.x1, gate_bits=eng.x1_bits))
            timed('gemm_dw1', lambda: ops.gemm_tn_rows(eng.a1, eng.dx1, rows=eng.rows1, out=eng.params[0].grad))
    # deletion-mask construction (setup path, A9): the 2-hop + 1-hop k_hop masks of delete_gnn.py:128-151 on the DIRECTED
    # edge list (rebuilt here from the symmetrised one); each call ends with a host read of its status word, so the
    # bracket includes one launch gap
    ei = data.train_pos_edge_index
    keep = ei[0] < ei[1]
    dir_ei, seed = ei[:, keep].contiguous(), data.df_mask[keep].contiguous()
    for _ in range(min(steps, 5)):
        timed('khop_masks_2hop_1hop', lambda: (MK.khop_masks(dir_ei, seed, 2, data.num_nodes), MK.khop_masks(dir_ei, seed, 1, data.num_nodes)))
    torch.cuda.synchronize()
    return {k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in rec.items()}


def load_tensor_peak():
    """(TF32 dense TFLOP/s, source): half of the measured bf16 figure of MEASURED_PEAKS.json (kind::tf32 runs at half the
    bf16 rate; the sustained number - the kernel is timed inside a loop), else half of the recipe's fallback."""
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['bf16_tflops_sustained']) / 2, 'measured (MEASURED_PEAKS.json bf16_tflops_sustained / 2: kind::tf32 runs at half the bf16 rate)'
    except (OSError, KeyError, ValueError):
        return 1400.0 / 2, 'fallback (B200_PROFILING.md sustained bf16 1.4 PFLOP/s / 2)'


def run_dense_ni_block(dev, lib, steps):
    """BASELINE config 1 as the reference runs it: train_fullbatch's dense S2 x S2 NI loss (gnndelete.py:163-193,
    239-241) on the Cora shape - the path's one tensor-bound contraction.  Epochs/s of the captured engine with that
    loss, the kernel alone (CUDA events over graph replays) against the TF32 tensor roofline, and the fp32 CUDA-core
    kernel for comparison."""
    from gnndelete_b200 import synthetic as S
    from gnndelete_b200.engine import GCNDeleteEngine
    from gnndelete_b200 import graph as G
    shape = S.SHAPES['cora']
    data, neg, model, z_ori = build_case(shape, 42, dev)
    logits_ori = z_ori @ z_ori.t()                                 # what base.py:288 stores in pred_proba.pt
    st = torch.cuda.current_stream()
    out = {'workload': f'GCNDelete edge unlearning, cora-shaped synthetic graph ({shape.num_nodes} nodes / {shape.num_edges} directed '
                       f'edges / {shape.num_deleted} deleted), dense-block NI loss of train_fullbatch, full graph per step'}
    for mode in ('tc', 'simt'):
        os.environ['GD_DENSE_NI'] = mode
        eng = GCNDeleteEngine(model, data, neg, z_ori=z_ori, hoist_layer1=False, static_negatives=True, logits_ori=logits_ori)
        plan = eng.dense
        n_s, pairs = plan.n_s, plan.num_pairs
        eng.capture(warmup=2)
        for _ in range(3):
            eng.epoch()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k = max(10, min(steps, 50) if mode == 'tc' else 10)
        torch.cuda.synchronize()
        a.record(st)
        for _ in range(k):
            eng.epoch()
        b.record(st)
        torch.cuda.synchronize()
        ms_epoch = a.elapsed_time(b) / k
        # the kernel alone: gather of z[S], the contraction kernel, the scatter-add of dz[S] (graph of 5 calls)
        z, dz = eng.z, torch.zeros_like(eng.z)
        g = torch.cuda.CUDAGraph()
        plan.forward_backward(z, dz)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            for _ in range(5):
                plan.forward_backward(z, dz)
        g.replay(); torch.cuda.synchronize()
        a.record(st); g.replay(); b.record(st); torch.cuda.synchronize()
        ms_k = a.elapsed_time(b) / 5
        flops = 2 * 2.0 * n_s * n_s * 64                               # both contractions over the full S x S block (each pair from both sides)
        rec = {'epochs_per_s': 1e3 / ms_epoch, 'ms_per_epoch': ms_epoch, 'kernel_ms': ms_k,
               'fp32_equivalent_tflops': flops / (ms_k * 1e-3) / 1e12}
        if mode == 'tc':
            peak, src = load_tensor_peak()
            issued = 3 * flops                                         # 3xTF32: three tensor-core products per fp32-accurate one
            out.update({'n_s': n_s, 'pairs': pairs, 'epochs_per_s': rec['epochs_per_s'], 'ms_per_epoch': ms_epoch,
                        'roofline': {'kernel': 'gd::tc::dense_ni_tc_kernel (tcgen05.mma kind::tf32, 3xTF32, accumulators in TMEM)',
                                     'bound': 'tensor', 'achieved': issued / (ms_k * 1e-3) / 1e12, 'peak': peak, 'unit': 'TFLOP/s',
                                     'frac': issued / (ms_k * 1e-3) / 1e12 / peak, 'peak_source': src,
                                     'flops_issued_per_launch': issued, 'fp32_equivalent_tflops': rec['fp32_equivalent_tflops'],
                                     'kernel_ms': ms_k, 'target_bytes_per_launch': n_s * n_s * 4,
                                     'traffic': (measured_traffic('dense_ni_tc') or (None, None))[0]},
                        'dtype': 'tf32x3 (fp32-level accuracy, 1e-5 vs the fp64 oracle)'})
        else:
            out['cuda_core_fp32_kernel'] = rec
        del eng, g
        G._GLOBAL_CACHE = G.PlanCache()
        torch.cuda.empty_cache()
    os.environ.pop('GD_DENSE_NI', None)
    return out


def timed_epochs(eng, steps, world):
    """K epochs between barrier+sync brackets, CUDA events on the launch stream; ms total."""
    st = torch.cuda.current_stream()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier_sync(world)
    a.record(st)
    for _ in range(steps):
        eng.epoch()
    b.record(st)
    barrier_sync(world)
    return a.elapsed_time(b)


# ---------------------------------------------------------------- CPU reference leg
def cpu_epoch_runner(shape, seed=42):
    """The reference's CPU path for the same epoch: the oracle restatement executed with
    PyG's op sequence (eager PyTorch fp32 autograd) on all host cores."""
    from gnndelete_b200 import synthetic as S
    from oracle import models as OM
    from oracle import unlearn as OU
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    raw = S.make_graph(shape, seed=seed, device='cpu')
    df = S.sample_df_mask(shape.num_edges, shape.num_deleted, seed=seed)
    data = OU.build_unlearning_data(raw, df, num_edge_type=shape.num_edge_type or None)
    neg = S.supplied_negatives(shape.num_nodes, int(data.df_mask.sum()), seed=seed + 1)
    args = types.SimpleNamespace(in_dim=shape.in_dim, hidden_dim=shape.hidden_dim, out_dim=shape.out_dim)
    torch.manual_seed(seed)
    if shape.gnn == 'rgcn':
        return _cpu_kg_step_runner(shape, data, args), cores
    model = OM.DELETE_MODELS[shape.gnn](args, data.sdf_node_1hop_mask, data.sdf_node_2hop_mask)
    opt = torch.optim.Adam([p for n, p in model.named_parameters() if 'del' in n], lr=1e-3)
    with torch.no_grad():
        z_ori = model.get_original_embeddings(data.x, data.train_pos_edge_index[:, data.dr_mask])

    def epoch():
        loss, _, _, _ = OU.edge_form_loss(model, data, neg, z_ori, masks_positional=False)
        loss.backward()
        opt.step()
        opt.zero_grad()
        return float(loss.detach())

    return epoch, cores


def _cpu_kg_step_runner(shape, data, args):
    """BASELINE config 4: the KG node-embedding step (gnndelete_nodeemb.py:744-798) with the per-relation Python loop of
    PyG's RGCNConv - forward, original embeddings, two backward passes, two Adam steps."""
    from oracle import models as OM
    from oracle import unlearn as OU
    net = shape.num_edge_type
    model = OM.RGCNDelete(args, shape.num_nodes, net, data.sdf_node_1hop_mask, data.sdf_node_2hop_mask)
    pos_ei, pos_et = data.edge_index[:, data.df_mask], data.edge_type[data.df_mask]
    dec = pos_et < net
    neg = OU.negative_sampling_kg(pos_ei[:, dec], pos_et[dec], generator=torch.Generator().manual_seed(43))
    o1 = torch.optim.Adam(model.deletion1.parameters(), lr=1e-3)
    o2 = torch.optim.Adam(model.deletion2.parameters(), lr=1e-3)

    def step():
        loss1, loss2, _ = OU.kg_step_losses(model, data, neg, net, alpha=0.5)
        loss1.backward(retain_graph=True); o1.step(); o1.zero_grad()
        loss2.backward(retain_graph=True); o2.step(); o2.zero_grad()
        return float((loss1 + loss2).detach())

    return step


def run_reference(args, shape):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    epoch, cores = cpu_epoch_runner(shape)
    for _ in range(max(1, min(args.warmup, 2))):
        epoch()
    # every step is one FULL epoch of the workload (about a second on 16 cores); the run is bounded in time instead:
    # a driver-supplied --steps sized for the GPU arm stops after --cpu-budget-s seconds and reports the steps it timed
    done, t0 = 0, time.perf_counter()
    while done < args.steps:
        epoch()
        done += 1
        if time.perf_counter() - t0 > args.cpu_budget_s:
            break
    dt = time.perf_counter() - t0
    requested, args.steps = args.steps, done
    value = args.steps / dt
    sample = (f'{args.steps} full epochs of the {shape.name} workload (whole graph, no sub-sampling)' +
              (f'; {requested} requested, stopped at the {args.cpu_budget_s:.0f} s CPU budget' if done < requested else ''))
    line = {
        'impl': 'reference', 'metric': metric_name(shape),
        'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(shape), 'where': 'host CPU',
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': 'PyTorch-Geometric is not installable here; the reference arm is the oracle restatement run with '
                "PyG's op sequence (eager PyTorch CPU autograd) on ONE host process with all host threads, parity "
                'unpinned; at N > 1 it is still one process (the N-GPU native line is N replicas)',
    }
    print(json.dumps(line), flush=True)


def metric_name(shape):
    return METRIC if shape.name.startswith('collab') else f'Del-training epochs/s on {shape.name} shape'


def workload_config(shape, where=None):
    """Identical for the native and the reference arm (the driver pairs their lines by metric and config)."""
    return {
        'workload': f'{shape.gnn.upper()}Delete edge unlearning, {shape.name}-shaped synthetic power-law graph '
                    f'({shape.num_nodes} nodes / {shape.num_edges} directed train edges / {shape.num_deleted} deleted), '
                    f'{shape.in_dim}->{shape.hidden_dim}->{shape.out_dim}, ' +
                    ('node-embedding DEC/NI losses (KG step)' if shape.gnn == 'rgcn' else 'edge-form NI loss') +
                    ', full graph per step',
        'epoch': ('fwd + original embeddings + two backward passes + two Adam steps' if shape.gnn == 'rgcn' else
                  'fwd (both convs recomputed) + decode + DEC/NI loss + bwd to Del weights + Adam'),
        'l2': 'per-epoch working set ~0.9 GB > 126 MB L2, no flush between steps',
        'arith': ARITH,
    }


# ---------------------------------------------------------- row-partitioned config
def _partition_parity(rank, world, dev, wire, scale=1.0 / 16):
    """Before anything is timed: the partitioned engine over the N ranks against the single-GPU engine on a
    1/16-scale power-law graph - losses and both Del gradients after the first step, weights after 3 steps."""
    import dataclasses
    from gnndelete_b200 import models as M
    from gnndelete_b200 import synthetic as S
    from gnndelete_b200.dist import PartitionedGCNDeleteEngine
    from gnndelete_b200.engine import GCNDeleteEngine
    shape = S.SHAPES['powerlaw10m'].scaled(scale)
    data, neg, model, z_ori = build_case(shape, 42, dev, gen_device=dev, with_eval_edges=False, same_on_all_ranks=world > 1)
    init = {k: v.clone() for k, v in model.state_dict().items()}
    margs = types.SimpleNamespace(in_dim=shape.in_dim, hidden_dim=shape.hidden_dim, out_dim=shape.out_dim)
    m2 = M.GCNDelete(margs, data.sdf_node_1hop_mask, data.sdf_node_2hop_mask).to(dev)
    m2.load_state_dict(init)
    part = PartitionedGCNDeleteEngine(model, data, neg, z_ori, wire=wire, world=world, rank=rank)
    one = GCNDeleteEngine(m2, data, neg, z_ori=z_ori, hoist_layer1=False, static_negatives=True)
    part.forward(); part.backward()
    one.forward_backward()

    def rel(a, b):
        return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))

    err = {'losses': rel(part.losses, one.loss.losses),
           'dW_del1': rel(model.deletion1.deletion_weight.grad, m2.deletion1.deletion_weight.grad),
           'dW_del2': rel(model.deletion2.deletion_weight.grad, m2.deletion2.deletion_weight.grad)}
    part.adam_step(); one.adam_step()
    for _ in range(2):
        part.epoch(); one.epoch()
    err['W_del1_after_3_steps'] = rel(model.deletion1.deletion_weight, m2.deletion1.deletion_weight)
    err['W_del2_after_3_steps'] = rel(model.deletion2.deletion_weight, m2.deletion2.deletion_weight)
    tol = 2e-2 if wire == 'bf16' else 1e-5
    ok = all(v <= tol for v in err.values())
    out = {'graph': f'{shape.name}: {shape.num_nodes} nodes / {shape.num_edges} directed edges', 'tolerance': tol,
           'wire': wire, 'rel_err_vs_single_gpu_engine': err, 'ok': ok}
    del part, one, data, model, m2
    from gnndelete_b200 import graph as G
    G._GLOBAL_CACHE = G.PlanCache()
    torch.cuda.empty_cache()
    return out


def run_partitioned(args, shape, rank, local, world, dev, lib, wire='bf16', one_gpu_point=True):
    """BASELINE config 5: GCNDelete on the power-law graph, 1-D row partitioned over the ranks (strong scaling: the
    graph is fixed, ranks split its rows).  The graph is generated on the device with the same seed on every rank.
    Returns the ``partitioned`` block (identical on every rank)."""
    from gnndelete_b200 import graph as G
    from gnndelete_b200.dist import PartitionedGCNDeleteEngine
    steps = max(3, min(args.steps, args.partition_steps))
    warm = max(3, min(args.warmup, 5))
    parity = _partition_parity(rank, world, dev, wire) if (world > 1 and shape.num_nodes >= 1_000_000) else None
    if parity is not None and not parity['ok']:
        raise RuntimeError(f'partitioned epoch disagrees with the single-GPU engine: {parity}')
    t_setup = time.perf_counter()
    data, neg, model, z_ori = build_case(shape, 42, dev, gen_device=dev, with_eval_edges=False, same_on_all_ranks=world > 1)
    G._GLOBAL_CACHE = G.PlanCache()                      # drop the dr-edge plan before the engines allocate
    torch.cuda.empty_cache()
    init = {k: v.clone() for k, v in model.state_dict().items()}
    n = shape.num_nodes
    nnz = int(data.sdf_mask.sum()) + n

    def measure(w, r, group_world):
        """epochs/s of the partitioned engine over ``w`` ranks (w == 1: no process group, this rank alone)."""
        model.load_state_dict(init)
        eng = PartitionedGCNDeleteEngine(model, data, neg, z_ori, wire=wire, world=w, rank=r, exchange=args.partition_exchange,
                                         overlap_layer1=None if args.partition_overlap == 'auto' else False)
        overlap, exchange = eng.overlap, eng.exchange + (f' (symm failed: {eng.exchange_error})' if hasattr(eng, 'exchange_error') else '')
        torch.cuda.synchronize()
        eng.epoch()                                       # first epoch: also builds the batch plans (one-time launches)
        c0 = lib.gd_launch_count()
        eng.epoch()
        launches = lib.gd_launch_count() - c0
        for _ in range(warm):
            eng.epoch()
        eng.comm_events = []
        ms = timed_epochs(eng, steps, group_world)
        comm = eng.comm_ms()
        eng.comm_events = None
        ms = max_over_ranks(ms, group_world, dev)
        out = {'ms_per_step': ms / steps, 'value': steps / (ms / 1e3), 'launches_per_epoch': int(launches),
               'comm_ms_per_epoch': {k: v / steps for k, v in comm.items()}, 'losses_last': eng.losses.tolist(),
               'halo_bytes_per_epoch_per_rank_received': int(eng.halo_bytes_per_epoch * (w - 1) / w) if w > 1 else 0,
               'rows_per_rank': [b[1] - b[0] for b in eng.plan.bounds], 'overlap_layer1': overlap, 'exchange': exchange}
        del eng
        torch.cuda.empty_cache()
        return out

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    res = measure(world, rank, world)
    clocks = sampler.stop() if rank == 0 else None
    setup_s = time.perf_counter() - t_setup
    block = {
        'workload': f'GCNDelete edge unlearning, {shape.name} power-law graph ({n} nodes / {shape.num_edges} directed edges / '
                    f'{shape.num_deleted} deleted), 128->128->64, 1-D row partition (work-balanced row blocks) over '
                    f'{world} GPU(s), halo exchange of H1 / z / dA2 in {wire} over NVLink (copy-engine pulls from symmetric peer '
                    f'buffers, or NCCL all-gather: see "exchange"; H0 exchanged once at setup: frozen, input-constant), layer-1 '
                    f'aggregation still run every epoch',
        'metric': 'Del-training epochs/s, row-partitioned power-law graph', 'unit': UNIT, 'scaling': 'strong',
        'n_gpus': world, 'comm_nranks_seen': world, 'steps': steps, 'warmup': warm, 'wire': wire,
        'dtype': 'bf16-gather (fp32 accumulate), tolerance 2e-2' if wire == 'bf16' else 'f32',
        'value': res['value'], 'ms_per_step': res['ms_per_step'], 'nnz_message_passing': nnz,
        'launches_per_epoch': res['launches_per_epoch'], 'comm_ms_per_epoch': res['comm_ms_per_epoch'],
        'halo_bytes_per_epoch_per_rank_received': res['halo_bytes_per_epoch_per_rank_received'],
        'rows_per_rank': res['rows_per_rank'], 'losses_last': res['losses_last'], 'parity': parity, 'clocks': clocks,
        'overlap_layer1_with_exchanges': res['overlap_layer1'], 'exchange': res['exchange'],
        'setup_s': setup_s,
    }
    if res['comm_ms_per_epoch']:
        lim = max(res['comm_ms_per_epoch'].items(), key=lambda kv: kv[1])
        block['limiting_collective'] = {'name': lim[0], 'ms_per_epoch': lim[1],
                                        'note': 'device time on rank 0 between the events bracketing the transfer (on the copy stream for pulls); '
                                                'pulls run underneath the next epoch\'s layer-1 aggregation, so this is not all exposed time'}
    if world > 1 and one_gpu_point:
        # the one-GPU point of the same engine (same kernels, same arithmetic), on rank 0's GPU, same invocation
        if rank == 0:
            one = measure(1, 0, 1)
            t = torch.tensor([one['value']], dtype=torch.float64, device=dev)
        else:
            t = torch.zeros(1, dtype=torch.float64, device=dev)
        import torch.distributed as dist
        dist.broadcast(t, 0)
        block['one_gpu_value'] = float(t.item())
        block['efficiency'] = block['value'] / (world * block['one_gpu_value'])
        block['speedup_vs_one_gpu'] = block['value'] / block['one_gpu_value']
    return block


# ------------------------------------------------------------------ KG config (4)
def rgcn_algo_bytes(n, nnz, num_rel, fin, fout):
    """SURVEY.md §8(d) RGCN row: rowptr + col + relation id (uint8) + read N*in + write N*out + block weights."""
    return 4 * (n + 1) + 4 * nnz + nnz + 4 * n * fin + 4 * n * fout + 4 * num_rel * fin * fout // 4


def run_kg(args, shape, rank, local, dev, lib):
    """BASELINE config 4: RGCNDelete on the BioKG shape, the KG node-embedding step (gnndelete_nodeemb.py:744-798)
    through framework.get_trainer(args).start(...): forward, two backward passes, two Adam steps per step."""
    import tempfile
    import framework
    from gnndelete_b200 import masks as MK
    from gnndelete_b200 import ops
    from gnndelete_b200 import synthetic as S
    from gnndelete_b200.graph import plan_for
    from gnndelete_b200.kg import negative_sampling_kg
    net = shape.num_edge_type
    raw = S.make_graph(shape, seed=42, device='cpu').to(dev)
    df = S.sample_df_mask(shape.num_edges, shape.num_deleted, seed=42, device='cpu').to(dev)
    data = MK.build_unlearning_data(raw, df, num_edge_type=net)
    pos_ei, pos_et = data.edge_index[:, data.df_mask], data.edge_type[data.df_mask]
    dec = pos_et < net
    neg = negative_sampling_kg(pos_ei[:, dec], pos_et[dec], torch.Generator(device=dev).manual_seed(43))   # supplied (§8(d))

    def session(capture, with_neg):
        targs = types.SimpleNamespace(unlearning_model='gnndelete', gnn='rgcn', dataset='ogbl-biokg-shaped', epochs=0,
                                      valid_freq=10 ** 9, checkpoint_dir=tempfile.mkdtemp(prefix='gd_bench_kg_'), in_dim=shape.in_dim,
                                      hidden_dim=shape.hidden_dim, out_dim=shape.out_dim, lr=1e-3, alpha=0.5, random_seed=42,
                                      num_edge_type=net, capture_step=capture, device=str(dev), loss_fct='mse_mean',
                                      loss_type='both_layerwise', eval_on_cpu=False)
        torch.manual_seed(42)
        model = framework.get_model(targs, data.sdf_node_1hop_mask, data.sdf_node_2hop_mask, num_nodes=data.num_nodes,
                                    num_edge_type=net).to(dev)
        opt = [torch.optim.Adam([model.deletion1.deletion_weight], lr=1e-3), torch.optim.Adam([model.deletion2.deletion_weight], lr=1e-3)]
        d = data.clone()
        if with_neg:
            d.neg_edge_index = neg
        trainer = framework.get_trainer(targs)
        return trainer, trainer.start(model, d, opt, targs), model

    trainer, sess, model = session(True, True)
    sess.step()
    c0 = lib.gd_launch_count()
    if sess.graph is None:
        sess.step()
    launches = lib.gd_launch_count() - c0
    for _ in range(args.warmup):
        sess.step()
    st = torch.cuda.current_stream()
    sampler = ClockSampler(local)
    sampler.start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record(st)
    for _ in range(args.steps):
        sess.step()
    b.record(st)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = a.elapsed_time(b)
    losses = sess.out.tolist()
    # the relation aggregation kernels on their own (layer-2 forward and its transpose): CUDA events, eager
    peak, peak_src = load_peaks()
    ei, et = sess.edge_index, sess.edge_type
    plan = plan_for(ei, shape.num_nodes, 'rgcn', et, 2 * net)
    n, nnz, R = shape.num_nodes, int(ei.shape[1]), 2 * net
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(n, shape.hidden_dim, generator=g, device=dev)
    go = torch.randn(n, shape.out_dim, generator=g, device=dev)
    c2 = model.conv2
    kt = {}
    for name, fn in (('rgcn_conv2_fwd_128_64', lambda: ops.rgcn_conv(plan, x, c2.weight, c2.root, c2.bias)),
                     ('rgcn_conv2_transposed_64_128', lambda: ops.rgcn_conv(plan, go, c2.weight, c2.root, None, transposed=True))):
        for _ in range(3):
            fn()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for e0, e1 in ev:
            e0.record(st); fn(); e1.record(st)
        torch.cuda.synchronize()
        kt[name] = sum(e0.elapsed_time(e1) for e0, e1 in ev) / len(ev)
    bts = {'rgcn_conv2_fwd_128_64': rgcn_algo_bytes(n, nnz, R, shape.hidden_dim, shape.out_dim),
           'rgcn_conv2_transposed_64_128': rgcn_algo_bytes(n, nnz, R, shape.out_dim, shape.hidden_dim)}
    flops = 2.0 * nnz * shape.hidden_dim * shape.out_dim / 4 + 2.0 * n * shape.hidden_dim * shape.out_dim
    kern = {k: {'ms': kt[k], 'algo_bytes': bts[k], 'gbs': bts[k] / (kt[k] * 1e-3) / 1e9, 'frac': bts[k] / (kt[k] * 1e-3) / 1e9 / peak,
                'fp32_tflops': flops / (kt[k] * 1e-3) / 1e12} for k in kt}
    dom = max(kt, key=kt.get)
    roofline = {'kernel': f'gd::rgcn_edge_kernel + reduce + root GEMM ({dom})', 'bound': 'hbm', 'achieved': kern[dom]['gbs'], 'peak': peak,
                'unit': 'GB/s', 'frac': kern[dom]['frac'], 'traffic': None, 'peak_source': peak_src,
                'algorithmic_bytes_per_launch': bts[dom], 'kernel_ms': kt[dom], 'other_kernels': {k: v for k, v in kern.items() if k != dom},
                'note': 'per-edge block products (2 nnz in out / 4 flops) on the fp32 FMA pipes: this kernel is FMA / latency bound, '
                        'far from the HBM roofline of its compulsory bytes (fp32_tflops column)'}
    # e2e: a session that takes this step's corrupted triples from pinned host memory (eager: the loss incidence is rebuilt)
    _, sess_e, _ = session(False, False)
    neg_host = neg.cpu().pin_memory()
    out_host = torch.empty(3, dtype=torch.float32).pin_memory()

    def e2e_step():
        out_host.copy_(sess_e.step(neg_host), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(3):
        e2e_step()
    k = min(args.steps, 30)
    t0 = time.perf_counter()
    for _ in range(k):
        e2e_step()
    e2e_s = time.perf_counter() - t0
    line = {'metric': metric_name(shape), 'value': args.steps / (ms / 1e3), 'unit': UNIT, 'n_gpus': 1, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(shape), 'where': 'B200', 'roofline': roofline,
            'e2e': {'value': k / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': neg_host.numel() * 8, 'd2h_bytes_per_step': 12, 'steps': k,
                    'what': 'KGNodeembSession.step(negatives=pinned host triples): H2D, rebuild of the two loss incidences, the step '
                            '(eager launches), losses -> pinned host'},
            'gpu_launches': int(launches * args.steps), 'launches_per_epoch': int(launches), 'captured_step': sess.graph is not None,
            'clocks': clocks, 'losses_last': losses, 'parallelism': 'single'}
    if not args.no_cpu_baseline:
        epoch, cores = cpu_epoch_runner(shape)
        epoch()
        t0 = time.perf_counter()
        epoch()
        dt = time.perf_counter() - t0
        line['cpu_baseline'] = {'value': 1.0 / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                                'sample': '1 full KG step of the same workload (after 1 warm-up)'}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=None)
    ap.add_argument('--warmup', type=int, default=None)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--workload', default='collab')
    ap.add_argument('--scale', type=float, default=1.0)
    ap.add_argument('--cpu-epochs', type=int, default=4, help='bounded CPU-baseline sample (epochs)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-budget-s', type=float, default=150.0, help='time cap of the --impl reference loop')
    ap.add_argument('--wire', default='bf16', choices=['bf16', 'fp32'], help='halo format of the row-partitioned epoch')
    ap.add_argument('--partition-steps', type=int, default=10, help='timed epochs of the row-partitioned block')
    ap.add_argument('--partition-workload', default='powerlaw10m')
    ap.add_argument('--partition-scale', type=float, default=1.0)
    ap.add_argument('--no-partitioned', action='store_true', help='N > 1: skip the row-partitioned config-5 block')
    ap.add_argument('--partition-exchange', default='symm', choices=['symm', 'nccl'], help='halo exchange of the row-partitioned epoch')
    ap.add_argument('--partition-overlap', default='auto', choices=['auto', 'off'],
                    help="'off': do not run the next epoch's layer-1 aggregation under the collectives (A/B measurements)")
    args = ap.parse_args()
    from gnndelete_b200 import synthetic as S
    shape = S.SHAPES[args.workload].scaled(args.scale)

    if args.impl == 'reference':
        args.steps = args.steps if args.steps is not None else 5
        args.warmup = args.warmup if args.warmup is not None else 1
        return run_reference(args, shape)

    args.steps = args.steps if args.steps is not None else 200
    args.warmup = max(3, args.warmup if args.warmup is not None else 10)
    rank, local, world = dist_setup(args.gpus)
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    from gnndelete_b200 import _lib, build
    if rank == 0:
        build.build()
    barrier_sync(world)
    lib = _lib.load()
    from gnndelete_b200.engine import GCNDeleteEngine
    if args.workload.startswith('powerlaw'):
        # the partitioned config on its own (any N, incl. 1): prints the block as the line
        block = run_partitioned(args, shape, rank, local, world, dev, lib, wire=args.wire, one_gpu_point=False)
        if rank == 0:
            line = {'metric': block['metric'], 'value': block['value'], 'unit': UNIT, 'n_gpus': world, 'steps': block['steps'],
                    'warmup': block['warmup'], 'ms_per_step': block['ms_per_step'], 'higher_is_better': True,
                    'scaling': 'strong', 'vs_baseline': None, 'dtype': block['dtype'], 'data': 'synthetic',
                    'config': {'workload': block['workload'], 'arith': ARITH, 'l2': 'feature matrices (>= 0.25 GB each) exceed the 126 MB L2, no flush'},
                    'gpu_launches': block['launches_per_epoch'] * block['steps'], 'partitioned': block, 'clocks': block['clocks']}
            print(json.dumps(line), flush=True)
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return

    if shape.gnn == 'rgcn':
        if world > 1:
            raise SystemExit('the KG config runs on one GPU')
        return run_kg(args, shape, rank, local, dev, lib)

    data, neg, model, z_ori = build_case(shape, 42 + rank, dev)
    n = shape.num_nodes
    nnz = data.train_pos_edge_index[:, data.sdf_mask].shape[1] + n

    # ---- value: device-resident epochs, CUDA-graph replay, both conv layers recomputed
    # (the supplied negatives are fixed for the run, SURVEY.md §8(d): one loss-gradient gather over one incidence)
    eng = GCNDeleteEngine(model, data, neg, z_ori=z_ori, hoist_layer1=False, static_negatives=True)
    eng.epoch()                                       # first epoch: also builds the batch plans (one-time launches)
    c0 = lib.gd_launch_count()
    eng.epoch()
    launches_per_epoch = lib.gd_launch_count() - c0
    eng.capture(warmup=2)
    for _ in range(args.warmup):
        eng.epoch()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed_epochs(eng, args.steps, world)
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(ms, world, dev)
    value = world * args.steps / (ms / 1e3)
    losses = eng.loss.losses.tolist()

    # ---- value_hoisted: layer-1 conv (frozen, input-constant) computed once outside the loop
    eng_h = GCNDeleteEngine(model, data, neg, z_ori=z_ori, hoist_layer1=True, static_negatives=True)
    eng_h.capture(warmup=2)
    for _ in range(args.warmup):
        eng_h.epoch()
    ms_h = max_over_ranks(timed_epochs(eng_h, args.steps, world), world, dev)
    del eng_h

    # ---- per-kernel durations (eager launches, events on the launch stream) -> roofline
    peak, peak_src = load_peaks()
    eng_k = GCNDeleteEngine(model, data, neg, z_ori=z_ori, hoist_layer1=False, static_negatives=True)
    time_kernels(eng_k, data, 3)
    kt = time_kernels(eng_k, data, min(args.steps, 50))
    hid, out = shape.hidden_dim, shape.out_dim
    n1, n2 = int(eng_k.rows1.numel()), int(eng_k.rows2.numel())
    n_df, n_ni = eng_k.loss.n_df, eng_k.loss.n_ni
    algo = {
        'spmm_l1_f128': spmm_algo_bytes(n, nnz, hid), 'spmm_l2_f64': spmm_algo_bytes(n, nnz, out),
        'spmm_bwd_f64': spmm_algo_bytes(n, nnz, out),
        'gemm_xw1': gemm_algo_bytes(n, shape.in_dim, hid, False), 'gemm_xw2': gemm_algo_bytes(n, hid, out, False),
        'gemm_del1': gemm_algo_bytes(n1, hid, hid, True), 'gemm_del2': gemm_algo_bytes(n2, out, out, True),
        'gemm_da2': gemm_algo_bytes(n2, out, out, True), 'gemm_dw2': gemm_algo_bytes(n2, out, out, True),
        'loss_fwd_bwd': loss_algo_bytes(n_df, n_ni, n, out),
        'khop_masks_2hop_1hop': khop_algo_bytes(shape.num_edges, n),
    }
    if eng_k.fused_dxdw:    # rows of dH1 [out] and a1 [hid] read once, row id + scale + gate bits per row; dX1 never touches HBM
        algo['gemm_dxdw1'] = n1 * (4 * out + 4 * hid + 8 + hid // 8) + 4 * out * hid + 4 * hid * hid
    else:
        algo['gemm_dx1'] = gemm_algo_bytes(n1, out, hid, True)
        algo['gemm_dw1'] = gemm_algo_bytes(n1, hid, hid, True)
    kernels = {k: {'ms': kt[k], 'algo_bytes': algo[k], 'gbs': algo[k] / (kt[k] * 1e-3) / 1e9,
                   'frac': algo[k] / (kt[k] * 1e-3) / 1e9 / peak} for k in algo}
    agg = ['spmm_l1_f128', 'spmm_l2_f64', 'spmm_bwd_f64']
    agg_bytes, agg_ms = sum(algo[k] for k in agg), sum(kt[k] for k in agg)
    dom = max(agg, key=lambda k: kt[k])               # the time-dominant aggregation of the epoch
    traffic, traffic_src = measured_traffic(dom) if shape.name == 'collab' else (None, None)
    gather_bytes = algo[dom] - 4 * n * (hid if dom == 'spmm_l1_f128' else out) + 4 * nnz * (hid if dom == 'spmm_l1_f128' else out)
    roofline = {
        'kernel': {'spmm_l1_f128': 'gd::spmm_batched_kernel<16,false,SCALE|BIAS> x2 (GCN layer-1 aggregation, F=128 as two 64-wide column passes)',
                   'spmm_l2_f64': 'gd::spmm_batched_kernel<16,false,SCALE|BIAS> (GCN layer-2 aggregation, F=64)',
                   'spmm_bwd_f64': 'gd::spmm_batched_kernel<16,true,0> (transpose-backward aggregation, F=64)'}[dom],
        'bound': 'hbm', 'achieved': kernels[dom]['gbs'], 'peak': peak, 'unit': 'GB/s', 'frac': kernels[dom]['frac'],
        'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src,
        'algorithmic_bytes_per_launch': algo[dom], 'kernel_ms': kt[dom], 'bytes_gather_per_launch': gather_bytes,
        'gathered_tbs': gather_bytes / (kt[dom] * 1e-3) / 1e12, 'gather_ceiling_tbs': 18.0,
        'aggregations_per_epoch': {'algo_bytes': agg_bytes, 'ms': agg_ms, 'frac': agg_bytes / (agg_ms * 1e-3) / 1e9 / peak},
        'gemm_share_of_epoch': sum(kt[k] for k in kt if k.startswith('gemm_')) / (ms / args.steps),
        'other_kernels': {k: v for k, v in kernels.items() if k != dom},
    }
    del eng_k

    # ---- e2e: through the drop-in trainer (framework.get_trainer(args).start(...)): per step pinned-host negatives -> device,
    #      epoch (one CUDA-graph launch incl. the negative-pair gradient), losses -> pinned host
    import tempfile
    import framework
    targs = types.SimpleNamespace(unlearning_model='gnndelete', gnn='gcn', dataset='ogbl-collab-shaped', epochs=0, valid_freq=10 ** 9,
                                  checkpoint_dir=tempfile.mkdtemp(prefix='gd_bench_'), in_dim=shape.in_dim, hidden_dim=hid,
                                  out_dim=out, lr=1e-3, random_seed=42, hoist_layer1=False, device=str(dev))
    trainer = framework.get_trainer(targs)
    d_e2e = data.clone()
    d_e2e.z_ori = z_ori
    opt = torch.optim.Adam([model.deletion1.deletion_weight, model.deletion2.deletion_weight], lr=1e-3)
    sess = trainer.start(model, d_e2e, opt, targs)
    neg_host = neg.cpu().pin_memory()
    out_host = torch.empty(3, dtype=torch.float32).pin_memory()

    def e2e_step():
        out_host.copy_(sess.step(neg_host), non_blocking=True)     # H2D negatives, one graph launch, losses -> pinned host
        torch.cuda.current_stream().synchronize()

    for _ in range(3):
        e2e_step()
    e2e_steps = min(args.steps, 50)
    barrier_sync(world)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier_sync(world)
    e2e_sync_s = max_over_ranks(time.perf_counter() - t0, world, dev)
    # the same steps with the transfers overlapped: H2D of step k and D2H of step k-1 ride copy streams while the graph of
    # the neighbouring step runs; the host reads every step's losses, one step late
    for i in range(3):
        sess.result(sess.submit(neg_host))
    barrier_sync(world)
    t0 = time.perf_counter()
    last = None
    for _ in range(e2e_steps):
        k = sess.submit(neg_host)
        if last is not None:
            sess.result(last)
        last = k
    sess.result(last)
    barrier_sync(world)
    e2e_s = max_over_ranks(time.perf_counter() - t0, world, dev)
    e2e = {'value': world * e2e_steps / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': neg_host.numel() * 8,
           'd2h_bytes_per_step': 12, 'steps': e2e_steps, 'value_step_synchronous': world * e2e_steps / e2e_sync_s,
           'captured_step': bool(trainer.trainer_log.get('captured_step')),
           'what': 'through framework.get_trainer(args).start(model, data, optimizer, args) -> EdgeFormSession (the object '
                   'GNNDeleteTrainer.train loops over): per step new negatives pinned-host -> device, epoch with both convs '
                   'recomputed (one CUDA-graph launch; the negative pairs\' gradient is added with vector float reductions, '
                   'no per-step sort), losses -> pinned host and read by the host; transfers of neighbouring steps overlap '
                   'compute (value_step_synchronous: same steps with a full host sync after every step)'}

    line = {
        'metric': metric_name(shape), 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(shape), 'where': 'B200',
        'value_hoisted': world * args.steps / (ms_h / 1e3), 'ms_per_step_hoisted': ms_h / args.steps,
        'roofline': roofline, 'e2e': e2e, 'gpu_launches': int(launches_per_epoch * args.steps),
        'launches_per_epoch': int(launches_per_epoch), 'clocks': clocks,
        'losses_last': losses, 'parallelism': 'replicas' if world > 1 else 'single',
    }
    del sess, trainer, eng

    # ---- bf16-gather mode (opt-in, separate number: never the headline): gathered operands stored in bf16, fp32 accumulate
    if world == 1 and shape.name == 'collab':
        from gnndelete_b200.dist import PartitionedGCNDeleteEngine
        from gnndelete_b200 import graph as G
        G._GLOBAL_CACHE = G.PlanCache()
        torch.cuda.empty_cache()
        try:
            e16 = PartitionedGCNDeleteEngine(model, data, neg, z_ori, wire='bf16', world=1, rank=0, hoist_gather=False)
            for _ in range(max(3, args.warmup)):
                e16.epoch()
            ms16 = timed_epochs(e16, args.steps, 1)
            st = torch.cuda.current_stream()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
            from gnndelete_b200 import ops
            for a, b in ev:
                a.record(st); ops.spmm(e16.csr, e16.h1, out=e16.a2, row_scale=e16.dinv, bias=model.conv2.bias.detach()); b.record(st)
            torch.cuda.synchronize()
            t16 = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
            b16 = spmm_algo_bytes(n, nnz, out, src_elt=2)
            line['bf16_gather'] = {'dtype': 'bf16-gather', 'tolerance': 2e-2, 'value': args.steps / (ms16 / 1e3), 'unit': UNIT,
                                   'ms_per_step': ms16 / args.steps, 'captured': False,
                                   'spmm_l2_f64': {'ms': t16, 'algo_bytes': b16, 'gbs': b16 / (t16 * 1e-3) / 1e9,
                                                   'frac': b16 / (t16 * 1e-3) / 1e9 / peak},
                                   'what': 'same epoch with H0 / H1 / z / dA2 rounded to bf16 before they are gathered (fp32 '
                                           'accumulation and outputs); eager launches (not graph-captured), so compare its '
                                           'spmm line, not its epochs/s, with the fp32 path'}
            del e16
        except Exception as exc:                       # the opt-in mode must never take the headline down
            line['bf16_gather'] = {'error': repr(exc)}

    # ---- config 1 with the reference's dense-block NI loss: the tensor-core contraction of the path (own roofline block)
    if world == 1 and shape.name == 'collab':
        try:
            line['dense_ni'] = run_dense_ni_block(dev, lib, args.steps)
        except Exception as exc:
            line['dense_ni'] = {'error': repr(exc)}

    if world > 1 and not args.no_partitioned:
        del data, model, z_ori, neg
        from gnndelete_b200 import graph as G
        G._GLOBAL_CACHE = G.PlanCache()
        torch.cuda.empty_cache()
        pshape = S.SHAPES[args.partition_workload].scaled(args.partition_scale)
        try:
            line['partitioned'] = run_partitioned(args, pshape, rank, local, world, dev, lib, wire=args.wire)
        except Exception as exc:
            line['partitioned'] = {'error': repr(exc)}
            raise
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        epoch, cores = cpu_epoch_runner(shape)
        epoch()
        t0 = time.perf_counter()
        for _ in range(args.cpu_epochs):
            epoch()
        dt = time.perf_counter() - t0
        line['cpu_baseline'] = {'value': args.cpu_epochs / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                                'sample': f'{args.cpu_epochs} full epochs of the same workload (after 1 warm-up)'}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
