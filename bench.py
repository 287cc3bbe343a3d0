#!/usr/bin/env python
"""Headline benchmark: GNNDelete edge-unlearning (Del-training) epochs/s on the
OGB-Collab-shaped synthetic graph (BASELINE.json configs[2]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload NAME]

One step = one epoch: GCNDelete forward on the fixed sdf-masked edge
set (BOTH conv layers recomputed — hoisting the frozen layer-1 conv is reported
separately as ``value_hoisted``) -> decode on (Df, supplied negatives) ->
0.5*MSE(pos,neg) + 0.5*NI(edge form) -> backward to deletion{1,2}.deletion_weight -> Adam.

N > 1: the Collab-shaped graph fits one GPU ("small graphs stay on one GPU"), so ranks
run independent replicas (different seeds = different deletion requests), no data-path
collective; value = all ranks' epochs / max-over-ranks device time ("scaling": "weak").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = 'Del-training epochs/s on OGB-Collab shape'
NCU_DRAM_BYTES_SPMM_L2 = 108_790_272 + 45_636_864     # profiles/r1_spmm_batched_ncu_full.md (dram read + write, F=64 launch)
UNIT = 'epochs/s'


# --------------------------------------------------------------------------- utils
def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


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
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
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
def build_case(shape, seed, dev):
    """Synthetic inputs + masks (CUDA mask pipeline) + random-init model + z_ori."""
    import types
    from gnndelete_b200 import synthetic as S
    from gnndelete_b200 import masks as MK
    from gnndelete_b200 import models as M
    raw = S.make_graph(shape, seed=seed, device='cpu').to(dev)
    df = S.sample_df_mask(shape.num_edges, shape.num_deleted, seed=seed, device='cpu').to(dev)
    data = MK.build_unlearning_data(raw, df)
    neg = S.supplied_negatives(shape.num_nodes, int(data.df_mask.sum()), seed=seed + 1, device='cpu').to(dev)
    args = types.SimpleNamespace(in_dim=shape.in_dim, hidden_dim=shape.hidden_dim, out_dim=shape.out_dim)
    torch.manual_seed(seed)
    model = M.GCNDelete(args, data.sdf_node_1hop_mask, data.sdf_node_2hop_mask).to(dev)
    with torch.no_grad():
        z_ori = model.get_original_embeddings(data.x, data.train_pos_edge_index[:, data.dr_mask])
    return data, neg, model, z_ori


def spmm_algo_bytes(n, nnz, feat):
    """rowptr + col + deg^-1/2 + read N*F + write N*F, fp32, int32 ids."""
    return 4 * (n + 1) + 4 * nnz + 4 * n + 4 * n * feat + 4 * n * feat


def time_kernels(eng, steps):
    """Eager epochs with CUDA events around each aggregation launch (on the launch stream)."""
    from gnndelete_b200 import ops
    st = torch.cuda.current_stream()
    rec = {}

    def timed(name, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st); fn(); b.record(st)
        rec.setdefault(name, []).append((a, b))

    m, p = eng.model, eng.plan
    for _ in range(steps):
        ops.gemm_rows(eng.x, m.conv1.lin.weight.detach(), True, out=eng.h0, out_scale=p.dinv)
        timed('spmm_l1_f128', lambda: ops.spmm(p.fwd, eng.h0, out=eng.a1, row_scale=p.dinv, bias=m.conv1.bias.detach()))
        eng._layer1_done = True
        hoist, eng.hoist = eng.hoist, True
        # forward with the layer-2 aggregation bracketed
        w1 = m.deletion1.deletion_weight.detach(); w2 = m.deletion2.deletion_weight.detach()
        ops.gemm_rows(eng.a1, w1, False, out=eng.x1, rows=eng.rows1, relu_mask_out=eng.x1_bits)
        ops.copy_rows(eng.a1, eng.x1, eng.comp1)
        ops.gemm_rows(eng.x1, m.conv2.lin.weight.detach(), True, out=eng.h1, out_scale=p.dinv, relu_in=True)
        timed('spmm_l2_f64', lambda: ops.spmm(p.fwd, eng.h1, out=eng.a2, row_scale=p.dinv, bias=m.conv2.bias.detach()))
        ops.gemm_rows(eng.a2, w2, False, out=eng.z, rows=eng.rows2)
        ops.copy_rows(eng.a2, eng.z, eng.comp2)
        timed('edge_loss_fwd', lambda: eng.loss.forward(eng.z))
        timed('edge_loss_bwd_spmm', lambda: eng.loss.backward(eng.z, out=eng.dz))
        ops.gemm_tn_rows(eng.a2, eng.dz, rows=eng.rows2, out=eng.params[1].grad)
        ops.gemm_rows(eng.dz, w2, True, out=eng.da2, rows=eng.rows2)
        ops.copy_rows(eng.dz, eng.da2, eng.comp2)
        timed('spmm_bwd_f64', lambda: ops.spmm(p.bwd, eng.da2, out=eng.dh1, col_scale=p.dinv))
        ops.gemm_rows(eng.dh1, m.conv2.lin.weight.detach(), False, out=eng.dx1, rows=eng.rows1,
                      out_scale=p.dinv, gate=None if eng.bitmask else eng.x1, gate_bits=eng.x1_bits)
        ops.gemm_tn_rows(eng.a1, eng.dx1, rows=eng.rows1, out=eng.params[0].grad)
        eng.hoist = hoist
    torch.cuda.synchronize()
    return {k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in rec.items()}


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
    import types
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
        'impl': 'reference', 'metric': METRIC if shape.name.startswith('collab') else f'Del-training epochs/s on {shape.name} shape',
        'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(shape), 'where': 'host CPU',
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': 'PyTorch-Geometric is not installable here; the reference arm is the oracle restatement run with '
                "PyG's op sequence (eager PyTorch CPU autograd), parity unpinned",
    }
    print(json.dumps(line), flush=True)


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
    }


# ---------------------------------------------------------- row-partitioned config
def run_partitioned(args, shape, rank, local, world, dev, lib):
    """BASELINE config 5: GCNDelete on the power-law graph, 1-D row partitioned over the ranks with an
    NCCL all-gather halo exchange per layer (strong scaling: the graph is fixed, ranks split its rows).
    The graph is generated on the device with the same seed on every rank."""
    import types
    from gnndelete_b200 import masks as MK
    from gnndelete_b200 import models as M
    from gnndelete_b200 import synthetic as S
    from gnndelete_b200.dist import PartitionedGCNDeleteEngine
    from gnndelete_b200.engine import GCNDeleteEngine
    t_setup = time.perf_counter()
    raw = S.make_graph(shape, seed=42, device=dev, with_eval_edges=False)
    df = S.sample_df_mask(shape.num_edges, shape.num_deleted, seed=42, device=dev)
    data = MK.build_unlearning_data(raw, df)
    del raw
    neg = S.supplied_negatives(shape.num_nodes, int(data.df_mask.sum()), seed=43, device=dev)
    margs = types.SimpleNamespace(in_dim=shape.in_dim, hidden_dim=shape.hidden_dim, out_dim=shape.out_dim)
    torch.manual_seed(42)
    model = M.GCNDelete(margs, data.sdf_node_1hop_mask, data.sdf_node_2hop_mask).to(dev)
    with torch.no_grad():
        z_ori = model.get_original_embeddings(data.x, data.train_pos_edge_index[:, data.dr_mask])
    from gnndelete_b200 import graph as G
    G._GLOBAL_CACHE = G.PlanCache()                      # drop the dr-edge plan before the engines allocate
    torch.cuda.empty_cache()
    if world > 1:
        eng = PartitionedGCNDeleteEngine(model, data, neg, z_ori)
    else:
        eng = GCNDeleteEngine(model, data, neg, z_ori=z_ori, hoist_layer1=False)
    setup_s = time.perf_counter() - t_setup
    eng.epoch()                                       # first epoch: also builds the batch plans (one-time launches)
    c0 = lib.gd_launch_count()
    eng.epoch()
    launches = lib.gd_launch_count() - c0
    for _ in range(args.warmup):
        eng.epoch()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = max_over_ranks(timed_epochs(eng, args.steps, world), world, dev)
    clocks = sampler.stop() if rank == 0 else None
    losses = (eng.losses if world > 1 else eng.loss.losses).tolist()
    n = shape.num_nodes
    nnz = int(data.sdf_mask.sum()) + n
    halo = 3 * 4 * shape.out_dim * n + 4 * shape.hidden_dim * n          # bytes all-gathered per epoch (H0, H1, z, dA2)
    line = {
        'metric': 'Del-training epochs/s, row-partitioned power-law graph', 'value': args.steps / (ms / 1e3), 'unit': UNIT,
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'GCNDelete edge unlearning, {shape.name} power-law graph ({n} nodes / {shape.num_edges} '
                               f'directed edges / {shape.num_deleted} deleted), 128->128->64, 1-D row partition over '
                               f'{world} GPU(s), NCCL all-gather halo exchange, both convs recomputed',
                   'nnz_message_passing': nnz, 'halo_bytes_per_epoch_total': halo if world > 1 else 0,
                   'l2': 'feature matrices (>= 0.25 GB each) exceed the 126 MB L2, no flush'},
        'gpu_launches': int(launches * args.steps), 'launches_per_epoch': int(launches), 'clocks': clocks,
        'losses_last': losses, 'setup_s': setup_s, 'parallelism': f'row-partition x{world}',
    }
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


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
        return run_partitioned(args, shape, rank, local, world, dev, lib)

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

    # ---- per-kernel durations (eager launches, events on the launch stream) -> roofline
    peak, peak_src = load_peaks()
    eng_k = GCNDeleteEngine(model, data, neg, z_ori=z_ori, hoist_layer1=False, static_negatives=True)
    time_kernels(eng_k, 3)
    kt = time_kernels(eng_k, min(args.steps, 50))
    b64 = spmm_algo_bytes(n, nnz, shape.out_dim)
    b128 = spmm_algo_bytes(n, nnz, shape.hidden_dim)
    achieved = b64 / (kt['spmm_l2_f64'] * 1e-3) / 1e9
    roofline = {
        'kernel': 'gd::spmm_batched_kernel<16,false> (GCN layer-2 aggregation, F=64)', 'bound': 'hbm',
        'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel on this workload, one
        # `ncu --set full` capture (profiles/r1_spmm_batched_ncu_full.md); only meaningful for the Collab shape
        'traffic': NCU_DRAM_BYTES_SPMM_L2 if shape.name == 'collab' else None,
        'gather_ceiling_tbs': 18.0, 'gathered_tbs': (4 * nnz * shape.out_dim) / (kt['spmm_l2_f64'] * 1e-3) / 1e12,
        'peak_source': peak_src, 'algorithmic_bytes_per_launch': b64, 'kernel_ms': kt['spmm_l2_f64'],
        'bytes_gather_per_launch': b64 - 4 * n * shape.out_dim + 4 * nnz * shape.out_dim,
        'other_kernels': {
            'spmm_l1_f128': {'ms': kt['spmm_l1_f128'], 'algo_bytes': b128,
                             'frac': b128 / (kt['spmm_l1_f128'] * 1e-3) / 1e9 / peak},
            'spmm_bwd_f64': {'ms': kt['spmm_bwd_f64'], 'algo_bytes': b64,
                             'frac': b64 / (kt['spmm_bwd_f64'] * 1e-3) / 1e9 / peak},
            'edge_loss_fwd': {'ms': kt['edge_loss_fwd']}, 'edge_loss_bwd_spmm': {'ms': kt['edge_loss_bwd_spmm']},
        },
    }

    # ---- e2e: public API with host buffers — per step: pinned negatives -> device, plan
    #      refresh, epoch, losses -> host
    eng_e = GCNDeleteEngine(model, data, neg, z_ori=z_ori, hoist_layer1=False)
    eng_e.capture(warmup=2, dynamic_negatives=True)      # the graph rebuilds the negative incidence from a staging buffer
    neg_host = neg.cpu().pin_memory()
    out_host = torch.empty(3, dtype=torch.float32).pin_memory()

    def e2e_step():
        eng_e.set_negatives(neg_host)                     # pinned host -> device staging buffer (async H2D)
        out_host.copy_(eng_e.epoch(), non_blocking=True)  # one graph launch, then losses -> pinned host
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

    # the same steps with the transfers overlapped (EpochPipeline): H2D of step k and D2H of step k-1 ride a copy
    # stream while the graph of the neighbouring step runs; the host reads every step's losses, one step late
    from gnndelete_b200.engine import EpochPipeline
    pipe = EpochPipeline(eng_e)
    for i in range(3):
        pipe.result(pipe.submit(neg_host))
    barrier_sync(world)
    t0 = time.perf_counter()
    last = None
    for _ in range(e2e_steps):
        k = pipe.submit(neg_host)
        if last is not None:
            pipe.result(last)
        last = k
    pipe.result(last)
    barrier_sync(world)
    e2e_s = max_over_ranks(time.perf_counter() - t0, world, dev)
    e2e = {'value': world * e2e_steps / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': neg_host.numel() * 8,
           'd2h_bytes_per_step': 12, 'steps': e2e_steps, 'value_step_synchronous': world * e2e_steps / e2e_sync_s,
           'what': 'per step through GCNDeleteEngine / EpochPipeline: new negatives pinned-host -> device, in-graph rebuild '
                   'of the negative-pair incidence (radix sort), epoch (one CUDA-graph launch), losses -> pinned host and '
                   'read by the host; transfers of neighbouring steps overlap compute (value_step_synchronous: same steps '
                   'with a full host sync after every step)'}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(shape), 'where': 'B200',
        'value_hoisted': world * args.steps / (ms_h / 1e3), 'ms_per_step_hoisted': ms_h / args.steps,
        'roofline': roofline, 'e2e': e2e, 'gpu_launches': int(launches_per_epoch * args.steps),
        'launches_per_epoch': int(launches_per_epoch), 'clocks': clocks,
        'losses_last': losses, 'parallelism': 'replicas' if world > 1 else 'single',
    }
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
