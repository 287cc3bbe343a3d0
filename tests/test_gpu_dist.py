"""Row-partitioned GCNDelete epoch (gnndelete_b200.dist, BASELINE config 5) on the GPU.

* world 1 (any box): the partitioned engine without a process group - same plan / slot-space code, the DEC item
  kernel, the node-side loss kernel with fp32 and bf16 rows, the bf16-source aggregation - against the fp64 oracle and
  the single-GPU engine.
* world 2 over NCCL (boxes with >= 2 GPUs; skipped otherwise): two ranks, both wires, against the same oracle.
Tolerances: fp32 wire 1e-5; bf16 wire 2e-2 (north_star's stated allowance for the bf16 path)."""
import os
import sys

import pytest
import torch

from tests import util as U

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = {'fp32': 1e-5, 'bf16': 2e-2}


def _reference(scale):
    from oracle import unlearn as OU
    shape, raw, df, data, neg = U.make_case('cora', scale)
    om = U.oracle_model('gcn', shape, data, dtype=torch.float64)
    init = {k: v.float().clone() for k, v in om.state_dict().items()}
    d64 = data.clone(); d64.x = data.x.double()
    with torch.no_grad():
        zo = om.get_original_embeddings(d64.x, d64.train_pos_edge_index[:, d64.dr_mask])
    loss_o, lr_o, ll_o, _ = OU.edge_form_loss(om, d64, neg, zo)
    loss_o.backward()
    ref = dict(losses=torch.stack([loss_o, lr_o, ll_o]).detach(), g1=om.deletion1.deletion_weight.grad.clone(),
               g2=om.deletion2.deletion_weight.grad.clone())
    return shape, data, neg, zo, init, ref


def _check_rank(rank, world, wire, dev, group=None, scale=0.2, exchange='symm'):
    from gnndelete_b200 import models as M
    from gnndelete_b200.dist import PartitionedGCNDeleteEngine
    from gnndelete_b200.engine import GCNDeleteEngine
    shape, data, neg, zo, init, ref = _reference(scale)

    def fresh():
        m = M.GCNDelete(U.args_for(shape), data.sdf_node_1hop_mask, data.sdf_node_2hop_mask)
        m.load_state_dict(init)
        return m.to(dev)

    dd = data.clone().to(dev)
    m_part = fresh()
    eng = PartitionedGCNDeleteEngine(m_part, dd, neg.to(dev), zo.float().to(dev), group=group, wire=wire,
                                     world=world, rank=rank, exchange=exchange)
    if world > 1:
        assert eng.exchange == exchange, getattr(eng, 'exchange_error', None)
    tol = TOL[wire]
    eng.forward(); eng.backward()
    U.assert_close(eng.losses, ref['losses'], tol=tol, what=f'{wire} partitioned losses vs oracle')
    U.assert_close(m_part.deletion1.deletion_weight.grad, ref['g1'], tol=tol, what=f'{wire} partitioned dW1 vs oracle')
    U.assert_close(m_part.deletion2.deletion_weight.grad, ref['g2'], tol=tol, what=f'{wire} partitioned dW2 vs oracle')
    eng.adam_step()
    m_one = fresh()
    one = GCNDeleteEngine(m_one, dd, neg.to(dev), z_ori=zo.float().to(dev), hoist_layer1=False, static_negatives=True)
    one.epoch()
    for _ in range(3):
        a = eng.epoch().clone()
        b = one.epoch().clone()
        U.assert_close(a, b, tol=tol, what=f'{wire} partitioned vs single-GPU losses')
    wtol = 1e-5 if wire == 'fp32' else 2e-2
    U.assert_close(m_part.deletion1.deletion_weight, m_one.deletion1.deletion_weight, tol=wtol, what='W_del1 after 4 steps')
    U.assert_close(m_part.deletion2.deletion_weight, m_one.deletion2.deletion_weight, tol=wtol, what='W_del2 after 4 steps')
    return [float(v) for v in a.tolist()]


@pytest.mark.parametrize('wire', ['fp32', 'bf16'])
def test_partitioned_engine_world1(lib, wire):
    _check_rank(0, 1, wire, torch.device('cuda'))


def test_bf16_source_aggregation(lib):
    """gd_spmm_batched_bf16 == the fp32 aggregation of the bf16-rounded source (exact up to summation order), incl.
    row scale, bias, weights and long split rows."""
    from gnndelete_b200 import ops
    from gnndelete_b200.graph import plan_for
    shape, raw, df, data, neg = U.make_case('cora', 0.3)
    dev = torch.device('cuda')
    ei = data.train_pos_edge_index[:, data.sdf_mask].to(dev)
    plan = plan_for(ei, data.num_nodes, 'gcn')
    g = torch.Generator().manual_seed(3)
    for f in (64, 128):
        x = torch.randn(data.num_nodes, f, generator=g).to(dev)
        xb = ops.cast_bf16(x)
        assert torch.equal(xb, x.to(torch.bfloat16))                             # round to nearest even, like torch
        bias = torch.randn(f, generator=g).to(dev)
        ref = ops.spmm(plan.fwd, xb.float(), row_scale=plan.dinv, bias=bias)
        out = ops.spmm(plan.fwd, xb, row_scale=plan.dinv, bias=bias)
        U.assert_close(out, ref, tol=1e-6, what=f'bf16-source aggregation F={f}')
        refw = ops.spmm(plan.bwd, xb.float(), col_scale=plan.dinv)
        outw = ops.spmm(plan.bwd, xb, col_scale=plan.dinv)
        U.assert_close(outw, refw, tol=1e-6, what=f'weighted bf16-source aggregation F={f}')
        # against the unrounded source: the stated 2e-2 of the bf16 path
        U.assert_close(out, ops.spmm(plan.fwd, x, row_scale=plan.dinv, bias=bias), tol=2e-2, what='bf16 rounding')
        sc = torch.rand(data.num_nodes, generator=g).to(dev) + 0.5
        assert torch.equal(ops.cast_bf16(x, row_scale=sc), (x * sc.view(-1, 1)).to(torch.bfloat16))


def _nccl_worker(rank, world, port, wire, exchange, ret):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    from gnndelete_b200.dist import nccl_options
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev, pg_options=nccl_options())
    try:
        ret[rank] = _check_rank(rank, world, wire, dev, exchange=exchange)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('wire,exchange', [('fp32', 'symm'), ('bf16', 'symm'), ('bf16', 'nccl')])
def test_partitioned_engine_nccl_world2(lib, wire, exchange):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (run under gpurun --gpus 2)')
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    ret = ctx.Manager().dict()
    port = 29700 + (os.getpid() % 1000) + (1 if wire == 'bf16' else 0) + (2 if exchange == 'nccl' else 0)
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, wire, exchange, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert len(ret) == 2 and ret[0] == ret[1]            # both ranks report the same all-reduced losses
