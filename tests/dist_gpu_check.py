"""Multi-GPU check of the row-partitioned epoch (run under torchrun on >= 2 GPUs, e.g.
    gpurun --gpus 2 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
                        --master-port 29511 tests/dist_gpu_check.py'
Every rank runs PartitionedGCNDeleteEngine over NCCL; the losses and the all-reduced Del gradients of
3 epochs are compared with the single-GPU GCNDeleteEngine (same kernels, no partition) and the fp64 oracle."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, local, world = int(os.environ['RANK']), int(os.environ['LOCAL_RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    from gnndelete_b200 import models as M
    from gnndelete_b200.dist import PartitionedGCNDeleteEngine
    from gnndelete_b200.engine import GCNDeleteEngine
    from oracle import unlearn as OU
    from tests import util as U

    shape, raw, df, data, neg = U.make_case('cora', 0.2)
    om = U.oracle_model('gcn', shape, data, dtype=torch.float64)
    init = {k: v.float().clone() for k, v in om.state_dict().items()}
    d64 = data.clone(); d64.x = data.x.double()
    with torch.no_grad():
        zo = om.get_original_embeddings(d64.x, d64.train_pos_edge_index[:, d64.dr_mask])
    loss_o, lr_o, ll_o, _ = OU.edge_form_loss(om, d64, neg, zo)
    loss_o.backward()

    def fresh():
        m = M.GCNDelete(U.args_for(shape), data.sdf_node_1hop_mask, data.sdf_node_2hop_mask)
        m.load_state_dict(init)
        return m.to(dev)

    dd = data.clone().to(dev)
    m_part = fresh()
    eng = PartitionedGCNDeleteEngine(m_part, dd, neg.to(dev), zo.float().to(dev))
    m_one = fresh()
    one = GCNDeleteEngine(m_one, dd, neg.to(dev), z_ori=zo.float().to(dev), hoist_layer1=False)
    # first step against the oracle at 1e-5
    l_part = eng.forward().clone(); eng.backward()
    U.assert_close(l_part, torch.stack([loss_o, lr_o, ll_o]), what='partitioned losses vs oracle')
    U.assert_close(m_part.deletion1.deletion_weight.grad, om.deletion1.deletion_weight.grad, what='partitioned dW1 vs oracle')
    U.assert_close(m_part.deletion2.deletion_weight.grad, om.deletion2.deletion_weight.grad, what='partitioned dW2 vs oracle')
    eng.adam_step()
    one.epoch()
    for _ in range(3):
        a = eng.epoch().clone()
        b = one.epoch().clone()
        U.assert_close(a, b, tol=1e-5, what='partitioned vs single-GPU losses')
    U.assert_close(m_part.deletion1.deletion_weight, m_one.deletion1.deletion_weight, tol=1e-5, what='W_del1 after 4 steps')
    U.assert_close(m_part.deletion2.deletion_weight, m_one.deletion2.deletion_weight, tol=1e-5, what='W_del2 after 4 steps')
    dist.barrier()
    if rank == 0:
        print(f'dist_gpu_check ok: world={world} losses={[round(float(v), 6) for v in a.tolist()]}')
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
