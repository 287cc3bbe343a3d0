"""Time the fused decoder + DEC/NI loss (+ dz) of the Collab-shaped epoch: CUDA-graph replays, CUDA events.
usage: [GD_NL_CFG=16x1@4] python tools/loss_bench.py [workload] [reps]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from gnndelete_b200 import synthetic as S
from gnndelete_b200.losses import EdgeLossPlan

wl = sys.argv[1] if len(sys.argv) > 1 else 'collab'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = torch.device('cuda')
shape = S.SHAPES[wl]
data, neg, model, z_ori = B.build_case(shape, 42, dev)
ei = data.train_pos_edge_index
sdf = ei[:, data.sdf_mask]
ni = sdf[:, sdf[0] < sdf[1]]
z = (z_ori + 0.01 * torch.randn_like(z_ori)).contiguous()
out = {}
for static in (True, False):
    plan = EdgeLossPlan(ei[:, data.df_mask], neg, ni, shape.num_nodes, z_ori=z_ori, static_negatives=static)
    dz = torch.empty_like(z)
    for _ in range(3):
        plan.forward(z, dz_out=dz); plan.backward(z, out=dz)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        plan.forward(z, dz_out=dz); plan.backward(z, out=dz)
    for _ in range(5):
        g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record(); torch.cuda.synchronize()
    P = plan.num_pairs
    out['static' if static else 'dynamic'] = {'us': 1e3 * a.elapsed_time(b) / reps, 'pairs': P, 'mode': plan.mode,
                                               'entries': int(plan.nnz_fixed), 'losses': plan.losses.tolist()}
print(json.dumps({'cfg': os.environ.get('GD_NL_CFG', 'default'), **out}))
