"""Time gd_spmm on the Collab-shaped sdf edge set (F = 64 / 128, unweighted and weighted)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnndelete_b200 import ops, synthetic as S, masks as MK
from gnndelete_b200.graph import plan_for
dev = 'cuda'
shape = S.SHAPES['collab']
raw = S.make_graph(shape, seed=42).to(dev)
df = S.sample_df_mask(shape.num_edges, shape.num_deleted, seed=42).to(dev)
data = MK.build_unlearning_data(raw, df)
ei = data.train_pos_edge_index[:, data.sdf_mask].contiguous()
n = shape.num_nodes
plan = plan_for(ei, n, 'gcn')
nnz = plan.fwd.nnz
import gnndelete_b200.graph as G
print('nnz', nnz, 'segments', plan.fwd.num_seg, 'mode', 'batched' if G.BATCHED else 'rows', 'oversub', G.OVERSUB)
for f in (64, 128):
    x = torch.randn(n, f, device=dev); out = torch.empty(n, f, device=dev); bias = torch.randn(f, device=dev)
    for name, kw in (('plain', dict(row_scale=plan.dinv, bias=bias)), ('col_scale', dict(col_scale=plan.dinv))):
        fn = lambda: ops.spmm(plan.fwd, x, out=out, **kw)
        for _ in range(5): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50): fn()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 50
        algo = 4 * (n + 1) + 4 * nnz + 4 * n + 8 * n * f
        print(f'F={f:3d} {name:9s} {ms*1e3:7.1f} us   algo {algo/ms/1e6:7.0f} GB/s ({algo/ms/1e6/6449.1:.3f} of measured HBM peak)   gather {nnz*f*4/ms/1e9:5.2f} TB/s')
