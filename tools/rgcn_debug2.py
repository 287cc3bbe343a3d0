"""RGCNDelete fwd/bwd at the full BioKG shape under GD_RGCN=edge vs transform: which tensors differ."""
import dataclasses, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnndelete_b200 import masks as MK, models as M, ops, synthetic as S
from tests import util as U
DEV = 'cuda'
net = 51
shape = dataclasses.replace(S.SHAPES['biokg'].scaled(float(sys.argv[1]) if len(sys.argv) > 1 else 1.0), num_edge_type=net)
raw = S.make_graph(shape, seed=42).to(DEV)
df = S.sample_df_mask(shape.num_edges, shape.num_deleted, seed=42).to(DEV)
data = MK.build_unlearning_data(raw, df, num_edge_type=net)
torch.manual_seed(0)
m = M.RGCNDelete(U.args_for(shape), shape.num_nodes, net, data.sdf_node_1hop_mask, data.sdf_node_2hop_mask)
U.randomize(m, 0)
m = m.to(DEV)
ei, et = data.edge_index[:, data.dr_mask].contiguous(), data.edge_type[data.dr_mask].contiguous()
g = torch.Generator().manual_seed(1)
t1 = torch.randn(shape.num_nodes, 128, generator=g).to(DEV)
t2 = torch.randn(shape.num_nodes, 64, generator=g).to(DEV)
res = {}
for mode in ('transform', 'edge', 'edge'):
    ops.RGCN_MODE = mode
    m.zero_grad()
    m._conv1_cache = M._FrozenCache()
    z1, z2 = m(data.x, ei, et, return_all_emb=True)
    z1.retain_grad()
    ((z1 - t1) ** 2).mean().backward(retain_graph=True)
    g1a = m.deletion1.deletion_weight.grad.clone()
    ((z2 - t2) ** 2).mean().backward()
    key = mode if mode not in res else mode + '2'
    res[key] = dict(z1=z1.detach().clone(), z2=z2.detach().clone(), dz1=z1.grad.clone(), g1a=g1a,
                    g1=m.deletion1.deletion_weight.grad.clone(), g2=m.deletion2.deletion_weight.grad.clone())
for k in res['edge']:
    a, b, c = res['edge'][k].double(), res['transform'][k].double(), res['edge2'][k].double()
    print(f'{k:5s} edge vs transform rel err {float((a - b).abs().max() / b.abs().max()):.3e}   edge vs edge (repeat) {float((a - c).abs().max() / a.abs().max()):.3e}', flush=True)
d = (res['edge']['dz1'] - res['transform']['dz1']).abs().max(1).values
bad = (d > 1e-4 * res['transform']['dz1'].abs().max()).nonzero().squeeze(1)
print('rows of dz1 that differ:', bad.numel(), bad[:20].tolist())
if bad.numel():
    from gnndelete_b200.graph import plan_for
    plan = plan_for(ei, shape.num_nodes, 'rgcn', et, 102)
    deg = plan.bwd.rowptr[1:] - plan.bwd.rowptr[:-1]
    print('their transposed degrees:', deg[bad[:20]].tolist())
    print('max transposed degree', int(deg.max()), 'num rows with deg > 512:', int((deg > 512).sum()))
