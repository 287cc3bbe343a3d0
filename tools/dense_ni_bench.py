"""Time of the dense-block NI (gd_dense_ni_fwd_bwd) at the full Cora shape (S2 x S2 pairs), CUDA events over graph replays."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnndelete_b200 import masks as MK, synthetic as S
from gnndelete_b200.losses import DenseNIPlan
dev = 'cuda'
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
shape = S.SHAPES['cora'].scaled(scale)
raw = S.make_graph(shape, seed=42, device='cpu').to(dev)
df = S.sample_df_mask(shape.num_edges, shape.num_deleted, seed=42, device='cpu').to(dev)
data = MK.build_unlearning_data(raw, df)
n = data.num_nodes
torch.manual_seed(1)
z_ori = torch.randn(n, 64, device=dev) * 0.3
logits = z_ori @ z_ori.t()
plan = DenseNIPlan(data.sdf_node_2hop_mask, data.train_pos_edge_index[:, data.df_mask], logits, n, 64, weight=0.5)
del logits
z = (z_ori + 0.05 * torch.randn_like(z_ori)).contiguous()
dz = torch.zeros_like(z)
for _ in range(3): plan.forward_backward(z, dz)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(5): plan.forward_backward(z, dz)
g.replay(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); g.replay(); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
ns = plan.n_s
flops = 2 * 2.0 * ns * ns * 64            # two contractions over the full S x S block
print(f'dense NI: n_s {ns}  pairs {plan.num_pairs}  {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s (fp32-equivalent, full block)  '
      f'target block {ns * ns * 4 / 1e9:.2f} GB -> {ns * ns * 4 / ms / 1e6:.0f} GB/s')
