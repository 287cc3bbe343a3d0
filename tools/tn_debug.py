"""Structured-input probe of gd_gemm_tn_rows_tc: A = identity rows, G[r, n] = 1000 r + n, so out[f, n] must equal G[f, n]."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnndelete_b200 import ops
torch.set_printoptions(linewidth=200, sci_mode=False)
m = 128
a = torch.eye(m, 128, device='cuda')
g = (torch.arange(m, device='cuda').float()[:, None] * 1000 + torch.arange(128, device='cuda').float()[None, :])
out = ops.gemm_tn_rows(a, g)
ref = a.double().t() @ g.double()
print('max abs err', (out.double() - ref).abs().max().item(), 'ref max', ref.abs().max().item())
for f in (0, 1, 2, 7, 8, 9, 31, 32, 33, 64, 127):
    print(f, out[f, :10].tolist(), '| cols 32..35', out[f, 32:36].tolist())
nz = (out != 0).sum().item()
print('nonzeros', nz, 'of', out.numel())
# where did G[r, n] land?  decode every output value v = 1000 r + n
v = out.round().long()
r_src, n_src = v // 1000, v % 1000
bad = (out.double() - ref).abs() > 0.5
print('bad entries', int(bad.sum()))
idx = bad.nonzero()[:20]
for f, n in idx.tolist():
    print(f'out[{f},{n}] = {out[f, n].item():.1f}  (looks like G[{r_src[f, n].item()},{n_src[f, n].item()}])  expected {ref[f, n].item():.1f}')
