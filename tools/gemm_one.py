"""One 128->128 all-rows tcgen05 GEMM (16 tiles per CTA), for ncu."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnndelete_b200 import ops
m = 148 * 128 * 16
x = torch.randn(m, 128, device='cuda'); o = torch.empty(m, 128, device='cuda'); w = torch.randn(128, 128, device='cuda')
for _ in range(3):
    ops.gemm_rows(x, w, True, out=o)
torch.cuda.synchronize()
