"""Chained dX / dW kernel (gemm_dxdw_wt.cu) at the Collab epoch shape: against the two kernels it replaces, or --once (for ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnndelete_b200 import ops
dev = 'cuda'
torch.manual_seed(0)
N = 235868
x = torch.randn(N, 64, device=dev); a = torch.randn(N, 128, device=dev); w = torch.randn(64, 128, device=dev)
sc = torch.rand(N, device=dev) + 0.5
rows = torch.nonzero(torch.rand(N, device=dev) < 0.88).squeeze(1).to(torch.int32)
bits = torch.randint(-2**31, 2**31 - 1, (N, 4), dtype=torch.int32, device=dev)
out = torch.zeros(128, 128, device=dev); dx = torch.zeros(N, 128, device=dev)
fn = lambda: ops.gemm_dxdw(x, w, False, a, rows=rows, in_scale=sc, gate_bits=bits, out=out)
if '--once' in sys.argv:
    for _ in range(3): fn()
    torch.cuda.synchronize(); sys.exit(0)

def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

def two():
    ops.gemm_rows(x, w, False, out=dx, rows=rows, out_scale=sc, gate_bits=bits)
    ops.gemm_tn_rows(a, dx, rows=rows, out=out)
print(f'rows {rows.numel()}  two kernels {t(two):.1f} us', flush=True)
print(f'chained {t(fn):.1f} us', flush=True)
