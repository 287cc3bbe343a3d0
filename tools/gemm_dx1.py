"""Which epilogue option of the 64 -> 128 gathered-row GEMM (dX1) costs what."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnndelete_b200 import ops
dev = 'cuda'
torch.manual_seed(0)
N = 235368
x64 = torch.randn(N, 64, device=dev); x128 = torch.randn(N, 128, device=dev); o128 = torch.empty(N, 128, device=dev); o64 = torch.empty(N, 64, device=dev)
w = torch.randn(64, 128, device=dev); w128 = torch.randn(128, 128, device=dev)
sc = torch.rand(N, device=dev) + 0.5
rows = torch.nonzero(torch.rand(N, device=dev) < 0.88).squeeze(1).to(torch.int32)
bits = torch.randint(-2**31, 2**31 - 1, (N, 4), dtype=torch.int32, device=dev)
mask = torch.zeros(N, 4, dtype=torch.int32, device=dev)
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
cases = {
    '64->128 rows': lambda: ops.gemm_rows(x64, w, False, out=o128, rows=rows),
    '64->128 rows scale': lambda: ops.gemm_rows(x64, w, False, out=o128, rows=rows, out_scale=sc),
    '64->128 rows gatebits': lambda: ops.gemm_rows(x64, w, False, out=o128, rows=rows, gate_bits=bits),
    '64->128 rows scale gatebits': lambda: ops.gemm_rows(x64, w, False, out=o128, rows=rows, out_scale=sc, gate_bits=bits),
    '64->128 all rows': lambda: ops.gemm_rows(x64, w, False, out=o128),
    '128->128 rows': lambda: ops.gemm_rows(x128, w128, False, out=o128, rows=rows),
    '128->128 rows maskbits': lambda: ops.gemm_rows(x128, w128, False, out=o128, rows=rows, relu_mask_out=mask),
    '128->128 all rows': lambda: ops.gemm_rows(x128, w128, False, out=o128),
}
for name, fn in cases.items():
    print(f'{name:32s} {t(fn):7.1f} us', flush=True)
