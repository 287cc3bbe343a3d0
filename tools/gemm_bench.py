"""Time the gathered-row GEMM variants of one Collab-shaped epoch in isolation (CUDA events)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnndelete_b200 import ops

N, M = 235368, 150000
dev = 'cuda'
torch.manual_seed(0)
rows = torch.randperm(N, device=dev)[:M].sort()[0].int()
x128 = torch.randn(N, 128, device=dev); x64 = torch.randn(N, 64, device=dev)
o128 = torch.empty(N, 128, device=dev); o64 = torch.empty(N, 64, device=dev)
w128 = torch.randn(128, 128, device=dev); w64_128 = torch.randn(64, 128, device=dev); w64 = torch.randn(64, 64, device=dev)
dinv = torch.rand(N, device=dev)

def bench(name, fn, bytes_):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print(f'{name:48s} {ms*1e3:8.1f} us   {bytes_/ms/1e6:8.1f} GB/s')

for backend in ('tc', 'simt'):
    ops.GEMM_BACKEND = backend
    print('---', backend)
    bench('X.W1^T all rows 128->128 +scale', lambda: ops.gemm_rows(x128, w128, True, out=o128, out_scale=dinv), N*1024)
    bench('Del1 gathered 128->128', lambda: ops.gemm_rows(x128, w128, False, out=o128, rows=rows), M*1024)
    bench('r.W2^T all rows 128->64 relu_in +scale', lambda: ops.gemm_rows(x128, w64_128, True, out=o64, out_scale=dinv, relu_in=True), N*768)
    bench('Del2 gathered 64->64', lambda: ops.gemm_rows(x64, w64, False, out=o64, rows=rows), M*512)
    bench('dx1 gathered 64->128 +scale (no gate)', lambda: ops.gemm_rows(x64, w64_128, False, out=o128, rows=rows, out_scale=dinv), M*768)
    bench('dx1 gathered 64->128 +scale +gate', lambda: ops.gemm_rows(x64, w64_128, False, out=o128, rows=rows, out_scale=dinv, gate=x128), M*1280)
    bench('all rows 64->128 +gate', lambda: ops.gemm_rows(x64, w64_128, False, out=o128, gate=x128), N*1280)
