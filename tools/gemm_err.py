"""Measure the numerical error of the tcgen05 3xTF32 GEMM against fp64 (and the SIMT fp32 path)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnndelete_b200 import ops

torch.manual_seed(0)
for (m, k, n) in [(4096, 128, 128), (4096, 128, 64), (4096, 64, 64)]:
    a = torch.randn(m, k, device='cuda'); b = torch.randn(n, k, device='cuda') / k ** 0.5
    ref = a.double() @ b.double().t()
    for backend in ('simt', 'tc'):
        ops.GEMM_BACKEND = backend
        out = ops.gemm_rows(a, b, True).double()
        err = (out - ref)
        print(f'{m}x{k}x{n} {backend:5s} max|err|/max|ref| = {err.abs().max().item() / ref.abs().max().item():.3e}  '
              f'rms err/rms ref = {err.pow(2).mean().sqrt().item() / ref.pow(2).mean().sqrt().item():.3e}  '
              f'mean err/rms ref (bias) = {err.mean().item() / ref.pow(2).mean().sqrt().item():.3e}')
    # positive operands expose accumulation bias
    a = torch.rand(m, k, device='cuda'); b = torch.rand(n, k, device='cuda')
    ref = a.double() @ b.double().t()
    for backend in ('simt', 'tc'):
        ops.GEMM_BACKEND = backend
        out = ops.gemm_rows(a, b, True).double()
        err = out - ref
        print(f'   positive {backend:5s} max rel = {(err.abs() / ref).max().item():.3e}  mean rel (bias) = {(err / ref).mean().item():.3e}')
