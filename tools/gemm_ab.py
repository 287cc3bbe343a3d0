"""Row GEMM A/B: weights-in-TMEM kernel (gemm_tc_wt.cu, default) vs the shared-memory-ring kernel (GD_GEMM_ROWS=ring), streaming sizes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnndelete_b200 import ops
dev = 'cuda'
torch.manual_seed(0)
N = 148 * 128 * 16
x128 = torch.randn(N, 128, device=dev); x64 = torch.randn(N, 64, device=dev)
o128 = torch.empty(N, 128, device=dev); o64 = torch.empty(N, 64, device=dev)
w128 = torch.randn(128, 128, device=dev); w64_128 = torch.randn(64, 128, device=dev); w64 = torch.randn(64, 64, device=dev); w128_64 = torch.randn(64, 128, device=dev)
sc = torch.rand(N, device=dev) + 0.5
rows = torch.nonzero(torch.rand(N, device=dev) < 0.88).squeeze(1).to(torch.int32)
bits = torch.zeros(N, 4, dtype=torch.int32, device=dev)

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
    '128->128 scale (xw1)': lambda: ops.gemm_rows(x128, w128, True, out=o128, out_scale=sc),
    '128->128 rows+maskbits (del1)': lambda: ops.gemm_rows(x128, w128, False, out=o128, rows=rows, relu_mask_out=bits),
    '128->64 relu_in scale (xw2)': lambda: ops.gemm_rows(x128, w64_128, True, out=o64, out_scale=sc, relu_in=True),
    '64->64 rows (del2)': lambda: ops.gemm_rows(x64, w64, False, out=o64, rows=rows),
    '64->128 rows scale gatebits (dx1)': lambda: ops.gemm_rows(x64, w128_64, False, out=o128, rows=rows, out_scale=sc, gate_bits=bits),
}
for name, fn in cases.items():
    res = {}
    for mode in ('wt', 'ring'):
        os.environ['GD_GEMM_ROWS'] = mode
        res[mode] = t(fn)
    print(f'{name:36s} wt {res["wt"]:7.1f} us   ring {res["ring"]:7.1f} us   x{res["ring"] / res["wt"]:.2f}', flush=True)
os.environ.pop('GD_GEMM_ROWS', None)
g128 = torch.zeros(128, 128, device=dev); g64 = torch.zeros(64, 64, device=dev)
tn = {'tn 128x128 rows (dw1)': lambda: ops.gemm_tn_rows(x128, o128, rows=rows, out=g128),
      'tn 64x64 rows (dw2)': lambda: ops.gemm_tn_rows(x64, o64, rows=rows, out=g64)}
for name, fn in tn.items():
    res = {}
    for mode in ('wt', 'ring'):
        os.environ['GD_GEMM_TN'] = mode
        res[mode] = t(fn)
    print(f'{name:36s} wt {res["wt"]:7.1f} us   ring {res["ring"]:7.1f} us   x{res["ring"] / res["wt"]:.2f}', flush=True)
os.environ.pop('GD_GEMM_TN', None)
# chained dX1 -> dW_del1 (gemm_dxdw_wt.cu) vs the two kernels it replaces
def two():
    ops.gemm_rows(x64, w128_64, False, out=o128, rows=rows, out_scale=sc, gate_bits=bits)
    ops.gemm_tn_rows(x128, o128, rows=rows, out=g128)
bits.fill_(-1)
g128b = torch.zeros(128, 128, device=dev)
two(); ops.gemm_dxdw(x64, w128_64, False, x128, rows=rows, in_scale=sc, gate_bits=bits, out=g128b)
torch.cuda.synchronize()
print('chained vs two kernels: max rel diff', ((g128b - g128).abs().max() / g128.abs().max()).item(), flush=True)
print(f'{"dx1 + dw1 two kernels":36s} {t(two):7.1f} us   chained {t(lambda: ops.gemm_dxdw(x64, w128_64, False, x128, rows=rows, in_scale=sc, gate_bits=bits, out=g128b)):7.1f} us', flush=True)
