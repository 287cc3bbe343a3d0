"""Phase knobs of the tcgen05 row GEMM (GD_TC_DEBUG, measurement only) + the weight-gradient kernel at the sweep size."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnndelete_b200 import ops
dev = 'cuda'
torch.manual_seed(0)
T = 16
N = 148 * 128 * T
x128 = torch.randn(N, 128, device=dev); x64 = torch.randn(N, 64, device=dev)
o128 = torch.empty(N, 128, device=dev); o64 = torch.empty(N, 64, device=dev)
w128 = torch.randn(128, 128, device=dev); w64_128 = torch.randn(64, 128, device=dev); w64 = torch.randn(64, 64, device=dev)
g128 = torch.zeros(128, 128, device=dev); g64 = torch.zeros(64, 64, device=dev)

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

for knob in (0, 1, 2, 4, 8, 3, 10, 30, 32):
    os.environ['GD_TC_DEBUG'] = str(knob)
    a = t(lambda: ops.gemm_rows(x128, w128, True, out=o128))
    b = t(lambda: ops.gemm_rows(x128, w64_128, True, out=o64, relu_in=True))
    c = t(lambda: ops.gemm_rows(x64, w64, False, out=o64))
    print(f'knob {knob:3d}   128->128 {a:7.1f} us   128->64 {b:7.1f}   64->64 {c:7.1f}', flush=True)
os.environ['GD_TC_DEBUG'] = '0'
d = t(lambda: ops.gemm_tn_rows(x128, o128, out=g128))
e = t(lambda: ops.gemm_tn_rows(x64, o64, out=g64))
print(f'tn128 {d:7.1f} us ({N*1024/d/1e6:6.2f} TB/s)   tn64 {e:7.1f} us ({N*512/e/1e6:6.2f})')
