"""Fixed cost vs streaming rate of the tcgen05 gathered-row GEMMs: time vs number of 128-row tiles per CTA."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnndelete_b200 import ops
dev = 'cuda'
torch.manual_seed(0)
N = 148 * 128 * 16
x128 = torch.randn(N, 128, device=dev); x64 = torch.randn(N, 64, device=dev)
o128 = torch.empty(N, 128, device=dev); o64 = torch.empty(N, 64, device=dev)
w128 = torch.randn(128, 128, device=dev); w64_128 = torch.randn(64, 128, device=dev); w64 = torch.randn(64, 64, device=dev)
g128 = torch.zeros(128, 128, device=dev); g64 = torch.zeros(64, 64, device=dev)

def t(fn, reps=20):
    """us per call, 20 calls captured in one CUDA graph (no host launch overhead in the timing)."""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3

for tiles in (1, 2, 4, 8, 12, 16):
    m = 148 * 128 * tiles
    a = t(lambda: ops.gemm_rows(x128[:m], w128, True, out=o128[:m]))
    b = t(lambda: ops.gemm_rows(x128[:m], w64_128, True, out=o64[:m], relu_in=True))
    c = t(lambda: ops.gemm_rows(x64[:m], w64, False, out=o64[:m]))
    d = t(lambda: ops.gemm_tn_rows(x128[:m], o128[:m], out=g128))
    e = t(lambda: ops.gemm_tn_rows(x64[:m], o64[:m], out=g64))
    print(f'tiles/CTA {tiles:3d}  m {m:8d}   128->128 {a:7.1f} us ({m*1024/a/1e6:6.2f} TB/s)   128->64 {b:7.1f} us ({m*768/b/1e6:6.2f})   '
          f'64->64 {c:7.1f} us ({m*512/c/1e6:6.2f})   tn128 {d:7.1f} us ({m*1024/d/1e6:6.2f})   tn64 {e:7.1f} us ({m*512/e/1e6:6.2f})')
