"""Graph-replayed time of parts of the Collab epoch: full epoch, without Adam, forward only, forward + loss."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
from gnndelete_b200 import synthetic as S
from gnndelete_b200.engine import GCNDeleteEngine
dev = torch.device('cuda', 0)
shape = S.SHAPES['collab']
data, neg, model, z_ori = B.build_case(shape, 42, dev)
eng = GCNDeleteEngine(model, data, neg, z_ori=z_ori, hoist_layer1=False, static_negatives=True)
for _ in range(3):
    eng.forward_backward(); eng.adam_step()
torch.cuda.synchronize()
def timed(fn, reps=50):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        a.record()
        for _ in range(reps): g.replay()
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / reps * 1e3)
    return round(best, 1)
out = {}
out['epoch'] = timed(lambda: (eng.forward_backward(), eng.adam_step()))
out['no_adam'] = timed(lambda: eng.forward_backward())
out['forward_only(no loss)'] = timed(lambda: (eng.layer1(), None))
def fwd():
    eng.forward()
out['forward+loss'] = timed(fwd)
out['adam_only'] = timed(lambda: eng.adam_step())
print(out)
