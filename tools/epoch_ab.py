"""value / value_hoisted of the Collab epoch only (A/B of library builds): python tools/epoch_ab.py [steps]"""
import os, sys, types, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
from gnndelete_b200 import synthetic as S
from gnndelete_b200.engine import GCNDeleteEngine
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device('cuda', 0)
shape = S.SHAPES['collab']
data, neg, model, z_ori = B.build_case(shape, 42, dev)
out = {}
for name, hoist in (('value', False), ('hoisted', True)):
    eng = GCNDeleteEngine(model, data, neg, z_ori=z_ori, hoist_layer1=hoist, static_negatives=True)
    eng.capture(warmup=2)
    for _ in range(10):
        eng.epoch()
    ts = [B.timed_epochs(eng, steps, 1) for _ in range(3)]
    out[name] = [round(1e3 * t / steps, 2) for t in ts]
    del eng
print(os.environ.get('GD_LIB_TAG', ''), 'us/epoch', out)
