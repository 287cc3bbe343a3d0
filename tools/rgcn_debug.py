"""Edge path vs transform / tile paths at the full BioKG shape: per-direction max error and the rows that differ."""
import dataclasses, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnndelete_b200 import graph as G, ops, synthetic as S
DEV = 'cuda'
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
shape = dataclasses.replace(S.SHAPES['biokg'].scaled(scale), num_edge_type=51)
raw = S.make_graph(shape, seed=42)
ei = torch.cat([raw.train_pos_edge_index, raw.train_pos_edge_index.flip(0)], 1).to(DEV)
et = torch.cat([raw.train_edge_type, raw.train_edge_type + 51]).to(DEV)
n = shape.num_nodes
g = torch.Generator().manual_seed(2)
plan = G.GraphPlan(ei, n, False, et, 102)
for (fin, fout) in ((128, 64), (128, 128)):
    w = (torch.randn(102, 4, fin // 4, fout // 4, generator=g) * 0.2).to(DEV)
    root = (torch.randn(fin, fout, generator=g) * 0.1).to(DEV)
    bias = torch.randn(fout, generator=g).to(DEV)
    x = torch.randn(n, fin, generator=g).to(DEV)
    gout = torch.randn(n, fout, generator=g).to(DEV)
    res = {}
    for mode in ('transform', 'edge'):
        ops.RGCN_MODE = mode
        res[mode] = (ops.rgcn_conv(plan, x, w, root, bias).clone(), ops.rgcn_conv(plan, gout, w, root, None, transposed=True).clone())
    for d, name in ((0, 'fwd'), (1, 'transposed')):
        a, b = res['edge'][d].double(), res['transform'][d].double()
        err = (a - b).abs()
        rowerr = err.max(1).values
        bad = (rowerr > 1e-4 * b.abs().max()).nonzero().squeeze(1)
        print(f'{fin}->{fout} {name}: max rel err {float(err.max() / b.abs().max()):.3e}; bad rows {bad.numel()} first {bad[:10].tolist()}', flush=True)
        if bad.numel():
            deg = (plan.bwd.rowptr[1:] - plan.bwd.rowptr[:-1]) if d else (plan.fwd.rowptr[1:] - plan.fwd.rowptr[:-1])
            print('   degrees of bad rows', deg[bad[:10]].tolist(), 'tiles', (bad[:10] // 16).tolist())
    # timing
    for mode in ('transform', 'edge'):
        ops.RGCN_MODE = mode
        for tr, inp in ((False, x), (True, gout)):
            for _ in range(2): ops.rgcn_conv(plan, inp, w, root, None if tr else bias, transposed=tr)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5): ops.rgcn_conv(plan, inp, w, root, None if tr else bias, transposed=tr)
            b.record(); torch.cuda.synchronize()
            print(f'   {mode:9s} {"transposed" if tr else "forward":10s} {fin}->{fout}: {a.elapsed_time(b) / 5:.3f} ms', flush=True)
