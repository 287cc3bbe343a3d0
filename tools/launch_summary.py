"""Per-kernel shares of ONE epoch from an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py.
usage: python tools/launch_summary.py launches.csv [first_kernel_regex] > summary.md
The epoch is located as the last run of launches that starts with the layer-1 GEMM (gemm_rows_tc_kernel<2>) and ends
with the second Adam bump."""
import csv, re, sys
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
for r in rd:
    if len(r) <= iv: continue
    v = float(r[iv].replace(',', ''))
    u = r[iu]
    ns = v * {'ns': 1, 'us': 1e3, 'usecond': 1e3, 'nsecond': 1, 'ms': 1e6, 'msecond': 1e6}.get(u, 1)
    name = re.sub(r'\(.*', '', r[ik]).replace('void ', '').replace('(int)', '').replace('(bool)', '')
    rows.append((name, ns))
# epochs end with the second Adam bump: cut the launch list there and take the last segment that looks like the recompute
# epoch `value` times (three forward aggregation launches - two 64-wide layer-1 passes + layer 2 - and the backward one, the
# fused loss, no dense-NI kernels, no set-up launches)
ends = [i for i, (n, _) in enumerate(rows) if 'adam_bump' in n]
segs, prev = [], -1
for k in range(1, len(ends), 2):
    segs.append(rows[prev + 1:ends[k] + 1]); prev = ends[k]
best = None
for seg in reversed(segs):
    if sum('spmm_batched_kernel' in n for n, _ in seg) == 4 and any('node_loss' in n for n, _ in seg) \
            and not any('dense_ni' in n or 'pair_scatter' in n for n, _ in seg) and len(seg) <= 40:
        best = seg; break
if best is None:
    sys.exit('no complete recompute epoch in the launch list')
agg = {}
for n, ns in best:
    c, t = agg.get(n, (0, 0.0)); agg[n] = (c + 1, t + ns)
tot = sum(t for _, t in agg.values())
print('| kernel | launches/epoch | ns | share |\n|---|---:|---:|---:|')
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'| `{n}` | {c} | {t:.0f} | {100 * t / tot:.1f}% |')
print(f'| **total** | {sum(c for c, _ in agg.values())} | {tot:.0f} | 100% |')
