"""Measured DRAM traffic of the aggregation launches -> profiles/ncu_traffic.json (read by bench.py's roofline.traffic).
usage: python tools/ncu_traffic.py <report.ncu-rep of `ncu --set full ... -k regex:spmm_batched_kernel python bench.py --steps 2 --warmup 3`>
       python tools/ncu_traffic.py --dense-ni <report.ncu-rep of the dense_ni_tc_kernel capture>   (merges one entry, keeps the rest)
The report must hold, in launch order, the aggregation launches of one eager or captured epoch with both convs
recomputed: layer-1 (two 64-wide passes, template <16, 0, 3>), layer-2 (<16, 0, 3>), transpose-backward (<16, 1, 0>)."""
import csv, io, json, os, subprocess, sys

dense_only = sys.argv[1] == '--dense-ni'
rep = sys.argv[2] if dense_only else sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
units = rows[1]


def to_bytes(v, u):
    f = float(v)
    return f * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]


launches = []
for r in rows[2:]:
    name = r[ix['Kernel Name']]
    rd = to_bytes(r[ix['dram__bytes_read.sum']], units[ix['dram__bytes_read.sum']])
    wr = to_bytes(r[ix['dram__bytes_write.sum']], units[ix['dram__bytes_write.sum']])
    launches.append((name, rd + wr))
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'ncu_traffic.json')
if dense_only:
    ni = [b for n, b in launches if 'dense_ni_tc_kernel' in n]
    if not ni:
        sys.exit('no dense_ni_tc_kernel launch in the report')
    with open(path) as f:
        res = json.load(f)
    commit = subprocess.run(['git', 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip()
    res['kernels']['dense_ni_tc'] = {'dram_bytes': int(ni[-1]), 'launches': 1, 'commit': commit, 'report': os.path.basename(rep)}
    with open(path, 'w') as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res['kernels']['dense_ni_tc']))
    sys.exit(0)
fwd = [b for n, b in launches if '<(int)16, (bool)0, (int)3>' in n or '<16, 0, 3>' in n]
# transpose-backward: unweighted plain flush <16, 0, 0> (the D^-1/2 factor is applied where dA2 is produced); <16, 1, 0> before that
bwd = [b for n, b in launches if any(t in n for t in ('<(int)16, (bool)0, (int)0>', '<16, 0, 0>', '<(int)16, (bool)1, (int)0>', '<16, 1, 0>'))]
if len(fwd) < 3 or not bwd:
    sys.exit(f'need >= 3 forward aggregation launches and one backward launch, got {len(fwd)} / {len(bwd)}: {[n for n, _ in launches]}')
commit = subprocess.run(['git', 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip()
res = {'commit': commit, 'report': os.path.basename(rep),
       'how': 'ncu --set full --clock-control none (cold cache, serialised): dram__bytes_read.sum + dram__bytes_write.sum',
       'kernels': {'spmm_l1_f128': {'dram_bytes': int(fwd[0] + fwd[1]), 'launches': 2},
                   'spmm_l2_f64': {'dram_bytes': int(fwd[2]), 'launches': 1},
                   'spmm_bwd_f64': {'dram_bytes': int(bwd[-1]), 'launches': 1}}}
try:
    with open(path) as f:
        old = json.load(f)
    if 'dense_ni_tc' in old.get('kernels', {}):
        res['kernels']['dense_ni_tc'] = old['kernels']['dense_ni_tc']
except (OSError, ValueError):
    pass
with open(path, 'w') as f:
    json.dump(res, f, indent=1)
print(json.dumps(res))
