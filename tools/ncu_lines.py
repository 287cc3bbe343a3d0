"""Per-source-line stall samples of one kernel in an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [kernel-id like :::0] [top N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; kid = sys.argv[2] if len(sys.argv) > 2 else '-'; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'] + ([] if kid == '-' else ['--kernel-id', kid]),
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if len(r) > 3 and r[0] == 'Line No')
hdr = rows[hi]
si, ie = hdr.index('# Samples'), hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_')]
lines, cur = {}, None
for r in rows[hi + 1:]:
    if len(r) <= si: continue
    if r[0].isdigit():
        cur = int(r[0]); lines.setdefault(cur, dict(src=r[1], s=0, n=0, st={})); continue
    if r[2].startswith('0x') and cur is not None:
        try: s = int(r[si]); n = int(r[ie])
        except ValueError: continue
        d = lines[cur]; d['s'] += s; d['n'] += n
        for c in stall_cols:
            try: v = int(r[c])
            except ValueError: v = 0
            if v: d['st'][hdr[c][6:]] = d['st'].get(hdr[c][6:], 0) + v
tot = sum(d['s'] for d in lines.values())
print('total samples', tot)
for l, d in sorted(lines.items(), key=lambda kv: -kv[1]['s'])[:top]:
    st = ' '.join(f'{k}:{v}' for k, v in sorted(d['st'].items(), key=lambda kv: -kv[1])[:3])
    print(f"{d['s']:6d} {100 * d['s'] / max(tot, 1):5.1f}%  inst {d['n']:>9d}  L{l}: {d['src'].strip()[:90]}   [{st}]")
