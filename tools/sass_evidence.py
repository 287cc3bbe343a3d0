"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (tcgen05 MMA / TMEM loads / bulk prefetch /
cp.async / vector float reductions) in the shipped library + the first lines of each tcgen05 instruction.
usage: python tools/sass_evidence.py > profiles/sass_gemm_tc.txt"""
import re, subprocess, collections, os
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gnndelete_b200', 'libgnndelete_b200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
pat = re.compile(r'\b(UTCHMMA|UTCQMMA|UTCIMMA|LDTM|STTM|UTCBAR|UBLKPF|UBLKCP|UTMALDG|UTMASTG|LDGSTS|HMMA|FFMA2|FADD2|RED|ATOMG)\b[\w.]*')
fn, counts, samples = None, collections.defaultdict(collections.Counter), collections.defaultdict(dict)
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        fn = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        fn = re.sub(r'\(.*', '', fn)
        continue
    m = pat.search(line)
    if m and fn:
        op = m.group(1)
        counts[fn][op] += 1
        if op.startswith(('UTC', 'LDTM', 'UBLK')) and op not in samples[fn]:
            samples[fn][op] = re.sub(r'\s+', ' ', re.sub(r'/\*[0-9a-f]+\*/', '', line)).strip()
print('cuobjdump -sass gnndelete_b200/libgnndelete_b200.so, counts of Blackwell-path mnemonics per kernel (sm_100a)\n')
for fn in sorted(counts):
    c = counts[fn]
    if not any(k.startswith(('UTC', 'LDTM', 'UBLK', 'LDGSTS', 'RED')) for k in c):
        continue
    print(fn)
    print('    ' + ', '.join(f'{k} x{v}' for k, v in sorted(c.items())))
    for op, l in samples[fn].items():
        print('      e.g. ' + l[:150])
