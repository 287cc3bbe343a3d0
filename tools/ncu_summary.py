"""Markdown table of the key counters of every kernel in an .ncu-rep (ncu --set full).
usage: python tools/ncu_summary.py report.ncu-rep > profiles/xxx.md"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = [('Kernel Name', 'kernel'), ('launch__grid_size', 'grid'), ('launch__registers_per_thread', 'regs'),
        ('gpu__time_duration.sum', 'duration'), ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
        ('lts__t_sector_hit_rate.pct', 'L2 hit %'), ('l1tex__t_sector_hit_rate.pct', 'L1 hit %'),
        ('l1tex__m_xbar2l1tex_read_bytes.sum.per_second', 'L2->L1'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occ %'),
        ('smsp__inst_executed.sum', 'warp instr'), ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'long_sb/issue'),
        ('sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor pipe %'),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem conflicts'),
        ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smem wavefronts')]
idx = {h: i for i, h in enumerate(hdr)}
cols = [(k, n) for k, n in want if k in idx]
print('| ' + ' | '.join(n for _, n in cols) + ' |')
print('|' + '---|' * len(cols))
def fmt(k, v):
    u = units[idx[k]]
    try:
        f = float(v)
        v = f'{f:.4g}' if abs(f) < 1e6 else f'{f:.4e}'
    except ValueError:
        v = v.replace('void ', '').replace('gd::', '')[:60]
    return (v + ' ' + u).strip()
for r in rows[2:]:
    print('| ' + ' | '.join(fmt(k, r[idx[k]]) for k, _ in cols) + ' |')
