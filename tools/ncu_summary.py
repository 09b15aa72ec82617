#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + per-opcode stall samples) into text for profiles/."""
import csv, io, subprocess, sys, re
from collections import Counter

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ''
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.max', 'smsp__cycles_active.avg']
for r in data:
    print('==', r[ci['Kernel Name']][:110])
    for w in want:
        if w in ci:
            print('   %-90s %s %s' % (w, r[ci[w]], units[ci[w]]))
    st = [(h, r[i]) for h, i in ci.items() if re.match(r'smsp__average_warps_issue_stalled_.*_per_issue_active.ratio', h)]
    st = sorted(st, key=lambda kv: -float(kv[1] or 0))[:7]
    print('   stalls/issue:', ', '.join('%s=%.2f' % (h.split('stalled_')[1].split('_per_issue')[0], float(v)) for h, v in st))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
acc = {}
kernel, cur = None, None
for r in rows:
    if r and r[0] == 'Kernel Name':
        kernel, cur = r[1], None
        continue
    if r and r[0] == 'Address':
        cur = {h: i for i, h in enumerate(r)}
        continue
    if cur is None or not r:
        continue
    try:
        n = float(r[cur['# Samples']])
    except Exception:
        continue
    srcl = r[cur['Source']].strip()
    toks = srcl.split()
    op = toks[1] if toks and toks[0].startswith('@') and len(toks) > 1 else (toks[0] if toks else '?')
    stl = {h: r[i] for h, i in cur.items() if h.startswith('stall_') and 'Not' not in h}
    acc.setdefault(kernel, []).append((n, op.split('.')[0], srcl, stl))
for kernel, items in acc.items():
    tot = sum(n for n, _, _, _ in items) or 1
    print('== samples by opcode:', kernel[:100], 'total', int(tot))
    c = Counter()
    for n, op, _, _ in items:
        c[op] += n
    print('   ' + ', '.join('%s %.1f%%' % (op, 100 * n / tot) for op, n in c.most_common(12)))
    for n, op, s, stl in sorted(items, key=lambda x: -x[0])[:12]:
        top = sorted(((k, float(v or 0)) for k, v in stl.items()), key=lambda kv: -kv[1])[:2]
        print('   %5.2f%%  %-60s %s' % (100 * n / tot, s[:60], top))
