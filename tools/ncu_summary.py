#!/usr/bin/env python
"""Summarise an .ncu-rep here (no GPU): per-launch key metrics, and for one launch the stall mix / hottest SASS lines.
   tools/ncu_summary.py rep.ncu-rep [--src <launch index> [--top N]]"""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size']
idx = [hdr.index(w) if w in hdr else None for w in want]
print(' | '.join(w.split('.')[0][-28:] for w in want))
for r in rows[2:]:
    print(' | '.join((r[i][:34] if i is not None else '-') for i in idx))
if '--src' in sys.argv:
    k = int(sys.argv[sys.argv.index('--src') + 1])
    top = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 25
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(k), "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    print(rows[0][1][:120])
    hdr = rows[1]
    i_src, i_s, i_ex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = Counter(); total = 0; data = []; ops = Counter(); opsamp = Counter()
    for r in rows[2:]:
        if len(r) < len(hdr) or not r[i_s].isdigit():
            continue
        n = int(r[i_s]); total += n
        st = {}
        for s_ in stalls:
            v = r[hdr.index(s_)]
            if v not in ('0', ''):
                st[s_[6:]] = int(v); tot[s_[6:]] += int(v)
        ex = int(r[i_ex]) if r[i_ex].isdigit() else 0
        toks = r[i_src].split()
        op = (toks[1] if toks and toks[0].startswith('@') else toks[0]).split('.')[0] if toks else '?'
        ops[op] += ex; opsamp[op] += n
        data.append((n, ex, r[i_src][:84], st))
    print('samples', total, {k: round(100 * v / total, 1) for k, v in tot.most_common(12)})
    te = sum(ops.values())
    print('executed by opcode %:', [(k, round(100 * v / te, 1)) for k, v in ops.most_common(18)])
    print('samples by opcode %:', [(k, round(100 * v / total, 1)) for k, v in opsamp.most_common(14)])
    for n, ex, s_, st in sorted(data, key=lambda x: -x[0])[:top]:
        print(n, ex, s_, st)
