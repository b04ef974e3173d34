#!/usr/bin/env python
"""Summarise ncu outputs into profiles/: launch-list shares and key full-set metrics."""
import csv, subprocess, sys, collections, json

def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split('(')[0][:70]
        v = float(r[vi].replace(',', ''))
        v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
        c, t = agg.get(name, (0, 0.0))
        agg[name] = (c + 1, t + v)
    tot = sum(t for c, t in agg.values())
    out = ['| kernel | launches | total us | share |', '|---|---|---|---|']
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append('| `%s` | %d | %.1f | %.1f%% |' % (k, c, t, 100 * t / tot))
    return '\n'.join(out)

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__inst_executed_pipe_tensor_op_dmma.sum', 'smsp__inst_executed_pipe_fp64_op_dmma.sum']

def full(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {'kernel': r[hdr.index('Kernel Name')][:80]}
        for w in WANT:
            if w in hdr:
                d[w] = r[hdr.index(w)] + ' ' + units[hdr.index(w)]
        dm = [h for h in hdr if 'dmma' in h.lower()]
        for h in dm[:4]:
            d[h] = r[hdr.index(h)]
        res.append(d)
    return res

if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        print(launches(sys.argv[2]))
    else:
        print(json.dumps(full(sys.argv[2]), indent=1))
