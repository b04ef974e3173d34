#!/usr/bin/env python
"""Aggregate warp-stall samples of one kernel by CUDA source line from an .ncu-rep.
usage: ncu_lines.py REP KERNEL_REGEX [launch_skip] [top]"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else '0'
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda', '--kernel-name',
                      'regex:' + kern, '--launch-skip', skip, '--launch-count', '1'], stdout=subprocess.PIPE,
                     stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = collections.OrderedDict()
cur_file, cur = None, None
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if r[0] == 'Line No':
        hdr = r
        si = hdr.index('# Samples'); ii = hdr.index('Instructions Executed')
        continue
    if hdr is None or len(r) <= si:
        continue
    if r[0] != '':
        cur = (cur_file, int(r[0]), r[1].strip()[:90])
        agg.setdefault(cur, [0, 0])
        # the line row itself carries the aggregated numbers
        try:
            agg[cur][0] += int(r[si]); agg[cur][1] += int(r[ii])
        except ValueError:
            pass
tot = sum(v[0] for v in agg.values())
print('total samples', tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%5.1f%% %8d inst=%10d  %s:%d  %s' % (100.0 * v[0] / max(tot, 1), v[0], v[1], k[0], k[1], k[2]))
