#!/usr/bin/env python
"""Warp-stall reasons of one kernel aggregated per source FILE (and the top lines of one file) from an .ncu-rep.
usage: ncu_stalls.py REP KERNEL_REGEX [file_for_line_detail] [top]"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
detail = sys.argv[3] if len(sys.argv) > 3 else None
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda', '--kernel-name',
                      'regex:' + kern, '--launch-count', '1'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                     text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, cur_file = None, None
byfile = collections.defaultdict(lambda: collections.Counter())
byline = collections.defaultdict(lambda: collections.Counter())
src = {}
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if r[0] == 'Line No':
        hdr = r
        cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
        si = hdr.index('# Samples'); ii = hdr.index('Instructions Executed')
        continue
    if hdr is None or len(r) <= si or r[0] == '':
        continue
    try:
        ns = int(r[si])
    except ValueError:
        continue
    key = (cur_file, int(r[0]))
    src[key] = r[1].strip()[:80]
    byfile[cur_file]['samples'] += ns
    byline[key]['samples'] += ns
    try:
        byfile[cur_file]['inst'] += int(r[ii]); byline[key]['inst'] += int(r[ii])
    except ValueError:
        pass
    for i, h in cols:
        try:
            v = int(r[i])
        except ValueError:
            v = 0
        byfile[cur_file][h] += v
        byline[key][h] += v
tot = sum(c['samples'] for c in byfile.values())
print('total samples', tot)


def fmt(c):
    parts = sorted(((k, v) for k, v in c.items() if k.startswith('stall_') and v), key=lambda kv: -kv[1])[:6]
    return ' '.join('%s=%d' % (k[6:], v) for k, v in parts)


for f, c in sorted(byfile.items(), key=lambda kv: -kv[1]['samples']):
    print('%5.1f%% %7d inst=%9d %-24s %s' % (100.0 * c['samples'] / max(tot, 1), c['samples'], c['inst'], f, fmt(c)))
if detail:
    print('--- lines of', detail)
    for k, c in sorted(((k, c) for k, c in byline.items() if k[0] == detail), key=lambda kv: -kv[1]['samples'])[:top]:
        print('%6d inst=%8d %s:%d  %-60s %s' % (c['samples'], c['inst'], k[0], k[1], src[k][:60], fmt(c)))
