import sys, os, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/efficient-tlbo_b200')
import bench
from tlbo_b200 import ops
peak = ops.fp64_peak_probe()
r = bench.predict_sweep(peak, int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**6)
for x in r['rows']:
    print(x['d'], x['n'], x['N'], round(x['ms'], 3), '%.3e' % x['candidates_per_s'], round(x['frac_of_dmma_peak'], 3))
