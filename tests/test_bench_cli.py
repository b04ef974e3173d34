"""bench.py's CPU arm prints one well-formed JSON line (tiny sizes; the GPU arm is exercised on
the B200 box by the driver)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '0', '--ref-procs', '1', '--pool', '2000', '--cpu-sample-rows', '500',
                          '--sources', '2', '--n0', '8', '--pair-loops', 'vectorised'],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better',
                'scaling', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in line, key
    assert line['impl'] == 'reference' and line['value'] > 0 and line['unit'] == 'iters/s'
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['cpu_baseline']['kind'] == 'port'
    assert 'workload' in line['config']
