"""Build ``lib/libtlbo_b200.so`` in-tree with nvcc for sm_100a.

``python -m tlbo_b200.build [--force]``.  The library is plain CUDA runtime code
behind the C ABI of ``include/tlbo_b200.h``; nvcc cross-compiles it without a GPU.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(_HERE)
CSRC = os.path.join(PKG, 'csrc')
LIB_DIR = os.path.join(PKG, 'lib')
OBJ_DIR = os.path.join(PKG, 'build')
LIB = os.path.join(LIB_DIR, 'libtlbo_b200.so')
SOURCES = ['c_api.cu', 'gp_fit.cu', 'gp_mcmc.cu', 'gp_predict.cu', 'host_search.cu', 'weights.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-ffp-contract=off', '-diag-suppress', '550']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return 'nvcc'


def _deps_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(PKG), 'include')):
        for fn in os.listdir(root):
            m = max(m, os.path.getmtime(os.path.join(root, fn)))
    return m


def build(force=False, verbose=False):
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    nvcc = _nvcc()

    def one(src):
        obj = os.path.join(OBJ_DIR, src.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s' % (src, r.stdout))
        if verbose:
            sys.stderr.write(r.stdout)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(one, SOURCES))
    cmd = [nvcc, '-shared', '-o', LIB] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n%s' % r.stdout)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
