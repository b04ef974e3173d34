"""The C-ABI library loads and exports every symbol include/tlbo_b200.h declares; the host
layer fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'tlbo_b200.h')


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    names = re.findall(r'\b(tlbo_[a-z0-9_]+)\s*\(', src)
    return sorted(set(names))


def test_header_symbols_are_exported_and_bound():
    import __graft_entry__ as ge
    ge.build()
    from tlbo_b200 import _cabi
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    declared = _declared_functions()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), 'library does not export %s' % name
        assert name in _cabi.SYMBOLS, 'no ctypes signature for %s' % name
    for name in _cabi.SYMBOLS:
        assert name in declared, '%s is bound but not declared in the header' % name
    assert _cabi.lib.tlbo_version() >= 100
    assert isinstance(_cabi.lib.tlbo_last_error(), bytes)


def test_library_is_sm100a_cuda_code():
    from tlbo_b200 import _cabi
    import subprocess
    out = subprocess.run(['cuobjdump', '-lelf', _cabi.LIB_PATH], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         text=True)
    if out.returncode != 0:
        pytest.skip('cuobjdump unavailable')
    assert 'sm_100a' in out.stdout


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA device present')
    from tlbo_b200 import ops
    from tlbo_b200.config_space import TableSpace
    from tlbo_b200.model.model_builder import build_model
    with pytest.raises(RuntimeError):
        ops.SurrogatePack(1, 8, 3)
    m = build_model('gp', TableSpace([0, 0, 0]), np.random.RandomState(1))
    with pytest.raises(RuntimeError):
        m.train(np.random.rand(6, 3), np.random.rand(6))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'efficient-tlbo_b200', 'tlbo_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith('.py'):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), fn
