"""The C-ABI library builds, loads and exports every symbol include/r2n2_b200.h declares, and
the ctypes prototypes mirror the header (no compute calls: runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def built():
    import __graft_entry__ as g
    g.build()
    from r2n2_b200 import _lib
    return _lib


def _header_functions():
    src = open(os.path.join(ROOT, 'include', 'r2n2_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    out = {}
    for m in re.finditer(r'\b(int|const char\*)\s+(r2n2_\w+)\s*\(([^;]*?)\)\s*;', src, flags=re.S):
        args = [a.strip() for a in m.group(3).split(',')]
        if args == ['void']:
            args = []
        out[m.group(2)] = args
    return out


def test_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built.LIB_PATH)
    fns = _header_functions()
    assert len(fns) >= 25
    for name in fns:
        assert hasattr(lib, name), 'missing export ' + name


def test_ctypes_prototypes_mirror_header(built):
    fns = _header_functions()
    kinds = {ctypes.c_void_p: 'p', ctypes.c_int: 'i', ctypes.c_longlong: 'l', ctypes.c_float: 'f'}
    for name, argtypes in built.PROTOTYPES.items():
        assert name in fns, name
        decl = fns[name]
        assert len(decl) == len(argtypes), (name, len(decl), len(argtypes))
        for d, a in zip(decl, argtypes):
            if '*' in d:
                want = 'p'
            elif d.startswith('long long'):
                want = 'l'
            elif d.startswith('float'):
                want = 'f'
            else:
                want = 'i'
            assert kinds[a] == want, (name, d)
    declared = set(fns) - {'r2n2_version', 'r2n2_last_error'}
    assert declared == set(built.PROTOTYPES), declared ^ set(built.PROTOTYPES)


def test_version_and_error_string(built):
    assert 'sm_100a' in built.version()


def test_product_has_no_cpu_fallback():
    """ops refuse CPU tensors; the engine refuses to build without CUDA."""
    import torch
    from r2n2_b200 import ops
    from r2n2_b200.engine import Engine
    with pytest.raises(RuntimeError):
        ops.add(torch.zeros(4), torch.zeros(4), torch.zeros(4))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            Engine('GRUNet', 1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, '3d-r2n2_b200')
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dp, f)).read()
                assert 'import oracle' not in txt and 'from oracle' not in txt, f
