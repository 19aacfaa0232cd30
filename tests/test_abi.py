"""CPU: the C-ABI library builds for sm_100a, loads, and exports exactly what include/pvsg.h
declares (no compute calls here -- there is no GPU in the build container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from openpvsg_b200 import lib as l
    if not os.path.exists(l.LIB_PATH):
        l.build()
    return l


def _declared():
    src = open(os.path.join(ROOT, 'include', 'pvsg.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(pvsg_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 20
    dll = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(dll, n), f'{n} declared in include/pvsg.h but not exported'
    # and the Python prototypes cover the same set
    assert sorted(lib.SIGNATURES) == names


def test_version_and_errors(lib):
    dll = lib.load()
    assert dll.pvsg_version() == 100
    assert dll.pvsg_error_string(0) == b'ok'
    assert b'invalid' in dll.pvsg_error_string(-1)
    assert b'unknown' in dll.pvsg_error_string(-99)


def test_argument_checks_launch_nothing(lib):
    """Null / non-positive arguments are rejected before any CUDA call."""
    dll = lib.load()
    assert dll.pvsg_linear(None, None, None, None, None, None, 1, 1, 1, 1, 1, 1, 1, 0, 1, 0, 0, 0, None) == -1
    assert dll.pvsg_layernorm(None, None, None, None, 1, 256, 1e-5, None) == -1
    assert dll.pvsg_top_pairs(None, 4, 4, None, None, None) == -1
    with pytest.raises(lib.PvsgError):
        lib.check(-2, 'x')


def test_sass_is_sm100_only(lib):
    out = os.popen(f'cuobjdump --list-elf {lib.LIB_PATH} 2>/dev/null').read()
    if not out:
        pytest.skip('cuobjdump not available')
    archs = set(re.findall(r'sm_(\d+a?)', out))
    assert archs == {'100a'}, archs


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under openpvsg_b200/ may import it."""
    pkg = os.path.join(ROOT, 'openpvsg_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f
