"""The C-ABI library builds for sm_100a, loads without a GPU, and exports every symbol that
include/pyjac_b200.h declares.  No compute call is made here."""
import ctypes
import os
import re

import pytest

from pyjac_b200 import lib, libgen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, 'include', 'pyjac_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    names = re.findall(r'^\s*(?:const\s+)?(?:int|void|long long|char\s*\*|const char\s*\*)\s*\*?\s*(\w+)\s*\(',
                       txt, flags=re.M)
    # the three functions of pyjacob.cuh are declared with C++ linkage: their symbols are mangled
    cxx = {'run': '_Z3runiiPKdS0_PdS1_S1_S1_S1_S1_S1_', 'init': '_Z4initi', 'cleanup': '_Z7cleanupv'}
    return sorted(set(cxx.get(n, n) for n in names))


def test_header_and_binding_agree():
    assert _declared() == sorted(lib.SIGNATURES)


def test_library_builds_and_exports_all_symbols():
    path = libgen.build_library()
    assert os.path.exists(path)
    cdll = ctypes.CDLL(path)
    for name in _declared():
        assert hasattr(cdll, name), name
    lib.load()


def test_no_cpu_fallback_without_device():
    L = lib.load()
    if L.pyjac_device_count() > 0:
        pytest.skip('a CUDA device is present')
    h = ctypes.c_void_p()
    from pyjac_b200 import blob, tables
    from pyjac_b200.mechanism import Mechanism
    mech = Mechanism.from_chemkin(os.path.join(ROOT, 'tests', 'golden', 'h2o2_n2.inp'))
    data = blob.pack(tables.build(mech))
    rc = L.pyjac_mech_create(data, len(data), 0, ctypes.byref(h))
    assert rc == -2 and b'no CPU fallback' in L.pyjac_last_error()


def test_bad_blob_rejected():
    L = lib.load()
    h = ctypes.c_void_p()
    assert L.pyjac_mech_create(b'x' * 64, 64, 0, ctypes.byref(h)) == -1


def test_stale_or_truncated_tables_rejected():
    """A blob written for another version of the tables, or with a table shorter than the dimensions
    it is indexed with, is refused before anything touches the device."""
    import numpy as np
    from pyjac_b200 import blob, tables
    from pyjac_b200.mechanism import Mechanism
    L = lib.load()
    mech = Mechanism.from_chemkin(os.path.join(ROOT, 'tests', 'golden', 'h2o2_n2.inp'))
    T = tables.build(mech)
    h = ctypes.c_void_p()
    for key, edit in (('meta', lambda a: a + np.array([1, 0, 0, 0], dtype=np.int32)),
                      ('meta', lambda a: a + np.array([0, 1, 0, 0], dtype=np.int32)),
                      ('p5_rx', lambda a: a[:-16]), ('sp_nasa', lambda a: a[:-1]), ('sp_fwd_map', lambda a: a[:-1])):
        bad = dict(T)
        bad[key] = np.ascontiguousarray(edit(T[key]))
        data = blob.pack(bad)
        assert L.pyjac_mech_create(data, len(data), 0, ctypes.byref(h)) == -1, key
        assert b'table blob' in L.pyjac_last_error()
    bad = {k: v for k, v in T.items() if k != 'meta'}
    data = blob.pack(bad)
    assert L.pyjac_mech_create(data, len(data), 0, ctypes.byref(h)) == -1
