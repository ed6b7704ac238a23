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
    return sorted(set(names))


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
