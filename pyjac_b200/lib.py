"""ctypes binding of include/pyjac_b200.h -- the only doorway from Python to the kernels.

Stands where the reference's Cython glue does (pyjac/pywrap/pyjacob_wrapper.pyx,
pyjac/pywrap/pyjacob_cuda_wrapper.pyx): typed buffers in, raw pointers out.  Loading fails
loudly if the in-tree CUDA library cannot be found or built; nothing here computes.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_longlong, c_size_t, c_void_p

from . import libgen

_dp = POINTER(c_double)

PYJAC_OK = 0
JAC_STATE_MAJOR = 0
JAC_STATE_FASTEST = 1

# every symbol include/pyjac_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    'pyjac_last_error': (c_char_p, []),
    'pyjac_device_count': (c_int, []),
    'pyjac_mech_create': (c_int, [c_void_p, c_size_t, c_int, POINTER(c_void_p)]),
    'pyjac_mech_destroy': (None, [c_void_p]),
    'pyjac_mech_dims': (c_int, [c_void_p, POINTER(c_int)]),
    'pyjac_mech_tune': (c_int, [c_void_p, c_int]),
    'pyjac_mech_launches': (c_longlong, [c_void_p]),
    'pyjac_mech_kernel_name': (c_int, [c_void_p, c_int, c_char_p, c_size_t]),
    'pyjac_eval_jacob_dev': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_longlong, c_longlong,
                                     c_void_p, c_int, c_longlong, c_void_p]),
    'pyjac_dydt_dev': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_longlong, c_longlong,
                               c_void_p, c_longlong, c_longlong, c_void_p]),
    'pyjac_dydt_conv_dev': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_longlong, c_longlong,
                                    c_void_p, c_longlong, c_longlong, c_void_p]),
    'pyjac_mech_set_conv': (c_int, [c_void_p, c_int]),
    'pyjac_fd_jacob_dev': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_double, c_void_p]),
    'pyjac_rates_dev': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_longlong, c_longlong,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_int, c_longlong, c_void_p]),
    'pyjac_set_mechanism': (c_int, [c_void_p]),
    'pyjac_cu_init': (c_int, [c_int]),
    'pyjac_cu_run': (None, [c_int, c_int] + [c_void_p] * 9),
    'pyjac_cu_cleanup': (None, []),
    'pyjac_eval_jacob_host': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    'pyjac_factored_size': (c_int, [c_void_p, POINTER(c_int), POINTER(c_int)]),
    'pyjac_factored_pattern': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'pyjac_eval_jacob_factored_dev': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_longlong, c_longlong,
                                              c_void_p, c_int, c_longlong, c_void_p]),
    'pyjac_eval_jacob_factored_host': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    'pyjac_jvp_dev': (c_int, [c_void_p, c_int, c_void_p, c_int, c_longlong, c_void_p, c_longlong, c_longlong,
                              c_void_p, c_longlong, c_longlong, c_void_p]),
    'pyjac_newton_solve_dev': (c_int, [c_void_p, c_int, c_void_p, c_int, c_longlong, c_double, c_void_p,
                                       c_void_p, c_longlong, c_longlong, c_void_p, c_longlong, c_longlong,
                                       c_void_p, c_void_p]),
    'pyjac_dydt_host': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    'eval_jacob': (None, [c_double, c_double, c_void_p, c_void_p]),
    'dydt': (None, [c_double, c_double, c_void_p, c_void_p]),
    'eval_conc': (None, [c_double, c_double, c_void_p, _dp, _dp, _dp, c_void_p]),
    'eval_rxn_rates': (None, [c_double, c_double, c_void_p, c_void_p, c_void_p]),
    'get_rxn_pres_mod': (None, [c_double, c_double, c_void_p, c_void_p]),
    'eval_spec_rates': (None, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'eval_h': (None, [c_double, c_void_p]),
    'eval_u': (None, [c_double, c_void_p]),
    'eval_cv': (None, [c_double, c_void_p]),
    'eval_cp': (None, [c_double, c_void_p]),
    'apply_mask': (None, [c_void_p]),
    'apply_reverse_mask': (None, [c_void_p]),
    'pyjac_register_tables': (c_int, [c_void_p, c_size_t]),
    # pyjac/pywrap/pyjacob.cuh:6-10, C++ linkage: int init(int); void run(int, int, const double*,
    # const double*, double* x 7); void cleanup()
    '_Z4initi': (c_int, [c_int]),
    '_Z3runiiPKdS0_PdS1_S1_S1_S1_S1_S1_': (None, [c_int, c_int] + [c_void_p] * 9),
    '_Z7cleanupv': (None, []),
}

_LIB = None


class PyjacError(RuntimeError):
    pass


def load(path: str = None) -> ctypes.CDLL:
    """The C-ABI library, built in-tree on first use (nvcc, sm_100a)."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    lib_path = path or libgen.build_library()
    try:
        lib = ctypes.CDLL(lib_path)
    except OSError as exc:
        raise PyjacError('cannot load the pyjac_b200 CUDA library %s: %s' % (lib_path, exc))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _LIB = lib
    return lib


def use(path: str) -> ctypes.CDLL:
    """Makes another build of the library (same C ABI) the one every later load() returns; the
    development tools under tools/ compare builds this way."""
    global _LIB
    _LIB = None
    lib = load(path)
    _LIB = lib
    return lib


def check(rc: int) -> None:
    if rc != PYJAC_OK:
        msg = load().pyjac_last_error()
        raise PyjacError('pyjac_b200 error %d: %s' % (rc, msg.decode() if msg else '?'))
