"""The ``pyjacob`` / ``cu_pyjacob`` Python surface of the reference, over the B200 library.

The reference builds two Cython extensions per mechanism (pyjac/pywrap/pyjacob_wrapper.pyx:18-55,
pyjac/pywrap/pyjacob_cuda_wrapper.pyx:13-34; driver pywrap_gen.py:66-128).  Here the same
functions -- same names, argument order, in-place outputs, ``None`` return -- are thin ctypes
calls into the fixed library's reference-named entry points (include/pyjac_b200.h surfaces
2 and 3).  ``generate_wrapper`` mirrors the reference's build call and returns an object that
carries them, bound to one mechanism:

    mod = generate_wrapper('cuda', build_path)      # build_path written by create_jacobian
    mod.py_eval_jacobian(t, P, y, jac)               # one state, runs on the GPU
    padded = mod.py_cuinit(num); mod.py_cujac(num, padded, pres, y, conc, fwd, rev, pm, sr, dy, jac)

Arrays are 1-D contiguous float64 numpy arrays, modified in place; species / reaction order is
pyJac's internal order (last species moved to the end).  There is no CPU implementation behind
any of these: without a CUDA device ``generate_wrapper`` raises.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import create_jacobian as _cj
from . import lib as _lib


def _buf(a: Optional[np.ndarray], what: str, size: int = None):
    if a is None:
        return None
    if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous):
        raise TypeError('%s must be a C-contiguous float64 numpy array' % what)
    if size is not None and a.size < size:
        raise ValueError('%s has %d elements, %d needed' % (what, a.size, size))
    return a.ctypes.data


class PyJacob:
    """Functions of ``pyjacob`` (scalar API) and ``cu_pyjacob`` (batched host API)."""

    def __init__(self, table_blob: bytes, device: Optional[int] = None):
        self._lib = _lib.load()
        if self._lib.pyjac_device_count() <= 0:
            raise _lib.PyjacError('no CUDA device: pyjac_b200 has no CPU fallback')
        h = ctypes.c_void_p()
        _lib.check(self._lib.pyjac_mech_create(table_blob, len(table_blob),
                                               -1 if device is None else int(device), ctypes.byref(h)))
        self._h = h
        dims = (ctypes.c_int * 4)()
        _lib.check(self._lib.pyjac_mech_dims(self._h, dims))
        self.NSP, self.FWD_RATES, self.REV_RATES, self.PRES_MOD_RATES = (int(v) for v in dims)

    def close(self):
        if getattr(self, '_h', None):
            self._lib.pyjac_mech_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _select(self):
        _lib.check(self._lib.pyjac_set_mechanism(self._h))

    # ---- pyjacob (pyjacob_wrapper.pyx:18-55)
    def py_dydt(self, t, pres, y, dy):
        self._select()
        self._lib.dydt(float(t), float(pres), _buf(y, 'y', self.NSP), _buf(dy, 'dy', self.NSP))

    def py_eval_jacobian(self, t, pres, y, jac):
        self._select()
        self._lib.eval_jacob(float(t), float(pres), _buf(y, 'y', self.NSP), _buf(jac, 'jac', self.NSP ** 2))

    def py_eval_rxn_rates(self, T, pres, C, fwd_rxn_rates, rev_rxn_rates):
        self._select()
        self._lib.eval_rxn_rates(float(T), float(pres), _buf(C, 'C', self.NSP),
                                 _buf(fwd_rxn_rates, 'fwd_rxn_rates', self.FWD_RATES),
                                 _buf(rev_rxn_rates, 'rev_rxn_rates', self.REV_RATES))

    def py_eval_spec_rates(self, fwd_rxn_rates, rev_rxn_rates, pres_mod, sp_rates):
        """The last species' rate lands in ``sp_rates[-1]`` (pyjacob_wrapper.pyx:36-40)."""
        self._select()
        base = _buf(sp_rates, 'sp_rates', self.NSP)
        self._lib.eval_spec_rates(_buf(fwd_rxn_rates, 'fwd_rxn_rates', self.FWD_RATES),
                                  _buf(rev_rxn_rates, 'rev_rxn_rates', self.REV_RATES),
                                  _buf(pres_mod, 'pres_mod', self.PRES_MOD_RATES),
                                  base, base + 8 * (sp_rates.shape[0] - 1))

    def py_get_rxn_pres_mod(self, T, pres, C, pres_mod):
        self._select()
        self._lib.get_rxn_pres_mod(float(T), float(pres), _buf(C, 'C', self.NSP),
                                   _buf(pres_mod, 'pres_mod', self.PRES_MOD_RATES))

    def py_eval_conc(self, T, pres, mass_frac, mw_avg, rho, conc):
        """Writes Y_N into ``mass_frac[-1]`` and C into ``conc``; ``mw_avg`` / ``rho`` are passed
        by value and lost, exactly as in the reference (pyjacob_wrapper.pyx:49-55)."""
        self._select()
        base = _buf(mass_frac, 'mass_frac', self.NSP)
        mw, r = ctypes.c_double(float(mw_avg)), ctypes.c_double(float(rho))
        self._lib.eval_conc(float(T), float(pres), base,
                            ctypes.cast(base + 8 * (mass_frac.shape[0] - 1), ctypes.POINTER(ctypes.c_double)),
                            ctypes.byref(mw), ctypes.byref(r), _buf(conc, 'conc', self.NSP))

    # ---- cu_pyjacob (pyjacob_cuda_wrapper.pyx:13-34)
    def py_cuinit(self, num: int) -> int:
        self._select()
        padded = self._lib.pyjac_cu_init(int(num))
        if padded < 0:
            _lib.check(padded)
        return padded

    def py_cujac(self, num, padded, pres, y, conc, fwd_rates, rev_rates, pres_mod, spec_rates, dy, jac):
        """Host arrays flattened in Fortran order of (num, width) -- variable-major, state-fastest
        (functional_tester/test.py:656-660,732-733)."""
        self._select()
        n = int(num)
        self._lib.pyjac_cu_run(n, int(padded), _buf(pres, 'pres', n), _buf(y, 'y', n * self.NSP),
                               _buf(conc, 'conc', n * self.NSP), _buf(fwd_rates, 'fwd_rates', n * self.FWD_RATES),
                               _buf(rev_rates, 'rev_rates', n * self.REV_RATES),
                               _buf(pres_mod, 'pres_mod', n * self.PRES_MOD_RATES),
                               _buf(spec_rates, 'spec_rates', n * self.NSP), _buf(dy, 'dy', n * self.NSP),
                               _buf(jac, 'jac', n * self.NSP ** 2))

    def py_cuclean(self):
        self._select()
        self._lib.pyjac_cu_cleanup()


def generate_wrapper(lang: str, source_dir: str, out_dir: Optional[str] = None, auto_diff: bool = False,
                     device: Optional[int] = None) -> PyJacob:
    """Signature of pyjac/pywrap/pywrap_gen.py:66.  ``source_dir`` was written by
    :func:`pyjac_b200.create_jacobian.create_jacobian`; nothing is compiled here (the library is
    fixed), the returned object carries the ``py_*`` functions for that mechanism."""
    if lang != 'cuda':
        raise ValueError("pyjac_b200 only targets CUDA (sm_100a); lang=%r" % (lang,))
    if auto_diff:
        raise NotImplementedError('auto_diff wrappers are outside the hot path')
    return PyJacob(_cj.load_tables(source_dir), device)
