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


    # ---- extras of the emitted chem_utils (no reference wrapper binds them; used by the tests)
    def eval_thermo(self, which: str, T: float) -> np.ndarray:
        """``which`` in {'h', 'u', 'cv', 'cp'}: NSP mass-based values (eval_h / eval_u / eval_cv / eval_cp)."""
        self._select()
        out = np.empty(self.NSP)
        getattr(self._lib, 'eval_' + which)(float(T), out.ctypes.data)
        return out


_MODULE = '''"""``%(name)s`` for the mechanism exported to %(src)r -- written by pyjac_b200.pywrap.generate_wrapper.

The reference builds a Cython extension of this name per mechanism (pyjac/pywrap/%(pyx)s);
this module carries the same functions, bound to the fixed sm_100a library over ctypes."""
from pyjac_b200 import create_jacobian as _cj
from pyjac_b200.pywrap import PyJacob as _PyJacob

_mod = _PyJacob(_cj.load_tables(%(src)r), %(device)r)
NSP, FWD_RATES, REV_RATES, PRES_MOD_RATES = _mod.NSP, _mod.FWD_RATES, _mod.REV_RATES, _mod.PRES_MOD_RATES
%(names)s
'''

_PYJACOB = ('py_dydt', 'py_eval_jacobian', 'py_eval_rxn_rates', 'py_eval_spec_rates', 'py_get_rxn_pres_mod', 'py_eval_conc')
_CU_PYJACOB = ('py_cuinit', 'py_cujac', 'py_cuclean')


def generate_wrapper(lang: str, source_dir: str, out_dir: Optional[str] = None, auto_diff: bool = False,
                     device: Optional[int] = None) -> PyJacob:
    """Signature of pyjac/pywrap/pywrap_gen.py:66.  ``source_dir`` was written by
    :func:`pyjac_b200.create_jacobian.create_jacobian`.  The reference compiles an extension module
    (``pyjacob`` for lang 'c', ``cu_pyjacob`` for 'cuda') into ``out_dir``, which its callers then import
    by name (functional_tester/test.py:432,740).  Here nothing needs compiling: a module of the same
    name is *written* into ``out_dir`` (default: the current directory, like the reference) that binds
    the same functions to the fixed library, and the bound object is returned as well.  Both names run
    on the GPU -- 'c' selects the one-state functions of pyjacob_wrapper.pyx, 'cuda' the batched ones of
    pyjacob_cuda_wrapper.pyx; there is no CPU implementation behind either."""
    if lang not in ('c', 'cuda'):
        raise ValueError("lang must be 'c' (pyjacob) or 'cuda' (cu_pyjacob); got %r" % (lang,))
    if auto_diff:
        raise NotImplementedError('auto_diff wrappers are outside the hot path')
    import os
    src = os.path.abspath(source_dir)
    name, pyx, fns = (('pyjacob', 'pyjacob_wrapper.pyx', _PYJACOB) if lang == 'c'
                      else ('cu_pyjacob', 'pyjacob_cuda_wrapper.pyx', _CU_PYJACOB))
    out_dir = os.path.abspath(out_dir or os.getcwd())
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, name + '.py'), 'w') as fh:
        fh.write(_MODULE % {'name': name, 'src': src, 'pyx': pyx, 'device': device,
                            'names': '\n'.join('%s = _mod.%s' % (f, f) for f in fns)})
    return PyJacob(_cj.load_tables(source_dir), device)
