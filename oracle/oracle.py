"""TEST INFRASTRUCTURE -- Python handles on the two CPU checkers.

* :class:`Oracle`  -- the table-driven C restatement (oracle/pyjac_oracle.c).
* :class:`RefLib`  -- the reference's own generated C, compiled by oracle/build_ref.py
  into oracle/_ref/<name>/libc_pyjac.so (present only where it was built).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
Batch arrays here are *row-major per state*: y[n, NSP] = [T, Y_0..Y_{NSP-2}],
jac[n, NSP*NSP] column-major inside each state, exactly what the scalar API takes.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np

from pyjac_b200 import blob as _blob
from pyjac_b200.mechanism import Mechanism

from . import ref_tables

HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(HERE, '_build')
_LIB = os.path.join(_BUILD, 'liboracle.so')
_dp = ctypes.POINTER(ctypes.c_double)


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_dp)


def build_oracle(force: bool = False) -> str:
    """gcc the restatement with the reference's FP semantics (-std=c99 => no FMA
    contraction; libgen.py:43)."""
    src = os.path.join(HERE, 'pyjac_oracle.c')
    if (not force and os.path.exists(_LIB)
            and os.path.getmtime(_LIB) >= os.path.getmtime(src)):
        return _LIB
    os.makedirs(_BUILD, exist_ok=True)
    cmd = ['gcc', '-std=c99', '-O2', '-ffp-contract=off', '-fPIC', '-fopenmp', '-shared',
           '-o', _LIB, src, '-lm']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('oracle build failed:\n' + r.stderr)
    return _LIB


class Oracle:
    def __init__(self, mech: Mechanism):
        self.mech = mech
        self.lib = ctypes.CDLL(build_oracle())
        L = self.lib
        L.oracle_load.restype = ctypes.c_void_p
        L.oracle_load.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        L.oracle_free.argtypes = [ctypes.c_void_p]
        for fn in ('oracle_eval_jacob_batch', 'oracle_dydt_batch', 'oracle_dydt_conv_batch'):
            getattr(L, fn).argtypes = [ctypes.c_void_p, ctypes.c_int, _dp, _dp, _dp, ctypes.c_int]
        L.oracle_rates_batch.argtypes = [ctypes.c_void_p, ctypes.c_int] + [_dp] * 7 + [ctypes.c_int]
        data = _blob.pack(ref_tables.build(mech))
        self._h = L.oracle_load(data, len(data))
        if not self._h:
            raise RuntimeError('oracle_load failed')
        self.NSP, self.NR = mech.NSP, mech.FWD_RATES
        self.NREV, self.NPD = mech.REV_RATES, mech.PRES_MOD_RATES

    def __del__(self):
        try:
            self.lib.oracle_free(self._h)
        except Exception:
            pass

    def _prep(self, P, y):
        y = np.ascontiguousarray(y, dtype=np.float64)
        P = np.ascontiguousarray(np.broadcast_to(np.asarray(P, dtype=np.float64), (y.shape[0],)))
        assert y.ndim == 2 and y.shape[1] == self.NSP
        return P, y

    def eval_jacob(self, P, y, nthreads: int = 0) -> np.ndarray:
        P, y = self._prep(P, y)
        jac = np.zeros((y.shape[0], self.NSP * self.NSP))
        self.lib.oracle_eval_jacob_batch(self._h, y.shape[0], _p(P), _p(y), _p(jac), nthreads)
        return jac

    def dydt(self, P, y, nthreads: int = 0) -> np.ndarray:
        P, y = self._prep(P, y)
        dy = np.zeros((y.shape[0], self.NSP))
        self.lib.oracle_dydt_batch(self._h, y.shape[0], _p(P), _p(y), _p(dy), nthreads)
        return dy

    def dydt_conv(self, rho, y, nthreads: int = 0) -> np.ndarray:
        """constant-volume dydt(t, rho, y, dy) (rate_subs.py:2340-2485)"""
        rho, y = self._prep(rho, y)
        dy = np.zeros((y.shape[0], self.NSP))
        self.lib.oracle_dydt_conv_batch(self._h, y.shape[0], _p(rho), _p(y), _p(dy), nthreads)
        return dy

    def rates(self, P, y, nthreads: int = 0):
        """conc, fwd, rev, pres_mod, spec_rates (internal species / reaction order)."""
        P, y = self._prep(P, y)
        n = y.shape[0]
        conc = np.zeros((n, self.NSP))
        fwd = np.zeros((n, self.NR))
        rev = np.zeros((n, max(self.NREV, 1)))
        pm = np.zeros((n, max(self.NPD, 1)))
        sr = np.zeros((n, self.NSP))
        # the C side strides by NREV / NPD, so allocate exactly that when non-zero
        if self.NREV:
            rev = np.zeros((n, self.NREV))
        if self.NPD:
            pm = np.zeros((n, self.NPD))
        self.lib.oracle_rates_batch(self._h, n, _p(P), _p(y), _p(conc), _p(fwd), _p(rev),
                                    _p(pm), _p(sr), nthreads)
        return conc, fwd, rev[:, :self.NREV], pm[:, :self.NPD], sr


class RefLib:
    """ctypes view of oracle/_ref/<name>/libc_pyjac.so (reference generated C)."""

    def __init__(self, name_or_path: str):
        path = name_or_path
        if not os.path.isabs(path):
            path = os.path.join(HERE, '_ref', name_or_path, 'libc_pyjac.so')
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = L = ctypes.CDLL(path)
        self.NSP, self.NR = L.ref_nsp(), L.ref_fwd_rates()
        self.NREV, self.NPD = L.ref_rev_rates(), L.ref_pres_mod_rates()
        L.ref_eval_jacob_batch.argtypes = [ctypes.c_int, _dp, _dp, _dp, ctypes.c_int]
        L.ref_dydt_batch.argtypes = [ctypes.c_int, _dp, _dp, _dp, ctypes.c_int]
        d = ctypes.c_double
        L.eval_conc.argtypes = [d, d, _dp, _dp, _dp, _dp, _dp]
        L.eval_rxn_rates.argtypes = [d, d, _dp, _dp, _dp]
        L.eval_spec_rates.argtypes = [_dp, _dp, _dp, _dp, _dp]
        if self.NPD:
            L.get_rxn_pres_mod.argtypes = [d, d, _dp, _dp]

    @classmethod
    def available(cls, name: str) -> bool:
        return os.path.exists(os.path.join(HERE, '_ref', name, 'libc_pyjac.so'))

    def _prep(self, P, y):
        y = np.ascontiguousarray(y, dtype=np.float64)
        P = np.ascontiguousarray(np.broadcast_to(np.asarray(P, dtype=np.float64), (y.shape[0],)))
        assert y.ndim == 2 and y.shape[1] == self.NSP
        return P, y

    def eval_jacob(self, P, y, nthreads: int = 0, keep: bool = True):
        P, y = self._prep(P, y)
        jac = np.zeros((y.shape[0], self.NSP * self.NSP)) if keep else None
        self.lib.ref_eval_jacob_batch(y.shape[0], _p(P), _p(y), _p(jac), nthreads)
        return jac

    def dydt(self, P, y, nthreads: int = 0):
        P, y = self._prep(P, y)
        dy = np.zeros((y.shape[0], self.NSP))
        self.lib.ref_dydt_batch(y.shape[0], _p(P), _p(y), _p(dy), nthreads)
        return dy

    def rates(self, P, y):
        P, y = self._prep(P, y)
        n = y.shape[0]
        conc = np.zeros((n, self.NSP))
        fwd = np.zeros((n, self.NR))
        rev = np.zeros((n, max(self.NREV, 1)))
        pm = np.zeros((n, max(self.NPD, 1)))
        sr = np.zeros((n, self.NSP))
        yN, mw, rho = (ctypes.c_double(), ctypes.c_double(), ctypes.c_double())
        L = self.lib
        for s in range(n):
            L.eval_conc(y[s, 0], P[s], _p(y[s, 1:].copy()), ctypes.byref(yN), ctypes.byref(mw),
                        ctypes.byref(rho), _p(conc[s]))
            L.eval_rxn_rates(y[s, 0], P[s], _p(conc[s]), _p(fwd[s]), _p(rev[s]))
            if self.NPD:
                L.get_rxn_pres_mod(y[s, 0], P[s], _p(conc[s]), _p(pm[s]))
            L.eval_spec_rates(_p(fwd[s]), _p(rev[s]), _p(pm[s]), _p(sr[s]), _p(sr[s, -1:]))
        return conc, fwd, rev[:, :self.NREV], pm[:, :self.NPD], sr
