"""Host-side handle on one mechanism loaded into the CUDA library.

``Evaluator`` owns a ``pyjac_mech`` (include/pyjac_b200.h) and exposes the batched
device-pointer calls on torch CUDA tensors (torch is used for device memory and streams
only) and the streamed host-pointer calls on numpy arrays.  Layout vocabulary:

* ``'rows'``          one row per state: ``y[n, NSP] = [T, Y_0..Y_{NSP-2}]`` -- what the
                      reference's scalar API takes (docs/faqs.rst:82-87)
* ``'state_fastest'`` variable-major ``y[NSP, ld]`` -- the reference GPU layout
                      (mech_auxiliary.py:418-420; docs/faqs.rst:163-172)

Jacobians: ``'rows'`` gives ``jac[n, NSP*NSP]``, each row one column-major NSP x NSP matrix
(``jac[s, i + NSP*j]``); ``'state_fastest'`` gives ``jac[NSP*NSP, ld]``.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import blob as _blob
from . import lib as _lib
from . import tables as _tables
from .mechanism import Mechanism


def _ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


class Evaluator:
    def __init__(self, mech: Mechanism, device: Optional[int] = None, gs: int = 0, threads: int = 0,
                 ws_global: Optional[bool] = None, streams: Optional[bool] = None):
        """gs / threads: states per thread block and block size of the Jacobian kernel's plan
        (pyjac_b200/plan.py); 0 = automatic.  ws_global: working set of a block in global memory
        instead of shared memory (None = only for mechanisms too large for shared memory).
        streams: True = eval_jacob on the record streams of k_jac6 (pyjac_b200/plan6.py) instead of the
        schedule tables of k_eval (the default, measured faster)."""
        import torch
        self._torch = torch
        self.mech = mech
        self.lib = _lib.load()
        if self.lib.pyjac_device_count() <= 0:
            raise _lib.PyjacError('no CUDA device: pyjac_b200 has no CPU fallback')
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.tables = _tables.build(mech, gs=gs, threads=threads, ws_global=ws_global, streams=streams)
        self.uses_streams = 'p6_str' in self.tables
        self.plan_gs, self.plan_threads = (int(v) for v in self.tables['p5_cfg'][:2])
        data = _blob.pack(self.tables)
        h = ctypes.c_void_p()
        _lib.check(self.lib.pyjac_mech_create(data, len(data), self.device, ctypes.byref(h)))
        self._h = h
        dims = (ctypes.c_int * 4)()
        _lib.check(self.lib.pyjac_mech_dims(self._h, dims))
        self.NSP, self.NR, self.NREV, self.NPD = (int(v) for v in dims)

    def close(self):
        if getattr(self, '_h', None):
            self.lib.pyjac_mech_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def tune(self, blocks_per_sm: int = 0):
        """Cap on resident blocks per SM (0 = automatic); states per block / block size are the
        ``gs`` / ``threads`` arguments of the constructor (they shape the plan tables)."""
        _lib.check(self.lib.pyjac_mech_tune(self._h, blocks_per_sm))

    @property
    def launches(self) -> int:
        return int(self.lib.pyjac_mech_launches(self._h))

    def kernel_name(self, mode: int = 0) -> str:
        """Demangled symbol of the kernel behind eval_jacob (mode 0), dydt (1), the rate routines (2) or the
        factored Jacobian (3)."""
        import subprocess
        buf = ctypes.create_string_buffer(512)
        _lib.check(self.lib.pyjac_mech_kernel_name(self._h, mode, buf, len(buf)))
        name = buf.value.decode()
        try:
            out = subprocess.run(['c++filt', name], capture_output=True, text=True, timeout=10).stdout.strip()
            return out or name
        except Exception:
            return name

    def make_current(self):
        """Select this mechanism for the reference-named entry points (pyjacob / cu_pyjacob)."""
        _lib.check(self.lib.pyjac_set_mechanism(self._h))

    def _stream(self, stream):
        s = self._torch.cuda.current_stream(self.device) if stream is None else stream
        return ctypes.c_void_p(s.cuda_stream)

    def _strides(self, y, layout: str):
        if layout == 'rows':
            assert y.dim() == 2 and y.shape[1] == self.NSP and y.is_contiguous()
            return y.shape[0], self.NSP, 1
        if layout == 'state_fastest':
            assert y.dim() == 2 and y.shape[0] == self.NSP and y.is_contiguous()
            return y.shape[1], 1, y.shape[1]
        raise ValueError(layout)

    def _check_dev(self, *ts):
        for t in ts:
            if t is not None:
                assert t.is_cuda and t.dtype == self._torch.float64 and t.device.index == self.device

    # ------------------------------------------------------------------ device batch API
    def eval_jacob(self, P, y, out=None, y_layout: str = 'rows', jac_layout: str = 'rows',
                   stream=None):
        torch = self._torch
        self._check_dev(P, y, out)
        n, ss, sv = self._strides(y, y_layout)
        nn = self.NSP * self.NSP
        if jac_layout == 'rows':
            if out is None:
                out = torch.empty((n, nn), dtype=torch.float64, device=y.device)
            assert out.shape == (n, nn) and out.is_contiguous()
            lay, ld = _lib.JAC_STATE_MAJOR, 0
        else:
            if out is None:
                out = torch.empty((nn, n), dtype=torch.float64, device=y.device)
            assert out.shape[0] == nn and out.shape[1] >= n and out.is_contiguous()
            lay, ld = _lib.JAC_STATE_FASTEST, out.shape[1]
        _lib.check(self.lib.pyjac_eval_jacob_dev(self._h, n, _ptr(P), _ptr(y), ss, sv, _ptr(out),
                                                 lay, ld, self._stream(stream)))
        return out

    def dydt(self, P, y, out=None, y_layout: str = 'rows', stream=None, conv: bool = False):
        """conv: constant volume -- ``P`` then holds one density per state (rate_subs.py:2340-2485)."""
        torch = self._torch
        self._check_dev(P, y, out)
        n, ss, sv = self._strides(y, y_layout)
        if out is None:
            out = torch.empty_like(y)
        assert out.shape == y.shape and out.is_contiguous()
        fn = self.lib.pyjac_dydt_conv_dev if conv else self.lib.pyjac_dydt_dev
        _lib.check(fn(self._h, n, _ptr(P), _ptr(y), ss, sv, _ptr(out), ss, sv, self._stream(stream)))
        return out

    def fd_jacob(self, P, y, order: int = 6, r_cap: float = 0.0, out=None, stream=None):
        """Finite-difference Jacobian of dydt (state-fastest ``y[NSP, n]`` -> ``jac[NSP*NSP, n]``):
        the on-device self-check of :meth:`eval_jacob` (performance_tester/fd_jacob.cu)."""
        torch = self._torch
        self._check_dev(P, y, out)
        n, _, _ = self._strides(y, 'state_fastest')
        if out is None:
            out = torch.empty((self.NSP * self.NSP, n), dtype=torch.float64, device=y.device)
        assert out.shape == (self.NSP * self.NSP, n) and out.is_contiguous()
        _lib.check(self.lib.pyjac_fd_jacob_dev(self._h, n, _ptr(P), _ptr(y), _ptr(out), int(order),
                                               float(r_cap), self._stream(stream)))
        return out

    def self_check(self, P, y, order: int = 6, r_cap: float = 0.0):
        """max over states and columns of |analytical - finite difference| / scale, scale = the
        column maximum, but at least 1e-6 of the state's largest entry (finite differences cannot
        resolve a column below their own round-off)."""
        torch = self._torch
        jac = self.eval_jacob(P, y, y_layout='state_fastest', jac_layout='state_fastest')
        fd = self.fd_jacob(P, y, order, r_cap)
        nsp, n = self.NSP, y.shape[1]
        a, f = jac.view(nsp, nsp, n), fd.view(nsp, nsp, n)          # [column j][row i][state]
        scale = a.abs().amax(dim=1, keepdim=True)
        scale = torch.maximum(scale, 1e-6 * scale.amax(dim=0, keepdim=True)).clamp_min(1e-300)
        return float(((a - f).abs() / scale).max())

    def rates(self, P, y, y_layout: str = 'rows', want_dy: bool = False, stream=None):
        """conc, fwd, rev, pres_mod, spec_rates[, dy] in the layout of ``y``."""
        torch = self._torch
        self._check_dev(P, y)
        n, ss, sv = self._strides(y, y_layout)
        widths = [self.NSP, self.NR, self.NREV, self.NPD, self.NSP] + ([self.NSP] if want_dy else [])
        sf = y_layout == 'state_fastest'
        outs = [torch.zeros((w, n) if sf else (n, w), dtype=torch.float64, device=y.device)
                for w in widths]
        ptrs = [_ptr(o) if o.numel() else 0 for o in outs] + ([] if want_dy else [0])
        _lib.check(self.lib.pyjac_rates_dev(self._h, n, _ptr(P), _ptr(y), ss, sv, *ptrs,
                                            1 if sf else 0, n, self._stream(stream)))
        return tuple(outs)

    # ------------------------------------------------------------------ factored Jacobian and its consumers
    @property
    def factored_size(self):
        """(NF, NNZ): doubles per state of the factored record, entries of its sparse block."""
        import ctypes
        nf, nnz = ctypes.c_int(), ctypes.c_int()
        _lib.check(self.lib.pyjac_factored_size(self._h, ctypes.byref(nf), ctypes.byref(nnz)))
        return nf.value, nnz.value

    def factored_pattern(self):
        """rows[NNZ], cols[NNZ] of the sparse block (indices into the NSP x NSP Jacobian) and the column
        factors ca[NSP], cb[NSP] of the rank-2 part (include/pyjac_b200.h)."""
        nf, nnz = self.factored_size
        rows, cols = np.zeros(max(nnz, 1), dtype=np.int32), np.zeros(max(nnz, 1), dtype=np.int32)
        ca, cb = np.zeros(self.NSP), np.zeros(self.NSP)
        _lib.check(self.lib.pyjac_factored_pattern(self._h, rows.ctypes.data, cols.ctypes.data,
                                                   ca.ctypes.data, cb.ctypes.data))
        return rows[:nnz], cols[:nnz], ca, cb

    def _fac_layout(self, fac, n, layout):
        nf = self.factored_size[0]
        if layout == 'rows':
            assert fac.shape == (n, nf) and fac.is_contiguous()
            return _lib.JAC_STATE_MAJOR, 0
        assert fac.shape[0] == nf and fac.shape[1] >= n and fac.is_contiguous()
        return _lib.JAC_STATE_FASTEST, fac.shape[1]

    def eval_jacob_factored(self, P, y, out=None, y_layout: str = 'rows', fac_layout: str = 'rows', stream=None):
        """The Jacobian as the record it is expanded from: NF doubles per state instead of NSP^2."""
        torch = self._torch
        self._check_dev(P, y, out)
        n, ss, sv = self._strides(y, y_layout)
        nf = self.factored_size[0]
        if out is None:
            out = torch.empty((n, nf) if fac_layout == 'rows' else (nf, n), dtype=torch.float64, device=y.device)
        lay, ld = self._fac_layout(out, n, fac_layout)
        _lib.check(self.lib.pyjac_eval_jacob_factored_dev(self._h, n, _ptr(P), _ptr(y), ss, sv, _ptr(out),
                                                          lay, ld, self._stream(stream)))
        return out

    def expand_factored(self, fac: np.ndarray) -> np.ndarray:
        """Host-side expansion of records (n, NF) to column-major dense Jacobians (n, NSP*NSP): what a caller of
        the reference API would be handed; the parity tests compare this with the oracle."""
        from . import factored
        return factored.expand(fac, self.NSP, *self.factored_pattern())

    def jvp(self, fac, v, out=None, fac_layout: str = 'rows', v_layout: str = 'rows', stream=None):
        """J v per state from the records of :meth:`eval_jacob_factored` (J is never formed)."""
        torch = self._torch
        self._check_dev(fac, v, out)
        n, ss, sv = self._strides(v, v_layout)
        lay, ld = self._fac_layout(fac, n, fac_layout)
        if out is None:
            out = torch.empty_like(v)
        assert out.shape == v.shape and out.is_contiguous()
        _lib.check(self.lib.pyjac_jvp_dev(self._h, n, _ptr(fac), lay, ld, _ptr(v), ss, sv, _ptr(out), ss, sv,
                                          self._stream(stream)))
        return out

    def newton_solve(self, fac, gamma, rhs, out=None, fac_layout: str = 'rows', v_layout: str = 'rows', stream=None):
        """x = (I - gamma J)^-1 rhs per state (the linear solve of an implicit integrator step);
        ``gamma``: a float or a device tensor with one value per state.  Returns (x, info)."""
        torch = self._torch
        self._check_dev(fac, rhs, out)
        n, ss, sv = self._strides(rhs, v_layout)
        lay, ld = self._fac_layout(fac, n, fac_layout)
        if out is None:
            out = torch.empty_like(rhs)
        assert out.shape == rhs.shape and out.is_contiguous()
        info = torch.zeros(n, dtype=torch.int32, device=rhs.device)
        if isinstance(gamma, (int, float)):
            g, gp = float(gamma), None
        else:
            self._check_dev(gamma)
            assert gamma.shape == (n,) and gamma.dtype == torch.float64 and gamma.is_contiguous()
            g, gp = 0.0, _ptr(gamma)
        _lib.check(self.lib.pyjac_newton_solve_dev(self._h, n, _ptr(fac), lay, ld, g, gp, _ptr(rhs), ss, sv,
                                                   _ptr(out), ss, sv, _ptr(info), self._stream(stream)))
        return out, info

    def eval_jacob_factored_host(self, P, y, out: Optional[np.ndarray] = None) -> np.ndarray:
        """numpy rows in, records (n, NF) out: the host API with 8 NF instead of 8 NSP^2 bytes per state back."""
        y = np.ascontiguousarray(y, dtype=np.float64)
        P = np.ascontiguousarray(np.broadcast_to(np.asarray(P, dtype=np.float64), (y.shape[0],)))
        assert y.ndim == 2 and y.shape[1] == self.NSP
        nf = self.factored_size[0]
        if out is None:
            out = np.empty((y.shape[0], nf))
        assert out.flags.c_contiguous and out.shape == (y.shape[0], nf)
        _lib.check(self.lib.pyjac_eval_jacob_factored_host(self._h, y.shape[0], P.ctypes.data, y.ctypes.data,
                                                           out.ctypes.data))
        return out

    # ------------------------------------------------------------------ host batch API
    def eval_jacob_host(self, P, y, out: Optional[np.ndarray] = None) -> np.ndarray:
        """numpy rows in, numpy rows out, streamed through pinned staging buffers."""
        y = np.ascontiguousarray(y, dtype=np.float64)
        P = np.ascontiguousarray(np.broadcast_to(np.asarray(P, dtype=np.float64), (y.shape[0],)))
        assert y.ndim == 2 and y.shape[1] == self.NSP
        if out is None:
            out = np.empty((y.shape[0], self.NSP * self.NSP))
        assert out.flags.c_contiguous and out.shape == (y.shape[0], self.NSP * self.NSP)
        _lib.check(self.lib.pyjac_eval_jacob_host(self._h, y.shape[0], P.ctypes.data, y.ctypes.data,
                                                  out.ctypes.data))
        return out

    def dydt_host(self, P, y, out: Optional[np.ndarray] = None) -> np.ndarray:
        y = np.ascontiguousarray(y, dtype=np.float64)
        P = np.ascontiguousarray(np.broadcast_to(np.asarray(P, dtype=np.float64), (y.shape[0],)))
        assert y.ndim == 2 and y.shape[1] == self.NSP
        if out is None:
            out = np.empty_like(y)
        _lib.check(self.lib.pyjac_dydt_host(self._h, y.shape[0], P.ctypes.data, y.ctypes.data,
                                            out.ctypes.data))
        return out
