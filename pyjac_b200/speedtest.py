"""``speedtest`` -- the reference's performance-test executable, over the B200 library.

The reference compiles ``tester.cu.in`` per mechanism into ``speedtest`` and its driver runs
``speedtest <num_odes> [<num_threads>]`` (pyjac/performance_tester/tester.cu.in:51-168,
performance_tester.py:500-508).  It reads ``data.bin`` -- rows of NN + 2 doubles
``[t, T, P, Y_0 .. Y_{NSP-1}]`` in the *original* species order, permuted on read by
``apply_mask`` (read_initial_conditions.cu:10-59, mech_auxiliary.py:189-196) --, evaluates
the Jacobian of the first ``num_odes`` states **including** the host<->device transfers, and
prints one line ``"%d,%.15le\\n" % (num_odes, elapsed_ms)``.

    python -m pyjac_b200.speedtest NUM_ODES [NUM_THREADS] [--build-path out] [--data data.bin]

``--build-path`` is a directory written by :func:`pyjac_b200.create_jacobian.create_jacobian`
(the reference bakes mechanism and data file into the binary instead); ``NUM_THREADS`` is
accepted for command-line compatibility and ignored (the GPU path has no OpenMP threads).
"""
from __future__ import annotations

import argparse
import os
import re
import sys
import time
from typing import List, Optional, Tuple

import numpy as np

from . import create_jacobian as _cj
from .pywrap import PyJacob


def write_data_bin(path: str, T, P, Y_original, t=None) -> None:
    """``data.bin`` as the reference's tools write it (performance_tester.py:320-338): one row
    ``[t, T, P, Y...]`` per state, all NSP mass fractions in the original species order."""
    T = np.asarray(T, dtype=np.float64)
    n = T.shape[0]
    rows = np.empty((n, Y_original.shape[1] + 3))
    rows[:, 0] = 0.0 if t is None else t
    rows[:, 1] = T
    rows[:, 2] = P
    rows[:, 3:] = Y_original
    rows.tofile(path)


def _fwd_spec_map(build_path: str, nsp: int) -> List[int]:
    """Internal position -> original index, from the ``//last_spec`` comment of mechanism.h
    (the move-to-end permutation of utils.get_species_mappings, utils.py:55-91)."""
    last = nsp - 1
    with open(os.path.join(build_path, _cj.HEADER_FILE)) as fh:
        for line in fh:
            m = re.search(r'^//last_spec (\d+)$', line)
            if m:
                last = int(m.group(1))
    return [i for i in range(nsp) if i != last] + [last]


def read_initial_conditions(path: str, num: int, fwd_map: List[int]) -> Tuple[np.ndarray, np.ndarray]:
    """(y, pres): y is (NSP, num) state-fastest -- T then Y_0..Y_{NSP-2} in internal order --,
    exactly the rows the reference copies to the device."""
    nsp = len(fwd_map)
    want = num * (nsp + 3)
    buf = np.fromfile(path, dtype=np.float64, count=want)
    if buf.size != want:
        sys.stderr.write('File (%s) is incorrectly formatted, %d doubles were expected but only %d were read.\n'
                         % (path, want, buf.size))
        sys.exit(-1)
    rows = buf.reshape(num, nsp + 3)
    Y = rows[:, 3:][:, fwd_map]                                   # apply_mask
    y = np.empty((nsp, num))
    y[0] = rows[:, 1]
    y[1:] = Y[:, :nsp - 1].T
    return np.ascontiguousarray(y), np.ascontiguousarray(rows[:, 2])


def run(num_odes: int, build_path: str, data: str, device: Optional[int] = None):
    """Returns (elapsed_ms, jac) with jac (NSP*NSP, num_odes) state-fastest."""
    mod = PyJacob(_cj.load_tables(build_path), device)
    try:
        y, pres = read_initial_conditions(data, num_odes, _fwd_spec_map(build_path, mod.NSP))
        jac = np.empty((mod.NSP * mod.NSP, num_odes))
        padded = mod.py_cuinit(num_odes)
        t0 = time.perf_counter()                                  # StartTimer(): transfers included
        mod.py_cujac(num_odes, padded, pres, y.ravel(), None, None, None, None, None, None, jac.ravel())
        ms = (time.perf_counter() - t0) * 1e3
        mod.py_cuclean()
    finally:
        mod.close()
    return ms, jac


def main(argv: Optional[List[str]] = None) -> int:
    ap = argparse.ArgumentParser(prog='speedtest')
    ap.add_argument('num_odes', type=int)
    ap.add_argument('num_threads', type=int, nargs='?', default=None)
    ap.add_argument('--build-path', default='out')
    ap.add_argument('--data', default='data.bin')
    a = ap.parse_args(argv)
    if a.num_odes <= 0:
        return 1
    ms, _ = run(a.num_odes, a.build_path, a.data)
    sys.stdout.write('%d,%.15e\n' % (a.num_odes, ms))
    return 0


if __name__ == '__main__':
    sys.exit(main())
