#!/usr/bin/env python
"""The reference's own generated CUDA (rebuilt for sm_100a by oracle/build_ref.py build_cuda)
on this GPU, next to pyjac_b200: a reported baseline (BASELINE.json config 5) and a
cross-check of the two GPU implementations.  Development / measurement tool: it executes
oracle/_ref, which the product path never does."""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch
import _devlib  # noqa: F401,E402  (PYJAC_B200_LIB: development builds)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import gates                                        # noqa: E402
from oracle import build_ref                        # noqa: E402
from pyjac_b200.evaluator import Evaluator          # noqa: E402
from pyjac_b200.mechanism import Mechanism          # noqa: E402
from pyjac_b200.states import synthetic_states      # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sizes', default='1024,16384,65536,262144')
    a = ap.parse_args()
    lib_path = build_ref.cuda_lib_path('gri30')
    if not os.path.exists(lib_path):
        raise SystemExit('missing %s (build it where /root/reference exists)' % lib_path)
    ref = ctypes.CDLL(lib_path)
    ref.refcu_run.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 3 + [ctypes.c_int] + [ctypes.POINTER(ctypes.c_double)] * 2
    mech = Mechanism.from_chemkin(os.path.join(ROOT, 'tests', 'golden', 'gri30_syn.inp'))
    nsp = mech.NSP
    assert ref.refcu_nsp() == nsp
    ev = Evaluator(mech, 0)
    print('| states | reference CUDA kernel ms | states/s | with transfers ms | states/s | pyjac_b200 kernel ms | states/s | ratio (kernel) |')
    print('|---|---|---|---|---|---|---|---|')
    for n in (int(v) for v in a.sizes.split(',')):
        P_h, y_h = synthetic_states(nsp, n, seed=0)
        y_sf = np.ascontiguousarray(y_h.T)
        jac_ref = np.empty((nsp * nsp, n)) if n <= 65536 else None
        k_ms, t_ms = ctypes.c_double(), ctypes.c_double()
        rc = ref.refcu_run(n, P_h.ctypes.data, y_sf.ctypes.data, None if jac_ref is None else jac_ref.ctypes.data,
                           3, ctypes.byref(k_ms), ctypes.byref(t_ms))
        if rc:
            print('| %d | reference failed: %d |' % (n, rc))
            continue
        P, y = torch.tensor(P_h, device='cuda'), torch.tensor(y_sf, device='cuda')
        out = torch.empty((nsp * nsp, n), dtype=torch.float64, device='cuda')
        best = 1e30
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ev.eval_jacob(P, y, out, y_layout='state_fastest', jac_layout='state_fastest')
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        if jac_ref is not None:       # the two GPU implementations agree within the parity gate
            gates.check_jac(np.ascontiguousarray(out.t().cpu().numpy()), np.ascontiguousarray(jac_ref.T), nsp,
                            'reference CUDA vs pyjac_b200, n=%d' % n, mech, y_h)
        print('| %d | %.3f | %.3e | %.1f | %.3e | %.3f | %.3e | %.1f |' %
              (n, k_ms.value, n / k_ms.value * 1e3, t_ms.value, n / t_ms.value * 1e3, best, n / best * 1e3,
               k_ms.value / best))
        del out


if __name__ == '__main__':
    main()
