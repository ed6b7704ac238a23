#!/usr/bin/env python
"""Launch-shape sweep for eval_jacob on one GPU (development tool, not the bench)."""
import argparse
import os
import sys

import torch
import _devlib  # noqa: F401,E402  (PYJAC_B200_LIB: development builds)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyjac_b200 import synth                       # noqa: E402
from pyjac_b200.evaluator import Evaluator         # noqa: E402
from pyjac_b200.mechanism import Mechanism         # noqa: E402
from pyjac_b200.states import synthetic_states     # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--mech', default=os.path.join(ROOT, 'tests', 'golden', 'gri30_syn.inp'))
    ap.add_argument('--shape', default=None, help='synthetic shape name instead of --mech')
    ap.add_argument('--n', type=int, default=262144)
    ap.add_argument('--configs', default='8:512:0,8:256:0,4:512:0,4:256:0,4:384:0,8:384:0,2:256:0',
                    help='gs:threads:blocks_per_sm[:streams] of the Jacobian plan (streams = 0: schedule-table kernel)')
    ap.add_argument('--layout', default='state_fastest')
    ap.add_argument('--reps', type=int, default=5)
    a = ap.parse_args()
    if a.shape:
        path = '/tmp/%s.inp' % a.shape
        synth.write(a.shape, path)
        a.mech = path
    mech = Mechanism.from_chemkin(a.mech)
    P_h, y_h = synthetic_states(mech.NSP, a.n, seed=0)
    P = torch.tensor(P_h, device='cuda')
    y = torch.tensor(y_h, device='cuda')
    if a.layout != 'rows':
        y = y.t().contiguous()
    nn = mech.NSP ** 2
    out = torch.empty((a.n, nn) if a.layout == 'rows' else (nn, a.n), dtype=torch.float64, device='cuda')
    bytes_per_state = 8 * nn + 8 * (mech.NSP + 1)
    print('mech NSP=%d NR=%d n=%d  bytes/state=%d' % (mech.NSP, mech.FWD_RATES, a.n, bytes_per_state))
    for cfg in a.configs.split(','):
        G, th, bp, *rest = (int(v) for v in cfg.split(':'))
        try:
            ev = Evaluator(mech, 0, gs=G, threads=th, streams=bool(rest[0]) if rest else None)
            ev.tune(bp)
            ev.eval_jacob(P, y, out, y_layout=a.layout, jac_layout=a.layout)
            torch.cuda.synchronize()
        except Exception as exc:
            print('gs=%d threads=%d bpsm=%d: %s' % (G, th, bp, exc))
            continue
        best = 1e30
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ev.eval_jacob(P, y, out, y_layout=a.layout, jac_layout=a.layout)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print('gs=%d threads=%3d bpsm=%d streams=%d: %8.3f ms  %.3e states/s  %7.1f GB/s' %
              (G, th, bp, ev.uses_streams, best, a.n / best * 1e3, a.n * bytes_per_state / best / 1e6))
    # dydt for reference
    ev = Evaluator(mech, 0)
    dy = torch.empty_like(y)
    ev.dydt(P, y, dy, y_layout=a.layout)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ev.dydt(P, y, dy, y_layout=a.layout)
    e1.record()
    torch.cuda.synchronize()
    print('dydt: %.3f ms  %.3e states/s' % (e0.elapsed_time(e1), a.n / e0.elapsed_time(e1) * 1e3))


if __name__ == '__main__':
    main()
