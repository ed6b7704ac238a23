#!/usr/bin/env python
"""eval_jacob throughput over mechanism sizes (development tool): GRI-3.0-, USC-Mech-II- and
n-heptane-shaped synthetic mechanisms, state-fastest layout, best of --reps launches."""
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
    ap.add_argument('--cases', default='gri30:262144,usc2:65536,usc2:65536:g8,nc7:8192')
    ap.add_argument('--reps', type=int, default=3)
    a = ap.parse_args()
    print('| mechanism | species | reactions | plan | states | ms | states/s | GB/s (algorithmic) |')
    print('|---|---|---|---|---|---|---|---|')
    for case in a.cases.split(','):
        parts = case.split(':')
        shape, n = parts[0], int(parts[1])
        # shape:n:gN[:tT] = working set in global memory, N states per block, T threads per block
        kw = dict(gs=int(parts[2][1:]), ws_global=True) if len(parts) > 2 else {}
        if len(parts) > 3:
            kw['threads'] = int(parts[3][1:])
        path = '/tmp/%s.inp' % shape
        synth.write(shape, path)
        mech = Mechanism.from_chemkin(path)
        ev = Evaluator(mech, 0, **kw)
        P_h, y_h = synthetic_states(mech.NSP, n, seed=0)
        P = torch.tensor(P_h, device='cuda')
        y = torch.tensor(y_h, device='cuda').t().contiguous()
        out = torch.empty((mech.NSP ** 2, n), dtype=torch.float64, device='cuda')
        ev.eval_jacob(P, y, out, y_layout='state_fastest', jac_layout='state_fastest')
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ev.eval_jacob(P, y, out, y_layout='state_fastest', jac_layout='state_fastest')
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        bps = 8 * mech.NSP ** 2 + 8 * (mech.NSP + 1)
        plan = 'gs=%d, %d threads, working set in %s' % (ev.plan_gs, ev.plan_threads,
                                                          'global memory' if int(ev.tables['p5_cfg'][14]) else 'shared memory')
        print('| %s | %d | %d | %s | %d | %.3f | %.3e | %.0f |' % (shape, mech.NSP, mech.FWD_RATES, plan, n, best,
                                                                  n / best * 1e3, n * bps / best / 1e6))
        ev.close()
        del out


if __name__ == '__main__':
    main()
