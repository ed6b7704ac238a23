#!/usr/bin/env python
"""Prints, per golden fixture, the worst scaled error and the elementwise-agreement fraction of the GPU
Jacobian (tests/gates.py check_jac) for the default plan and the global-memory plan: the measured
values behind gates.CASE_LIMITS.  Needs a GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import _devlib                                                 # noqa: E402,F401  (PYJAC_B200_LIB: development builds)
import gates                                                   # noqa: E402
import torch                                                   # noqa: E402
from pyjac_b200.evaluator import Evaluator                     # noqa: E402
from pyjac_b200.mechanism import Mechanism                     # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
CASES = [('h2o2_n2.inp', 'h2o2_pasr.npz'), ('torture.inp', 'torture_pasr.npz'), ('gri30_syn.inp', 'gri30_syn.npz'),
         ('usc2_syn.inp', 'usc2_syn.npz'), ('plog.inp', 'plog_syn.npz'), ('cheb.inp', 'cheb_syn.npz'),
         ('nega.inp', 'nega_pasr.npz')]
print('| fixture | plan | states | worst abs(d)/scale | elementwise <= 1e-10 | limits (cap, floor) |')
print('|---|---|---|---|---|---|')
for mf, npz in CASES:
    mech = Mechanism.from_chemkin(os.path.join(GOLD, mf))
    g = dict(np.load(os.path.join(GOLD, npz)))
    P, y = torch.tensor(g['P'], device='cuda'), torch.tensor(g['y'], device='cuda')
    for kw, label in (({}, 'default'), (dict(ws_global=True, gs=8), 'global memory'), (dict(streams=True), 'streams')):
        try:
            ev = Evaluator(mech, **kw)
        except Exception as exc:
            print('| %s | %s | - | %s | | |' % (mf, label, str(exc).splitlines()[0][:60]))
            continue
        jac = ev.eval_jacob(P, y).cpu().numpy()
        worst, frac = gates.check_jac(jac, g['jac'], mech.NSP, mf, mech, g['y'])
        print('| %s | %s gs=%d | %d | %.2e | %.5f | %s |' % (mf, label, ev.plan_gs, len(g['P']), worst, frac, gates.CASE_LIMITS[mf]))
        ev.close()
