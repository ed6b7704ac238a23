#!/usr/bin/env python
"""Sensitivity of eval_jacob to the static DE schedule: sweeps one cost constant of pyjac_b200/plan.py
(no rebuild: the schedule tables are made in Python).  usage: cost_probe.py NAME v1 v2 ..."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyjac_b200 import plan
from pyjac_b200.evaluator import Evaluator
from pyjac_b200.mechanism import Mechanism
from pyjac_b200.states import synthetic_states

name, vals = sys.argv[1], [float(v) for v in sys.argv[2:]]
mech = Mechanism.from_chemkin(os.path.join(ROOT, 'tests', 'golden', 'gri30_syn.inp'))
n = 262144
P_h, y_h = synthetic_states(mech.NSP, n, seed=0)
P, y = torch.tensor(P_h, device='cuda'), torch.tensor(y_h, device='cuda').t().contiguous()
out = torch.empty((mech.NSP ** 2, n), dtype=torch.float64, device='cuda')
for v in vals:
    setattr(plan, name, v)
    ev = Evaluator(mech, 0)
    for _ in range(2):
        ev.eval_jacob(P, y, out, y_layout='state_fastest', jac_layout='state_fastest')
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ev.eval_jacob(P, y, out, y_layout='state_fastest', jac_layout='state_fastest'); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print('%s = %-8g %.3f ms  %.3e states/s' % (name, v, best, n / best * 1e3), flush=True)
    ev.close()
