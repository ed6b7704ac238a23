#!/usr/bin/env python
"""eval_jacob through the record-stream kernel only, for compute-sanitizer."""
import os, sys
import torch
import _devlib  # noqa: F401,E402  (PYJAC_B200_LIB: development builds)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyjac_b200.evaluator import Evaluator
from pyjac_b200.mechanism import Mechanism
from pyjac_b200.states import synthetic_states

cases = sys.argv[1:] or ['h2o2_n2.inp:70:0', 'gri30_syn.inp:37:0']
for c in cases:
    f, n, gs, *rest = c.split(':')
    mech = Mechanism.from_chemkin(os.path.join(ROOT, 'tests', 'golden', f))
    P_h, y_h = synthetic_states(mech.NSP, int(n), seed=1)
    P, y = torch.tensor(P_h, device='cuda'), torch.tensor(y_h, device='cuda')
    ev = Evaluator(mech, 0, gs=int(gs), streams=True)
    assert ev.uses_streams
    if rest:
        ev.tune(int(rest[0]))
    a = ev.eval_jacob(P, y)
    torch.cuda.synchronize()
    print(f, 'gs', ev.plan_gs, 'finite', bool(torch.isfinite(a).all()), flush=True)
    ev.close()
