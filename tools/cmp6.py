#!/usr/bin/env python
"""Development: k_jac6 against k_eval on the same states (GRI-shaped), both layouts, ragged batch."""
import os, sys
import numpy as np
import torch
import _devlib  # noqa: F401,E402  (PYJAC_B200_LIB: development builds)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyjac_b200.evaluator import Evaluator
from pyjac_b200.mechanism import Mechanism
from pyjac_b200.states import synthetic_states

f, n, gs = (sys.argv[1].split(':') + ['8'])[:3] if len(sys.argv) > 1 else ('gri30_syn.inp', '4099', '8')
mech = Mechanism.from_chemkin(os.path.join(ROOT, 'tests', 'golden', f))
P_h, y_h = synthetic_states(mech.NSP, int(n), seed=4)
P, y = torch.tensor(P_h, device='cuda'), torch.tensor(y_h, device='cuda')
ev6, ev5 = Evaluator(mech, 0, gs=int(gs), streams=True), Evaluator(mech, 0, gs=int(gs), streams=False)
assert ev6.uses_streams and not ev5.uses_streams
a = ev6.eval_jacob(P, y).cpu().numpy().reshape(-1, mech.NSP, mech.NSP)
b = ev5.eval_jacob(P, y).cpu().numpy().reshape(-1, mech.NSP, mech.NSP)
err = np.abs(a - b) / (np.abs(b).max(axis=2, keepdims=True) + 1e-300)
print('rows: max |d|/colmax %.3e  nan %d' % (np.nanmax(err), np.isnan(a).sum()))
yt = y.t().contiguous()
c = ev6.eval_jacob(P, yt, y_layout='state_fastest', jac_layout='state_fastest').t().cpu().numpy().reshape(-1, mech.NSP, mech.NSP)
print('state-fastest == rows:', np.array_equal(a, c))
