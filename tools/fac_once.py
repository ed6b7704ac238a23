#!/usr/bin/env python
"""One launch each of the factored-record kernel, k_jvp and k_newton on the GRI-shaped mechanism (for ncu)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyjac_b200.evaluator import Evaluator
from pyjac_b200.mechanism import Mechanism
from pyjac_b200.states import synthetic_states

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
mech = Mechanism.from_chemkin(os.path.join(ROOT, 'tests', 'golden', 'gri30_syn.inp'))
ev = Evaluator(mech, 0)
P_h, y_h = synthetic_states(mech.NSP, n, seed=0)
P, y = torch.tensor(P_h, device='cuda'), torch.tensor(y_h, device='cuda').t().contiguous()
v = torch.randn((mech.NSP, n), dtype=torch.float64, device='cuda')
for _ in range(2):
    fac = ev.eval_jacob_factored(P, y, y_layout='state_fastest', fac_layout='state_fastest')
    jv = ev.jvp(fac, v, fac_layout='state_fastest', v_layout='state_fastest')
    x, info = ev.newton_solve(fac, 1e-6, v, fac_layout='state_fastest', v_layout='state_fastest')
torch.cuda.synchronize()
print('ok', float(x.abs().max()), int(info.max()))
