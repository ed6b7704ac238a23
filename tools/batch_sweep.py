#!/usr/bin/env python
"""eval_jacob throughput against batch size on one GPU (BASELINE.json config 5, GRI-shaped)."""
import os, sys
import torch
import _devlib  # noqa: F401,E402  (PYJAC_B200_LIB: development builds)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyjac_b200.evaluator import Evaluator
from pyjac_b200.mechanism import Mechanism
from pyjac_b200.states import synthetic_states

mech = Mechanism.from_chemkin(os.path.join(ROOT, 'tests', 'golden', 'gri30_syn.inp'))
ev = Evaluator(mech, 0)
nn, bps = mech.NSP ** 2, 8 * mech.NSP ** 2 + 8 * (mech.NSP + 1)
nmax = 1 << 22
P_h, y_h = synthetic_states(mech.NSP, 1 << 16, seed=0)
P = torch.tensor(P_h, device='cuda').repeat(nmax >> 16)
y = torch.tensor(y_h, device='cuda').t().contiguous().repeat(1, nmax >> 16)
out = torch.empty((nn, nmax), dtype=torch.float64, device='cuda')
print('| states | ms | states/s | GB/s (algorithmic) |\n|---|---|---|---|')
for e in range(10, 23):
    n = 1 << e
    yy, oo, pp = y[:, :n].contiguous(), out[:, :n].contiguous() if n < nmax else out, P[:n]
    for _ in range(3):
        ev.eval_jacob(pp, yy, oo, y_layout='state_fastest', jac_layout='state_fastest')
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ev.eval_jacob(pp, yy, oo, y_layout='state_fastest', jac_layout='state_fastest')
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print('| %d | %.3f | %.3e | %.0f |' % (n, best, n / best * 1e3, n * bps / best / 1e6))
    del yy, oo
