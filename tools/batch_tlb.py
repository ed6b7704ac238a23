#!/usr/bin/env python
"""Why is eval_jacob 16 % slower at 2^22 states (profiles/r01_batch_sweep.md)?  Separates batch size from the
stride between consecutive Jacobian elements of a state (8 * ld bytes in the state-fastest layout):
    - 2^21 states with ld = 2^21 and with ld = 2^22 (same work, doubled stride)
    - 2^22 states state-fastest (stride 32 MB) and one-Jacobian-per-state ('rows': stride 8 bytes)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyjac_b200.evaluator import Evaluator
from pyjac_b200.mechanism import Mechanism
from pyjac_b200.states import synthetic_states

mech = Mechanism.from_chemkin(os.path.join(ROOT, 'tests', 'golden', 'gri30_syn.inp'))
ev = Evaluator(mech, 0)
nsp, nn = mech.NSP, mech.NSP ** 2
nmax = 1 << 22
P_h, y_h = synthetic_states(nsp, 1 << 16, seed=0)
P = torch.tensor(P_h, device='cuda').repeat(nmax >> 16)
y_sf = torch.tensor(y_h, device='cuda').t().contiguous().repeat(1, nmax >> 16)
y_rows = y_sf.t().contiguous()
out = torch.empty(nn * (nmax + 64), dtype=torch.float64, device='cuda')


def timed(fn):
    for _ in range(2):
        fn()
    best = 1e30
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


print('| states | layout | ld | stride between elements of a state | ms | states/s |\n|---|---|---|---|---|---|')
for n, ld in ((1 << 20, 1 << 20), (1 << 21, 1 << 21), (1 << 21, 1 << 22), (1 << 22, 1 << 22), (1 << 20, 1 << 22), (3 << 20, 3 << 20),
              (1 << 22, (1 << 22) + 64), (1 << 20, (1 << 22) + 64)):
    o = out[:nn * ld].view(nn, ld)
    yy = y_sf[:, :n].contiguous()
    ms = timed(lambda: ev.eval_jacob(P[:n], yy, o, y_layout='state_fastest', jac_layout='state_fastest'))
    print('| %d | state-fastest | %d | %.4f MB | %.2f | %.3e |' % (n, ld, ld * 8 / 2 ** 20, ms, n / ms * 1e3))
    del yy
for n in (1 << 20, 1 << 22):
    o = out[:nn * n].view(n, nn)
    yy = y_rows[:n].contiguous()
    ms = timed(lambda: ev.eval_jacob(P[:n], yy, o))
    print('| %d | rows | - | 8 B | %.2f | %.3e |' % (n, ms, n / ms * 1e3))
    del yy
