#!/usr/bin/env python
"""Per-warp, per-phase cycle counts of block 0 of k_eval in Jacobian mode (development tool).
Slots (a clock read may be scheduled before the barrier that precedes it, so a slot holds the
warp's own work plus the wait at the previous barrier): 0 A1, 1 B, 2 -, 3 C, 4 dots/A0 + class S,
5 class D, 6 class T, 7 -."""
import os, sys
# instrumented build: tools/devbuild.sh clk -DPJ_PHASE_CLOCKS, then PYJAC_B200_LIB=pyjac_b200/_build/dev_clk.so
if 'PYJAC_B200_LIB' not in os.environ:
    os.environ['PYJAC_B200_NVCC_EXTRA'] = '-DPJ_PHASE_CLOCKS'     # (set before the library is built)
import torch
import _devlib  # noqa: F401,E402  (PYJAC_B200_LIB: development builds)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyjac_b200.evaluator import Evaluator
from pyjac_b200.mechanism import Mechanism
from pyjac_b200.states import synthetic_states

mech = Mechanism.from_chemkin(os.path.join(ROOT, 'tests', 'golden', 'gri30_syn.inp'))
n = 131072
P_h, y_h = synthetic_states(mech.NSP, n, seed=0)
P = torch.tensor(P_h, device='cuda'); y = torch.tensor(y_h, device='cuda').t().contiguous()
out = torch.empty((mech.NSP ** 2, n), dtype=torch.float64, device='cuda')
clk = torch.zeros(32 * 8, dtype=torch.int64, device='cuda')
from pyjac_b200 import libgen
if 'PYJAC_B200_LIB' not in os.environ:
    libgen.build_library(force=True)
ev = Evaluator(mech, 0)
ev.eval_jacob(P, y, out, y_layout='state_fastest', jac_layout='state_fastest')
os.environ['PYJAC_DEBUG_CLK'] = str(clk.data_ptr())
ev.eval_jacob(P, y, out, y_layout='state_fastest', jac_layout='state_fastest')
torch.cuda.synchronize()
c = clk.cpu().numpy().reshape(32, 8)[:16, :8]
groups = (n // 8 + 147) // 148
print('groups per block', groups)
print('warp   A1     B      -      C      S      D      T      -    (cycles per group)')
for w in range(16):
    print('%3d ' % w + ' '.join('%6d' % (v / groups) for v in c[w]))
print('sum over phases (warp 0): %d cycles/group' % (c[0].sum() / groups))
