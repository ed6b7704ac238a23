#!/usr/bin/env python
"""Development: run eval_jacob of a PJ_DEV build with the record checks of k_jac6 writing to pinned host memory."""
import os, sys
import torch
import _devlib  # noqa: F401,E402  (PYJAC_B200_LIB: development builds)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyjac_b200.evaluator import Evaluator
from pyjac_b200.mechanism import Mechanism
from pyjac_b200.states import synthetic_states

f, n, gs, *rest = sys.argv[1].split(':')
buf = torch.zeros(1 + 64 * 16, dtype=torch.int64).pin_memory()
os.environ['PYJAC_DEBUG_CLK'] = str(buf.data_ptr())
mech = Mechanism.from_chemkin(os.path.join(ROOT, 'tests', 'golden', f))
P_h, y_h = synthetic_states(mech.NSP, int(n), seed=1)
P, y = torch.tensor(P_h, device='cuda'), torch.tensor(y_h, device='cuda')
ev = Evaluator(mech, 0, gs=int(gs), streams=True)
if rest:
    ev.tune(int(rest[0]))
hdr = ev.tables['p6_hdr'].reshape(-1, 8)
print('hdr', hdr.tolist())
try:
    a = ev.eval_jacob(P, y)
    torch.cuda.synchronize()
    print('ok finite', bool(torch.isfinite(a).all()))
except Exception as exc:
    print('FAILED', str(exc).splitlines()[0])
cnt = int(buf[0])
print('events', cnt)
for i in range(min(cnt, 64)):
    e = buf[1 + 16 * i:1 + 16 * i + 12].tolist()
    print('blk %d tid %d (warp %d sub %d) phase %d grp %d chunks %d rec %d slot %d  rec=%08x %08x %08x %08x extra %d' %
          (e[0], e[1], e[1] // 32, (e[1] % 32) // 4, e[2], e[3], e[4], e[5], e[6], e[7], e[8], e[9], e[10], e[11]))
