import os, sys, torch
sys.path.insert(0, '/root/repo')
from pyjac_b200.evaluator import Evaluator
from pyjac_b200.mechanism import Mechanism
from pyjac_b200.states import synthetic_states
mech = Mechanism.from_chemkin('/root/repo/tests/golden/gri30_syn.inp')
ev = Evaluator(mech, 0)
n = 262144
P_h, y_h = synthetic_states(mech.NSP, n, seed=0)
P = torch.tensor(P_h, device='cuda'); y_rows = torch.tensor(y_h, device='cuda'); y_sf = y_rows.t().contiguous()
v_rows = torch.randn((n, mech.NSP), dtype=torch.float64, device='cuda'); v_sf = v_rows.t().contiguous()
def t(fn):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best
for lay, y, v in (('state_fastest', y_sf, v_sf), ('rows', y_rows, v_rows)):
    fac = ev.eval_jacob_factored(P, y, y_layout=lay, fac_layout=lay)
    x = torch.empty_like(v)
    t1 = t(lambda: ev.eval_jacob_factored(P, y, out=fac, y_layout=lay, fac_layout=lay))
    t2 = t(lambda: ev.newton_solve(fac, 1e-6, v, out=x, fac_layout=lay, v_layout=lay))
    t3 = t(lambda: ev.jvp(fac, v, out=x, fac_layout=lay, v_layout=lay))
    print('%-14s record %.2f ms (%.2e/s)  newton %.2f ms (%.2e/s)  jvp %.2f ms' % (lay, t1, n/t1*1e3, t2, n/t2*1e3, t3))
# mixed: record state-fastest y in, rows fac out
fac = ev.eval_jacob_factored(P, y_sf, y_layout='state_fastest', fac_layout='rows')
print('sf in / rows out record %.2f ms' % t(lambda: ev.eval_jacob_factored(P, y_sf, out=fac, y_layout='state_fastest', fac_layout='rows')))
