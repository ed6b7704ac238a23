#!/usr/bin/env python
"""Small eval_jacob / dydt / rates / factored-record / J v / Newton-solve launches for compute-sanitizer (memcheck, racecheck,
synccheck)."""
import os, sys
import torch
import _devlib  # noqa: F401,E402  (PYJAC_B200_LIB: development builds)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyjac_b200.evaluator import Evaluator
from pyjac_b200.mechanism import Mechanism
from pyjac_b200.states import synthetic_states

# (mechanism, states, working set forced into global memory?)
for f, n, g in (('gri30_syn.inp', 37, None), ('h2o2_n2.inp', 70, None), ('usc2_syn.inp', 5, False),
                ('usc2_syn.inp', 21, None), ('gri30_syn.inp', 37, True), ('plog.inp', 19, None), ('cheb.inp', 19, True)):
    mech = Mechanism.from_chemkin(os.path.join(ROOT, 'tests', 'golden', f))
    P_h, y_h = synthetic_states(mech.NSP, n, seed=1)
    P, y = torch.tensor(P_h, device='cuda'), torch.tensor(y_h, device='cuda')
    ev = Evaluator(mech, 0, ws_global=g)
    a = ev.eval_jacob(P, y)
    b = ev.eval_jacob(P, y.t().contiguous(), y_layout='state_fastest', jac_layout='state_fastest')
    c = ev.dydt(P, y)
    d = ev.rates(P, y, want_dy=True)
    torch.cuda.synchronize()
    assert torch.equal(a, b.t()) and torch.isfinite(a).all() and torch.isfinite(c).all()
    # the factored record and its consumers (k_eval M_FACT, k_jvp, k_newton: warp-level sharing of shared memory)
    if mech.NSP <= 160:
        v = torch.randn((n, mech.NSP), dtype=torch.float64, device='cuda')
        fac = ev.eval_jacob_factored(P, y)
        jv = ev.jvp(fac, v)
        x, info = ev.newton_solve(fac, 1e-6, v)
        yT, vT = y.t().contiguous(), v.t().contiguous()
        facT = ev.eval_jacob_factored(P, yT, y_layout='state_fastest', fac_layout='state_fastest')
        jvT = ev.jvp(facT, vT, fac_layout='state_fastest', v_layout='state_fastest')
        xT, infoT = ev.newton_solve(facT, 1e-6, vT, fac_layout='state_fastest', v_layout='state_fastest')
        torch.cuda.synchronize()
        assert torch.equal(fac, facT.t()) and torch.equal(jv, jvT.t()) and torch.equal(x, xT.t()) and int(info.max()) == 0
    print(f, 'ws_global', g, 'ok')
    ev.close()
