"""GPU parity: the CUDA path, called through the C ABI, against (a) the committed golden
vectors made by the reference's own generated C and (b) the CPU oracle on seeded states."""
import os

import numpy as np
import pytest

import gates
from pyjac_b200.mechanism import Mechanism
from pyjac_b200.states import synthetic_states

pytestmark = pytest.mark.gpu

CASES = [('h2o2_n2.inp', 'h2o2_pasr.npz'), ('torture.inp', 'torture_pasr.npz'),
         ('gri30_syn.inp', 'gri30_syn.npz')]
KEYS = ['conc', 'fwd', 'rev', 'pres_mod', 'spec_rates']


@pytest.fixture(scope='module')
def torch():
    import torch
    if not torch.cuda.is_available():
        pytest.fail('GPU tests need a CUDA device (there is no CPU fallback to test)')
    return torch


def _evaluator(golden_dir, mech_file):
    from pyjac_b200.evaluator import Evaluator
    mech = Mechanism.from_chemkin(os.path.join(golden_dir, mech_file))
    return mech, Evaluator(mech)


@pytest.mark.parametrize('mech_file,npz', CASES)
@pytest.mark.parametrize('layout', ['rows', 'state_fastest'])
def test_device_api_vs_reference_golden(torch, golden_dir, mech_file, npz, layout):
    mech, ev = _evaluator(golden_dir, mech_file)
    g = dict(np.load(os.path.join(golden_dir, npz)))
    P = torch.tensor(g['P'], device='cuda')
    y_rows = torch.tensor(g['y'], device='cuda')
    y = y_rows if layout == 'rows' else y_rows.t().contiguous()
    n0 = ev.launches
    outs = ev.rates(P, y, y_layout=layout, want_dy=True)
    jac = ev.eval_jacob(P, y, y_layout=layout, jac_layout=layout)
    dy2 = ev.dydt(P, y, y_layout=layout)
    torch.cuda.synchronize()
    assert ev.launches == n0 + 3
    host = [o.cpu().numpy() for o in outs]
    jac, dy2 = jac.cpu().numpy(), dy2.cpu().numpy()
    if layout == 'state_fastest':
        host = [o.T for o in host]
        jac, dy2 = jac.T, dy2.T
    new = dict(zip(KEYS + ['dydt'], host))
    gates.check_rates(mech, g['P'], g['y'], new, g, mech_file)
    gates.check_dydt(mech, g['y'], dy2, g, mech_file + ' dydt kernel')
    worst, frac = gates.check_jac(np.ascontiguousarray(jac), g['jac'], mech.NSP, mech_file, mech, g['y'])
    assert frac > 0.97
    ev.close()


def _evaluator_with(golden_dir, mech_file, **kw):
    from pyjac_b200.evaluator import Evaluator
    mech = Mechanism.from_chemkin(os.path.join(golden_dir, mech_file))
    return mech, Evaluator(mech, **kw)


@pytest.mark.parametrize('gs,threads', [(8, 256), (4, 512), (4, 256), (2, 128), (2, 512), (8, 64)])
@pytest.mark.parametrize('layout', ['rows', 'state_fastest'])
def test_launch_shapes_give_identical_results(torch, golden_dir, gs, threads, layout):
    """Results must not depend on states per block / block size (the plan); ragged batch
    sizes and odd leading dimensions included."""
    mech, ev0 = _evaluator(golden_dir, 'gri30_syn.inp')
    P_h, y_h = synthetic_states(mech.NSP, 203, seed=3)
    P, y = torch.tensor(P_h, device='cuda'), torch.tensor(y_h, device='cuda')

    def run(ev, n, ld=None):
        yy = y[:n].contiguous()
        if layout == 'rows':
            return ev.eval_jacob(P[:n].contiguous(), yy).cpu().numpy()
        out = torch.full((mech.NSP ** 2, ld or n), float('nan'), dtype=torch.float64, device='cuda')
        ev.eval_jacob(P[:n].contiguous(), yy.t().contiguous(), out, y_layout=layout, jac_layout=layout)
        return out[:, :n].t().cpu().numpy()

    ref = run(ev0, 203)
    assert np.isfinite(ref).all()
    mech, ev = _evaluator_with(golden_dir, 'gri30_syn.inp', gs=gs, threads=threads)
    assert (ev.plan_gs, ev.plan_threads) == (gs, threads)
    # Different plans sum in different orders: agreement to rounding, not bitwise.
    full = run(ev, 203)
    a, b = full.reshape(-1, mech.NSP, mech.NSP), ref.reshape(-1, mech.NSP, mech.NSP)
    err = np.abs(a - b) / (np.abs(b).max(axis=2, keepdims=True) + 1e-300)
    assert err.max() <= 5e-11, (gs, threads, err.max())
    # Within one plan the result of a state does not depend on the batch around it.
    for n, ld in ((203, 205), (1, 1), (2, 2), (3, 4), (5, 7), (64, 64)):
        out = run(ev, n, ld)
        assert np.array_equal(out, full[:n]), (gs, threads, n)
    ev.close()
    ev0.close()


def test_against_oracle_on_synthetic_states(torch, golden_dir):
    """GRI-3.0-shaped mechanism, 2048 seeded synthetic states, oracle = CPU restatement."""
    from oracle.oracle import Oracle
    mech, ev = _evaluator(golden_dir, 'gri30_syn.inp')
    P_h, y_h = synthetic_states(mech.NSP, 2048, seed=11)
    ora = Oracle(mech)
    ref = dict(zip(KEYS, ora.rates(P_h, y_h)))
    ref['dydt'] = ora.dydt(P_h, y_h)
    ref_jac = ora.eval_jacob(P_h, y_h)
    P, y = torch.tensor(P_h, device='cuda'), torch.tensor(y_h, device='cuda')
    new = dict(zip(KEYS + ['dydt'], [o.cpu().numpy() for o in ev.rates(P, y, want_dy=True)]))
    gates.check_rates(mech, P_h, y_h, new, ref, 'gri30 synthetic')
    worst, frac = gates.check_jac(ev.eval_jacob(P, y).cpu().numpy(), ref_jac, mech.NSP, 'gri30 synthetic', mech, y_h)
    assert frac > 0.999, frac
    ev.close()


def test_empty_batch(torch, golden_dir):
    mech, ev = _evaluator(golden_dir, 'h2o2_n2.inp')
    P = torch.empty(0, dtype=torch.float64, device='cuda')
    y = torch.empty((0, mech.NSP), dtype=torch.float64, device='cuda')
    assert ev.eval_jacob(P, y).shape == (0, mech.NSP ** 2)
    assert ev.dydt(P, y).shape == (0, mech.NSP)
    ev.close()


def test_host_batch_api(torch, golden_dir):
    mech, ev = _evaluator(golden_dir, 'h2o2_n2.inp')
    g = np.load(os.path.join(golden_dir, 'h2o2_pasr.npz'))
    jac = ev.eval_jacob_host(g['P'], g['y'])
    gates.check_jac(jac, g['jac'], mech.NSP, 'host api', mech, g['y'])
    dy = ev.dydt_host(g['P'], g['y'])
    gates.check_dydt(mech, g['y'], dy, g)
    # device and host entry points run the same kernel
    dev = ev.eval_jacob(torch.tensor(g['P'], device='cuda'), torch.tensor(g['y'], device='cuda'))
    assert np.array_equal(dev.cpu().numpy(), jac)
    ev.close()
