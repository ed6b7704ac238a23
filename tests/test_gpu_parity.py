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
         ('gri30_syn.inp', 'gri30_syn.npz'), ('usc2_syn.inp', 'usc2_syn.npz'),
         ('plog.inp', 'plog_syn.npz'), ('cheb.inp', 'cheb_syn.npz'), ('nega.inp', 'nega_pasr.npz'),
         ('mini.inp', 'mini_syn.npz')]
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
    gates.check_case(mech_file, worst, frac)
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


@pytest.mark.parametrize('mech_file,npz,gs,threads', [('gri30_syn.inp', 'gri30_syn.npz', 0, 0), ('gri30_syn.inp', 'gri30_syn.npz', 4, 256),
                                                      ('gri30_syn.inp', 'gri30_syn.npz', 8, 256), ('h2o2_n2.inp', 'h2o2_pasr.npz', 0, 0),
                                                      ('h2o2_n2.inp', 'h2o2_pasr.npz', 16, 128), ('h2o2_n2.inp', 'h2o2_pasr.npz', 8, 384),
                                                      ('torture.inp', 'torture_pasr.npz', 4, 64), ('torture.inp', 'torture_pasr.npz', 32, 512),
                                                      ('plog.inp', 'plog_syn.npz', 8, 256), ('cheb.inp', 'cheb_syn.npz', 16, 384)])
def test_stream_kernel_vs_table_kernel(torch, golden_dir, mech_file, npz, gs, threads):
    """eval_jacob through the record streams (k_jac6) against the golden vectors and against the
    schedule-table kernel (k_eval) on the same states, ragged batch included."""
    mech, ev = _evaluator_with(golden_dir, mech_file, gs=gs, threads=threads, streams=True)
    assert ev.uses_streams
    mech, ev5 = _evaluator_with(golden_dir, mech_file, gs=gs, threads=threads, streams=False)
    assert not ev5.uses_streams
    g = dict(np.load(os.path.join(golden_dir, npz)))
    P, y = torch.tensor(g['P'], device='cuda'), torch.tensor(g['y'], device='cuda')
    jac = ev.eval_jacob(P, y).cpu().numpy()
    worst, frac = gates.check_jac(jac, g['jac'], mech.NSP, mech_file + ' streams', mech, g['y'])
    print('%s gs=%d: |d|/colmax %.2e, elementwise <= 1e-10: %.5f' % (mech_file, ev.plan_gs, worst, frac))
    ref = ev5.eval_jacob(P, y).cpu().numpy()
    a, b = jac.reshape(-1, mech.NSP, mech.NSP), ref.reshape(-1, mech.NSP, mech.NSP)
    err = np.abs(a - b) / (np.abs(b).max(axis=2, keepdims=True) + 1e-300)
    assert err.max() <= 5e-11, err.max()
    n = len(g['P'])
    for m_ in (1, 3, min(n, 2 * ev.plan_gs + 1)):
        out = torch.full((mech.NSP ** 2, m_ + 1), float('nan'), dtype=torch.float64, device='cuda')
        ev.eval_jacob(P[:m_].contiguous(), y[:m_].t().contiguous(), out, y_layout='state_fastest', jac_layout='state_fastest')
        assert np.array_equal(out[:, :m_].t().cpu().numpy(), jac[:m_]), m_
        assert torch.isnan(out[:, m_]).all()
    ev.close()
    ev5.close()


def test_two_handles_share_one_kernel_instantiation(torch, golden_dir):
    """The opt-in shared-memory size belongs to the kernel instantiation, not to a handle: a large plan,
    then a small one, then the large one again (same instantiation) must all launch."""
    mech_l, ev_l = _evaluator_with(golden_dir, 'gri30_syn.inp', gs=8, threads=384, streams=True)
    mech_s, ev_s = _evaluator_with(golden_dir, 'h2o2_n2.inp', gs=8, threads=384, streams=True)
    for mech, ev in ((mech_l, ev_l), (mech_s, ev_s), (mech_l, ev_l), (mech_s, ev_s)):
        P_h, y_h = synthetic_states(mech.NSP, 40, seed=2)
        P, y = torch.tensor(P_h, device='cuda'), torch.tensor(y_h, device='cuda')
        assert np.isfinite(ev.eval_jacob(P, y).cpu().numpy()).all()
        assert np.isfinite(ev.dydt(P, y).cpu().numpy()).all()
    ev_l.close()
    ev_s.close()


@pytest.mark.parametrize('mech_file,npz', [('h2o2_n2.inp', 'h2o2_conv.npz'), ('gri30_syn.inp', 'gri30_conv.npz')])
@pytest.mark.parametrize('layout', ['rows', 'state_fastest'])
def test_constant_volume_dydt_vs_reference_golden(torch, golden_dir, mech_file, npz, layout):
    """pyjac_dydt_conv_dev against the reference's CONV dydt (golden vectors): dY/dt scaled by the gross
    production like the constant-pressure gate, dT/dt by sum |u_k W_k| gross_k / (rho cv_avg)."""
    from oracle.oracle import Oracle
    mech, ev = _evaluator(golden_dir, mech_file)
    g = dict(np.load(os.path.join(golden_dir, npz)))
    rho_h, y_h = g['rho'], g['y']
    rho, y = torch.tensor(rho_h, device='cuda'), torch.tensor(y_h, device='cuda')
    yy = y if layout == 'rows' else y.t().contiguous()
    dy = ev.dydt(rho, yy, y_layout=layout, conv=True).cpu().numpy()
    if layout != 'rows':
        dy = dy.T
    # gross rates from the oracle at the pressure the density implies
    Y = np.concatenate([y_h[:, 1:], 1.0 - y_h[:, 1:].sum(axis=1, keepdims=True)], axis=1)
    w = np.array([sp.mw for sp in mech.specs])
    from pyjac_b200.chem import RU
    P_h = rho_h * RU * y_h[:, 0] * (Y / w[None, :]).sum(axis=1)
    ora = Oracle(mech)
    conc, fwd, rev, pm, sr = ora.rates(P_h, y_h)
    gross = gates.gross_rates(mech, fwd, rev, pm)
    cp, h = gates._thermo(mech, y_h[:, 0])
    u = h - (RU / w)[None, :] * y_h[:, 0:1]
    cv_avg = (Y * (cp - (RU / w)[None, :])).sum(axis=1)
    scale = np.empty_like(dy)
    scale[:, 1:] = gross[:, :-1] * w[None, :-1] / rho_h[:, None]
    scale[:, 0] = (np.abs(u * w[None, :]) * gross).sum(axis=1) / (rho_h * cv_avg)
    d = np.abs(dy - g['dydt'])
    e = d / (scale + 1e-300)
    e[scale == 0] = d[scale == 0]
    assert e.max() <= gates.RTOL, e.max()
    # and it differs from the constant-pressure dT/dt
    cp_dy = ev.dydt(torch.tensor(P_h, device='cuda'), y).cpu().numpy()
    assert np.abs(cp_dy[:, 0] - g['dydt'][:, 0]).max() > 1e-3 * np.abs(g['dydt'][:, 0]).max()
    ev.close()


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


def test_plog_against_oracle_across_pressures(torch, golden_dir):
    """PLOG mechanism: 4096 states with pressures log-uniform over 0.001 - 1000 atm (below,
    between and above every pressure table), plus states exactly at the table pressures."""
    from oracle.oracle import Oracle
    mech, ev = _evaluator(golden_dir, 'plog.inp')
    P_h, y_h = synthetic_states(mech.NSP, 4096, seed=21)
    rng = np.random.default_rng(5)
    P_h = 101325.0 * 10.0 ** rng.uniform(-3.0, 3.0, size=P_h.shape)
    exact = sorted({e[0] for rx in mech.reacs if rx.plog for e in rx.plog_par})
    P_h[:len(exact)] = exact
    ora = Oracle(mech)
    ref = dict(zip(KEYS, ora.rates(P_h, y_h)))
    ref['dydt'] = ora.dydt(P_h, y_h)
    P, y = torch.tensor(P_h, device='cuda'), torch.tensor(y_h, device='cuda')
    new = dict(zip(KEYS + ['dydt'], [o.cpu().numpy() for o in ev.rates(P, y, want_dy=True)]))
    gates.check_rates(mech, P_h, y_h, new, ref, 'plog pressures')
    worst, frac = gates.check_jac(ev.eval_jacob(P, y).cpu().numpy(), ora.eval_jacob(P_h, y_h), mech.NSP,
                                  'plog pressures', mech, y_h)
    assert frac > 0.999, frac
    ev.close()


@pytest.mark.parametrize('mech_file,npz,gs', [('gri30_syn.inp', 'gri30_syn.npz', 8), ('torture.inp', 'torture_pasr.npz', 8),
                                              ('cheb.inp', 'cheb_syn.npz', 2), ('usc2_syn.inp', 'usc2_syn.npz', 8)])
def test_working_set_in_global_memory(torch, golden_dir, mech_file, npz, gs):
    """The plan variant for mechanisms too large for shared memory (per-block working set in global
    memory), forced on mechanisms that have golden vectors."""
    mech, ev = _evaluator_with(golden_dir, mech_file, gs=gs, ws_global=True)
    assert int(ev.tables['p5_cfg'][14]) == 1
    g = dict(np.load(os.path.join(golden_dir, npz)))
    P = torch.tensor(g['P'], device='cuda')
    y = torch.tensor(g['y'], device='cuda').t().contiguous()
    outs = ev.rates(P, y, y_layout='state_fastest', want_dy=True)
    jac = ev.eval_jacob(P, y, y_layout='state_fastest', jac_layout='state_fastest')
    dy2 = ev.dydt(P, y, y_layout='state_fastest')
    new = dict(zip(KEYS + ['dydt'], [o.cpu().numpy().T for o in outs]))
    gates.check_rates(mech, g['P'], g['y'], new, g, mech_file)
    gates.check_dydt(mech, g['y'], dy2.cpu().numpy().T, g, mech_file + ' dydt kernel')
    worst, frac = gates.check_jac(np.ascontiguousarray(jac.cpu().numpy().T), g['jac'], mech.NSP, mech_file, mech, g['y'])
    gates.check_case(mech_file, worst, frac)
    # the host-pointer API runs its chunks on two streams: they share the working sets
    jh = ev.eval_jacob_host(g['P'], g['y'])
    gates.check_jac(jh, g['jac'], mech.NSP, mech_file + ' host api', mech, g['y'])
    ev.close()


def test_usc2_plan_choice_and_shared_memory_plan(torch, golden_dir):
    """USC-II-shaped mechanism: the automatic plan keeps the working set in global memory (8 states
    per block); the shared-memory plan (2 states per block) stays available and agrees."""
    mech, ev = _evaluator(golden_dir, 'usc2_syn.inp')
    assert int(ev.tables['p5_cfg'][14]) == 1 and ev.plan_gs == 8
    mech, ev2 = _evaluator_with(golden_dir, 'usc2_syn.inp', ws_global=False)
    assert int(ev2.tables['p5_cfg'][14]) == 0 and ev2.plan_gs == 2
    g = dict(np.load(os.path.join(golden_dir, 'usc2_syn.npz')))
    P, y = torch.tensor(g['P'], device='cuda'), torch.tensor(g['y'], device='cuda')
    for e in (ev, ev2):
        gates.check_jac(e.eval_jacob(P, y).cpu().numpy(), g['jac'], mech.NSP, 'usc2', mech, g['y'])
        e.close()


def test_n_heptane_sized_mechanism_vs_oracle(torch, tmp_path):
    """654 species / 2827 reactions (the shape of the LLNL n-heptane mechanism): the working set of
    one state pair exceeds shared memory, so the plan puts it in global memory automatically."""
    from oracle.oracle import Oracle
    from pyjac_b200 import synth
    from pyjac_b200.evaluator import Evaluator
    path = str(tmp_path / 'nc7.inp')
    synth.write('nc7', path)
    mech = Mechanism.from_chemkin(path)
    ev = Evaluator(mech)
    assert int(ev.tables['p5_cfg'][14]) == 1 and ev.plan_gs == 16
    P_h, y_h = synthetic_states(mech.NSP, 20, seed=13)          # a tail group of 4 states
    ora = Oracle(mech)
    ref = dict(zip(KEYS, ora.rates(P_h, y_h)))
    ref['dydt'] = ora.dydt(P_h, y_h)
    P, y = torch.tensor(P_h, device='cuda'), torch.tensor(y_h, device='cuda')
    new = dict(zip(KEYS + ['dydt'], [o.cpu().numpy() for o in ev.rates(P, y, want_dy=True)]))
    gates.check_rates(mech, P_h, y_h, new, ref, 'nc7')
    worst, frac = gates.check_jac(ev.eval_jacob(P, y).cpu().numpy(), ora.eval_jacob(P_h, y_h), mech.NSP, 'nc7', mech, y_h)
    assert frac > 0.999, frac
    ev.close()


def test_n_heptane_sized_mechanism_vs_reference_golden(torch, golden_dir, tmp_path):
    """The same mechanism against the REFERENCE's generated C (tests/golden/nc7_syn.npz: rates and dydt in full, the
    Jacobian on all rows of 95 columns and all columns of 32 rows, with the per-column maxima the gate scales by)."""
    from pyjac_b200 import synth
    from pyjac_b200.evaluator import Evaluator
    path = str(tmp_path / 'nc7.inp')
    synth.write('nc7', path, seed=0)
    mech = Mechanism.from_chemkin(path)
    g = dict(np.load(os.path.join(golden_dir, 'nc7_syn.npz')))
    nsp = mech.NSP
    ev = Evaluator(mech)
    P, y = torch.tensor(g['P'], device='cuda'), torch.tensor(g['y'], device='cuda')
    new = dict(zip(KEYS + ['dydt'], [o.cpu().numpy() for o in ev.rates(P, y, want_dy=True)]))
    gates.check_rates(mech, g['P'], g['y'], new, g, 'nc7 golden')
    jac = ev.eval_jacob(P, y).cpu().numpy().reshape(-1, nsp, nsp)           # [state, col, row]
    # the gate of gates.check_jac on the stored sample: per-column maximum, the energy row's own scale
    cp, h = gates._thermo(mech, g['y'][:, 0])
    Y = np.concatenate([g['y'][:, 1:], 1.0 - g['y'][:, 1:].sum(axis=1, keepdims=True)], axis=1)
    cp_avg = (Y * cp).sum(axis=1)
    colmax = g['jac_colmax']                                                 # [state, col]
    for got, ref, cm in ((jac[:, g['cols'], :], g['jac_cols'], colmax[:, g['cols']][:, :, None]),
                         (jac[:, :, g['rows']], g['jac_rows'], colmax[:, :, None])):
        scale = np.broadcast_to(cm, ref.shape).copy()
        err = np.abs(got - ref) / (scale + 1e-300)
        rows0 = np.broadcast_to((np.arange(ref.shape[2]) if ref.shape[2] == nsp else g['rows'])[None, None, :] == 0, ref.shape)
        assert err[~rows0].max() <= gates.RTOL, err[~rows0].max()
    # energy-equation row of the sampled columns: scale max(colmax, sum_k |h_k J[k, j]| / cp_avg)
    ref_c = g['jac_cols']
    row0 = (np.abs(h[:, None, :nsp - 1]) * np.abs(ref_c[:, :, 1:])).sum(axis=2) / cp_avg[:, None]
    sc0 = np.maximum(colmax[:, g['cols']], row0)
    assert (np.abs(jac[:, g['cols'], 0] - ref_c[:, :, 0]) / (sc0 + 1e-300)).max() <= gates.RTOL
    rel = np.abs(jac[:, g['cols'], :] - ref_c) / (np.abs(ref_c) + 1e-300)
    frac = float((rel[ref_c != 0] <= gates.RTOL).mean())
    assert frac > 0.999, frac
    ev.close()


def test_fd_self_check(torch, golden_dir):
    """On-device finite-difference Jacobian of dydt (the reference's fd_jacob.cu comparison) against
    the analytical Jacobian: an oracle-free check.  Sixth-order central differences with CVODE-style
    increments; tolerance = what finite differences in fp64 can resolve."""
    for mech_file, n in (('gri30_syn.inp', 256), ('plog.inp', 512), ('cheb.inp', 512)):
        mech, ev = _evaluator(golden_dir, mech_file)
        P_h, y_h = synthetic_states(mech.NSP, n, seed=31)
        P = torch.tensor(P_h, device='cuda')
        y = torch.tensor(y_h, device='cuda').t().contiguous()
        # these states are far from equilibrium: the reference's r0 term would exceed the mass fractions
        err6 = ev.self_check(P, y, order=6, r_cap=1e-5)
        err1 = ev.self_check(P, y, order=1, r_cap=1e-5)
        print('%s: fd order 6 %.2e, order 1 %.2e' % (mech_file, err6, err1))
        assert err6 < 1e-3, (mech_file, err6)
        assert err1 < 0.2, (mech_file, err1)
        ev.close()
    # the bundled PaSR states (many exact-zero mass fractions: the reference's r0 / ewt increment
    # is unbounded there -- its finite-difference build is a timing comparison, not an accuracy one)
    mech, ev = _evaluator(golden_dir, 'h2o2_n2.inp')
    g = np.load(os.path.join(golden_dir, 'h2o2_pasr.npz'))
    P = torch.tensor(g['P'], device='cuda')
    y = torch.tensor(g['y'], device='cuda').t().contiguous()
    err = ev.self_check(P, y, order=6, r_cap=1e-5)
    print('h2o2 PaSR states: %.2e' % err)
    assert err < 1e-3, err
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


def test_pyjacob_wrappers_vs_reference_golden(torch, golden_dir, tmp_path):
    """The reference's Python surface (pyjacob / cu_pyjacob function names and calling
    conventions, pyjacob_wrapper.pyx:18-55, pyjacob_cuda_wrapper.pyx:13-34) on the PaSR states,
    through create_jacobian -> generate_library -> generate_wrapper."""
    from pyjac_b200 import libgen
    from pyjac_b200.create_jacobian import create_jacobian
    from pyjac_b200.pywrap import generate_wrapper
    out = str(tmp_path / 'out')
    mech = create_jacobian('cuda', os.path.join(golden_dir, 'h2o2_n2.inp'), build_path=out)
    libgen.generate_library('cuda', out)
    mod = generate_wrapper('cuda', out)
    g = dict(np.load(os.path.join(golden_dir, 'h2o2_pasr.npz')))
    nsp, nr, nrev, npd = mech.NSP, mech.FWD_RATES, mech.REV_RATES, mech.PRES_MOD_RATES

    # scalar API, a few states (each call is a batch of one on the GPU)
    pick = [0, 17, 511, 1019]
    new = {k: np.zeros((len(pick), w)) for k, w in
           (('conc', nsp), ('fwd', nr), ('rev', nrev), ('pres_mod', npd), ('spec_rates', nsp), ('dydt', nsp))}
    jac = np.zeros((len(pick), nsp * nsp))
    for r, s in enumerate(pick):
        y, P = np.ascontiguousarray(g['y'][s]), float(g['P'][s])
        mf = np.zeros(nsp)
        mf[:nsp - 1] = y[1:]
        mod.py_eval_conc(y[0], P, mf, 0.0, 0.0, new['conc'][r])
        assert abs(mf[-1] - (1.0 - y[1:].sum())) < 1e-15
        mod.py_eval_rxn_rates(y[0], P, new['conc'][r], new['fwd'][r], new['rev'][r])
        mod.py_get_rxn_pres_mod(y[0], P, new['conc'][r], new['pres_mod'][r])
        mod.py_eval_spec_rates(new['fwd'][r], new['rev'][r], new['pres_mod'][r], new['spec_rates'][r])
        mod.py_dydt(0.0, P, y, new['dydt'][r])
        mod.py_eval_jacobian(0.0, P, y, jac[r])
    sub = {k: v[pick] for k, v in g.items()}
    gates.check_rates(mech, sub['P'], sub['y'], new, sub, 'pyjacob scalar API')
    gates.check_jac(jac, sub['jac'], nsp, 'pyjacob scalar API', mech, sub['y'])

    # batched host API: Fortran-order (state-fastest) flat arrays
    num = g['y'].shape[0]
    padded = mod.py_cuinit(num)
    assert padded >= num
    flat = lambda a: np.ascontiguousarray(a.T).ravel()
    outs = {k: np.zeros(num * w) for k, w in
            (('conc', nsp), ('fwd', nr), ('rev', nrev), ('pres_mod', npd), ('spec_rates', nsp), ('dydt', nsp),
             ('jac', nsp * nsp))}
    mod.py_cujac(num, padded, np.ascontiguousarray(g['P']), flat(g['y']), outs['conc'], outs['fwd'], outs['rev'],
                 outs['pres_mod'], outs['spec_rates'], outs['dydt'], outs['jac'])
    mod.py_cuclean()
    back = {k: v.reshape(-1, num).T for k, v in outs.items()}
    gates.check_rates(mech, g['P'], g['y'], back, g, 'cu_pyjacob')
    worst, frac = gates.check_jac(np.ascontiguousarray(back['jac']), g['jac'], nsp, 'cu_pyjacob', mech, g['y'])
    gates.check_case('h2o2_n2.inp', worst, frac)
    mod.close()


def test_full_size_properties(torch, golden_dir):
    """BASELINE size (2^20 states, GRI-shaped): properties that need no oracle.  The batch is a
    small seeded set tiled many times, so every replica must equal the first one bitwise, and
    the first one is checked against the oracle."""
    from oracle.oracle import Oracle
    mech, ev = _evaluator(golden_dir, 'gri30_syn.inp')
    base, reps = 1024, 1024
    P_h, y_h = synthetic_states(mech.NSP, base, seed=21)
    P = torch.tensor(P_h, device='cuda').repeat(reps)
    y = torch.tensor(y_h, device='cuda').t().contiguous().repeat(1, reps)      # [NSP][n]
    n = base * reps
    jac = torch.empty((mech.NSP ** 2, n), dtype=torch.float64, device='cuda')
    ev.eval_jacob(P, y, jac, y_layout='state_fastest', jac_layout='state_fastest')
    torch.cuda.synchronize()
    first = jac[:, :base]
    assert torch.isfinite(first).all()
    for r in (1, 2, 513, reps - 1):
        assert torch.equal(jac[:, r * base:(r + 1) * base], first), r
    assert torch.equal(jac.view(mech.NSP ** 2, reps, base).amax(dim=1), first)
    assert torch.equal(jac.view(mech.NSP ** 2, reps, base).amin(dim=1), first)
    ref = Oracle(mech).eval_jacob(P_h, y_h)
    worst, frac = gates.check_jac(np.ascontiguousarray(first.t().cpu().numpy()), ref, mech.NSP,
                                  'full size, first replica', mech, y_h)
    assert frac > 0.999
    ev.close()


def test_speedtest_cli(torch, golden_dir, tmp_path, capsys):
    """The performance-test executable of the reference (tester.cu.in:51-168): data.bin rows in
    the original species order, masked on read, one 'N,ms' line on stdout."""
    from pyjac_b200 import speedtest
    from pyjac_b200.create_jacobian import create_jacobian
    out = str(tmp_path / 'out')
    mech = create_jacobian('cuda', os.path.join(golden_dir, 'h2o2_n2.inp'), build_path=out)
    g = dict(np.load(os.path.join(golden_dir, 'h2o2_pasr.npz')))
    n = g['y'].shape[0]
    # internal order -> original order (the inverse of apply_mask)
    Y_int = np.concatenate([g['y'][:, 1:], 1.0 - g['y'][:, 1:].sum(axis=1, keepdims=True)], axis=1)
    Y_orig = np.empty_like(Y_int)
    Y_orig[:, mech.fwd_spec_map] = Y_int
    data = str(tmp_path / 'data.bin')
    speedtest.write_data_bin(data, g['y'][:, 0], g['P'], Y_orig)
    assert speedtest.main([str(n), '4', '--build-path', out, '--data', data]) == 0
    line = capsys.readouterr().out.strip()
    num, ms = line.split(',')
    assert int(num) == n and float(ms) > 0.0
    ms2, jac = speedtest.run(n, out, data)
    gates.check_jac(np.ascontiguousarray(jac.T), g['jac'], mech.NSP, 'speedtest', mech, g['y'])


def test_edge_states_vs_oracle(torch, golden_dir):
    """States at the edges of what the reference's tests feed it: the NASA range limits
    (300 K / 3500 K and exactly T_mid), vacuum-like and 100 atm pressures, pure species and
    mixtures with exact zeros (most concentrations 0, so whole rate products vanish)."""
    from oracle.oracle import Oracle
    mech, ev = _evaluator(golden_dir, 'gri30_syn.inp')
    nsp = mech.NSP
    rng = np.random.default_rng(5)
    rows, pres = [], []
    for T in (300.0, 1000.0, 1000.0000001, 999.9999999, 3500.0):
        for P in (1.0e3, 101325.0, 1.0e7):
            Y = np.zeros(nsp)
            idx = rng.choice(nsp, size=4, replace=False)
            Y[idx] = rng.dirichlet(np.ones(4))
            Y[-1] = 1.0 - Y[:-1].sum()
            rows.append(np.concatenate([[T], Y[:-1]]))
            pres.append(P)
            Y = np.zeros(nsp)                     # one pure species (the last one gets 1 - sum = 0 or 1)
            k = int(rng.integers(nsp))
            Y[k] = 1.0
            rows.append(np.concatenate([[T], Y[:-1]]))
            pres.append(P)
    y_h, P_h = np.ascontiguousarray(rows), np.asarray(pres)
    ora = Oracle(mech)
    ref = dict(zip(KEYS, ora.rates(P_h, y_h)))
    ref['dydt'] = ora.dydt(P_h, y_h)
    ref_jac = ora.eval_jacob(P_h, y_h)
    assert np.isfinite(ref_jac).all()
    P, y = torch.tensor(P_h, device='cuda'), torch.tensor(y_h, device='cuda')
    new = dict(zip(KEYS + ['dydt'], [o.cpu().numpy() for o in ev.rates(P, y, want_dy=True)]))
    jac = ev.eval_jacob(P, y).cpu().numpy()
    assert np.isfinite(jac).all()
    gates.check_rates(mech, P_h, y_h, new, ref, 'edge states')
    gates.check_jac(jac, ref_jac, nsp, 'edge states', mech, y_h)
    ev.close()


@pytest.mark.parametrize('mech_file', ['mini.cti', 'mini.yaml'])
def test_cantera_format_mechanism_on_the_gpu_vs_reference_golden(torch, golden_dir, mech_file):
    """A mechanism read from a Cantera .cti / YAML file (no Cantera) runs through the kernel and matches what the
    REFERENCE's generated C gives for the Chemkin statement of the same mechanism (tests/golden/mini_syn.npz: third
    body, Troe, SRI, Lindemann with a specific collider, chemically activated, PLOG, Chebyshev, duplicates)."""
    from pyjac_b200.evaluator import Evaluator
    mech = Mechanism.from_file(os.path.join(golden_dir, mech_file))
    twin = Mechanism.from_file(os.path.join(golden_dir, 'mini.inp'))
    g = dict(np.load(os.path.join(golden_dir, 'mini_syn.npz')))
    ev = Evaluator(mech)
    P, y = torch.tensor(g['P'], device='cuda'), torch.tensor(g['y'], device='cuda')
    new = dict(zip(KEYS + ['dydt'], [o.cpu().numpy() for o in ev.rates(P, y, want_dy=True)]))
    gates.check_rates(twin, g['P'], g['y'], new, g, mech_file)
    worst, frac = gates.check_jac(ev.eval_jacob(P, y).cpu().numpy(), g['jac'], mech.NSP, mech_file, twin, g['y'])
    gates.check_case('mini.inp', worst, frac)
    ev.close()
