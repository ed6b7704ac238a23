"""GPU: the factored Jacobian record, J v from it, and the consumer step
x = (I - gamma J)^-1 r (8 f1) -- all through the C ABI, checked against the reference's golden Jacobians."""
import os

import numpy as np
import pytest

import gates
import kernel_model
from pyjac_b200 import factored
from pyjac_b200.mechanism import Mechanism
from pyjac_b200.states import synthetic_states

pytestmark = pytest.mark.gpu

CASES = [('h2o2_n2.inp', 'h2o2_pasr.npz', {}), ('torture.inp', 'torture_pasr.npz', {}),
         ('gri30_syn.inp', 'gri30_syn.npz', {}), ('usc2_syn.inp', 'usc2_syn.npz', {}),
         ('gri30_syn.inp', 'gri30_syn.npz', dict(ws_global=True)),
         ('plog.inp', 'plog_syn.npz', {}), ('nega.inp', 'nega_pasr.npz', {})]


@pytest.fixture(scope='module')
def torch():
    import torch
    if not torch.cuda.is_available():
        pytest.fail('GPU tests need a CUDA device (there is no CPU fallback to test)')
    return torch


def _setup(torch, golden_dir, mech_file, npz, kw, cap=256):
    from pyjac_b200.evaluator import Evaluator
    mech = Mechanism.from_chemkin(os.path.join(golden_dir, mech_file))
    ev = Evaluator(mech, **kw)
    g = {k: v[:cap] for k, v in np.load(os.path.join(golden_dir, npz)).items()}
    return mech, ev, g, torch.tensor(g['P'], device='cuda'), torch.tensor(g['y'], device='cuda')


@pytest.mark.parametrize('mech_file,npz,kw', CASES)
@pytest.mark.parametrize('layout', ['rows', 'state_fastest'])
def test_factored_record_vs_reference_golden(torch, golden_dir, mech_file, npz, kw, layout):
    mech, ev, g, P, y_rows = _setup(torch, golden_dir, mech_file, npz, kw)
    nsp, n = mech.NSP, y_rows.shape[0]
    nf, nnz = ev.factored_size
    assert nf == nsp + 3 * (nsp - 1) + nnz
    y = y_rows if layout == 'rows' else y_rows.t().contiguous()
    n0 = ev.launches
    if layout == 'rows':
        fac = ev.eval_jacob_factored(P, y)
        fac_h = fac.cpu().numpy()
    else:
        # a leading dimension larger than the batch: the padding must stay untouched
        buf = torch.full((nf, n + 6), float('nan'), dtype=torch.float64, device='cuda')
        fac = ev.eval_jacob_factored(P, y, out=buf, y_layout=layout, fac_layout=layout)
        assert torch.isnan(buf[:, n:]).all()
        fac_h = buf[:, :n].t().cpu().numpy()
    nm = ev.kernel_name(3)
    assert ev.launches == n0 + 1 and ('Li3E' in nm or ', 3, ' in nm), nm          # k_eval<.., M_FACT, ..>
    assert np.isfinite(fac_h).all()
    dense = ev.expand_factored(fac_h)
    gates.check_jac(dense, g['jac'], nsp, mech_file + ' factored', mech, g['y'])
    # and against the dense kernel of the same library: the same arithmetic, regrouped only in the last addition
    jac = ev.eval_jacob(P, y_rows).cpu().numpy()
    scale = np.abs(jac).reshape(n, nsp, nsp).max(axis=2, keepdims=True)
    err = np.abs(dense - jac).reshape(n, nsp, nsp) / (scale + 1e-300)
    assert err.max() <= 1e-14, err.max()
    ev.close()


@pytest.mark.parametrize('mech_file,npz,kw', CASES)
@pytest.mark.parametrize('layout', ['rows', 'state_fastest'])
def test_jvp_vs_reference_jacobian_times_v(torch, golden_dir, mech_file, npz, kw, layout):
    mech, ev, g, P, y_rows = _setup(torch, golden_dir, mech_file, npz, kw)
    nsp, n = mech.NSP, y_rows.shape[0]
    rng = np.random.default_rng(5)
    v_h = rng.standard_normal((n, nsp))
    y = y_rows if layout == 'rows' else y_rows.t().contiguous()
    v = torch.tensor(v_h if layout == 'rows' else np.ascontiguousarray(v_h.T), device='cuda')
    fac = ev.eval_jacob_factored(P, y, y_layout=layout, fac_layout=layout)
    out = ev.jvp(fac, v, fac_layout=layout, v_layout=layout).cpu().numpy()
    out = out if layout == 'rows' else out.T
    J = g['jac'].reshape(n, nsp, nsp)                       # [state, col, row]
    ref = np.einsum('ncr,nc->nr', J, v_h)
    mag = np.einsum('ncr,nc->nr', np.abs(J), np.abs(v_h))
    assert (np.abs(out - ref) <= gates.RTOL * mag + 1e-300).all(), (np.abs(out - ref) / (mag + 1e-300)).max()
    # the numpy statement of the same contraction
    rows, cols, ca, cb = ev.factored_pattern()
    fac_h = fac.cpu().numpy() if layout == 'rows' else fac.t().cpu().numpy()
    mine = kernel_model.factored_jvp(fac_h, v_h, nsp, rows, cols, ca, cb)
    assert (np.abs(out - mine) <= 1e-13 * mag + 1e-300).all()
    ev.close()


@pytest.mark.parametrize('mech_file,npz,kw', [c for c in CASES if c[0] != 'usc2_syn.inp'] +
                         [('usc2_syn.inp', 'usc2_syn.npz', {})])
@pytest.mark.parametrize('gamma', [1.0e-7, 1.0e-4])
def test_newton_solve_residual_against_reference_jacobian(torch, golden_dir, mech_file, npz, kw, gamma):
    """x = (I - gamma J)^-1 r on the device, J never stored; the residual is formed on the host
    with the REFERENCE's Jacobian (golden), so the solve and the record are both under test."""
    mech, ev, g, P, y = _setup(torch, golden_dir, mech_file, npz, kw, cap=64)
    nsp, n = mech.NSP, y.shape[0]
    rng = np.random.default_rng(9)
    r_h = rng.standard_normal((n, nsp)) * np.maximum(np.abs(g['y']), 1e-6)
    fac = ev.eval_jacob_factored(P, y)
    x, info = ev.newton_solve(fac, gamma, torch.tensor(r_h, device='cuda'))
    assert int(info.abs().max()) == 0
    x = x.cpu().numpy()
    J = g['jac'].reshape(n, nsp, nsp).transpose(0, 2, 1)    # [state, row, col]
    M = np.eye(nsp)[None] - gamma * J
    res = np.einsum('nrc,nc->nr', M, x) - r_h
    # normwise backward error of an LU solve with partial pivoting, plus the 1e-10 the Jacobian itself is held to
    bound = (np.abs(M) * np.abs(x)[:, None, :]).sum(axis=2) + np.abs(r_h)
    assert (np.abs(res) <= 1e-9 * bound.max(axis=1, keepdims=True)).all(), (np.abs(res) / bound.max(axis=1, keepdims=True)).max()
    # the same systems through LAPACK
    ref = np.linalg.solve(M, r_h[:, :, None])[:, :, 0]
    fwd = np.abs(x - ref).max(axis=1) / np.abs(ref).max(axis=1)
    cond = np.linalg.cond(M)
    assert (fwd <= 1e-9 * cond).all()
    # one gamma per state
    gam = torch.full((n,), gamma, dtype=torch.float64, device='cuda')
    x2, info2 = ev.newton_solve(fac, gam, torch.tensor(r_h, device='cuda'))
    assert np.array_equal(x2.cpu().numpy(), x) and int(info2.abs().max()) == 0
    ev.close()


def test_factored_host_api_and_ragged_batches(torch, golden_dir):
    from pyjac_b200.evaluator import Evaluator
    mech = Mechanism.from_chemkin(os.path.join(golden_dir, 'gri30_syn.inp'))
    ev = Evaluator(mech)
    P_h, y_h = synthetic_states(mech.NSP, 1003, seed=12)
    full = ev.eval_jacob_factored(torch.tensor(P_h, device='cuda'), torch.tensor(y_h, device='cuda')).cpu().numpy()
    host = ev.eval_jacob_factored_host(P_h, y_h)
    assert np.array_equal(host, full)
    for n in (1, 2, 3, 9, 64):
        part = ev.eval_jacob_factored(torch.tensor(P_h[:n], device='cuda'), torch.tensor(y_h[:n], device='cuda')).cpu().numpy()
        assert np.array_equal(part, full[:n])
    dense = ev.eval_jacob_host(P_h[:64], y_h[:64])
    err = np.abs(ev.expand_factored(full[:64]) - dense) / (np.abs(dense).max(axis=1, keepdims=True))
    assert err.max() <= 1e-14
    ev.close()


def test_newton_solve_reports_a_singular_matrix(torch, golden_dir):
    """gamma = 0 with a zero right-hand side is the identity; a record of NaNs has no pivot."""
    from pyjac_b200.evaluator import Evaluator
    mech = Mechanism.from_chemkin(os.path.join(golden_dir, 'h2o2_n2.inp'))
    ev = Evaluator(mech)
    nf = ev.factored_size[0]
    fac = torch.zeros((4, nf), dtype=torch.float64, device='cuda')
    r = torch.arange(4 * mech.NSP, dtype=torch.float64, device='cuda').reshape(4, mech.NSP)
    x, info = ev.newton_solve(fac, 0.5, r)
    assert torch.equal(x, r) and int(info.abs().max()) == 0
    fac[2] = float('nan')
    x = torch.full_like(r, -1.0)
    x, info = ev.newton_solve(fac, 0.5, r, out=x)
    assert info.cpu().tolist() == [0, 0, 1, 0] and torch.equal(x[2], torch.full_like(x[2], -1.0)) and torch.equal(x[3], r[3])
    ev.close()
