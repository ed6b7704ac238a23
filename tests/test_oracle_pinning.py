"""Pin the CPU restatement oracle (oracle/pyjac_oracle.c) to the REAL reference.

The golden .npz files under tests/golden/ hold outputs of the reference's own generated C
(made by tests/golden/make_golden.py in the build container).  The restatement follows the
emitted code's evaluation order, so agreement is expected to be bit-exact; the assertion
allows a few ulp so that a different libm build cannot turn it red.
"""
import os

import numpy as np
import pytest

from oracle.oracle import Oracle, RefLib
from pyjac_b200.mechanism import Mechanism

CASES = [('h2o2_n2.inp', 'h2o2_pasr.npz', 'h2o2'),
         ('torture.inp', 'torture_pasr.npz', 'torture'),
         ('gri30_syn.inp', 'gri30_syn.npz', 'gri30'),
         ('usc2_syn.inp', 'usc2_syn.npz', 'usc2'),
         ('plog.inp', 'plog_syn.npz', 'plog'),
         ('cheb.inp', 'cheb_syn.npz', 'cheb'),
         ('nega.inp', 'nega_pasr.npz', 'nega'),
         ('mini.inp', 'mini_syn.npz', 'mini')]
KEYS = ['conc', 'fwd', 'rev', 'pres_mod', 'spec_rates']


def _close(a, b, what):
    assert a.shape == b.shape, what
    assert np.isfinite(a).all(), what
    scale = np.abs(b).max(axis=-1, keepdims=True) + 1e-300
    err = (np.abs(a - b) / scale).max()
    assert err <= 4e-15, '%s: |d|/rowmax = %.3e' % (what, err)
    return float((a == b).mean())


@pytest.mark.parametrize('mech_file,npz,name', CASES)
def test_oracle_matches_reference_golden(golden_dir, mech_file, npz, name):
    mech = Mechanism.from_chemkin(os.path.join(golden_dir, mech_file))
    g = np.load(os.path.join(golden_dir, npz))
    ora = Oracle(mech)
    assert (ora.NSP, ora.NR, ora.NREV, ora.NPD) == (
        g['y'].shape[1], g['fwd'].shape[1], g['rev'].shape[1], g['pres_mod'].shape[1])
    P, y = g['P'], g['y']
    exact = {}
    for key, arr in zip(KEYS, ora.rates(P, y)):
        exact[key] = _close(arr, g[key], name + ':' + key)
    exact['dydt'] = _close(ora.dydt(P, y), g['dydt'], name + ':dydt')
    jac = ora.eval_jacob(P, y)
    assert (g['jac'] != 0).mean() > 0.5
    exact['jac'] = _close(jac, g['jac'], name + ':jac')
    # every output was bit-identical when the goldens were made
    assert min(exact.values()) > 0.999, exact


@pytest.mark.parametrize('mech_file,npz,name', CASES)
def test_oracle_matches_live_reference_build(golden_dir, mech_file, npz, name):
    """When oracle/_ref/<name> was built here (reference present), run it live."""
    if not RefLib.available(name):
        pytest.skip('oracle/_ref/%s not built' % name)
    mech = Mechanism.from_chemkin(os.path.join(golden_dir, mech_file))
    g = np.load(os.path.join(golden_dir, npz))
    P, y = g['P'][:64], g['y'][:64]
    ref = RefLib(name)
    ora = Oracle(mech)
    assert np.array_equal(ref.eval_jacob(P, y, 1), ora.eval_jacob(P, y, 1))
    assert np.array_equal(ref.dydt(P, y, 1), ora.dydt(P, y, 1))


@pytest.mark.parametrize('mech_file,npz', [('h2o2_n2.inp', 'h2o2_conv.npz'), ('gri30_syn.inp', 'gri30_conv.npz')])
def test_oracle_constant_volume_dydt_matches_reference(golden_dir, mech_file, npz):
    """Constant-volume dydt (rate_subs.py:2340-2485) against the reference's own CONV branch (the emitted copy
    with its two syntax slips repaired, oracle/build_ref.py conv=True)."""
    mech = Mechanism.from_chemkin(os.path.join(golden_dir, mech_file))
    g = np.load(os.path.join(golden_dir, npz))
    dy = Oracle(mech).dydt_conv(g['rho'], g['y'])
    assert _close(dy, g['dydt'], npz) > 0.999


def test_oracle_matches_reference_on_the_n_heptane_sized_mechanism(golden_dir, tmp_path):
    """654 species / 2827 reactions through the reference's own generator (tests/golden/make_golden.py nc7): rates and
    dydt in full, the Jacobian on the stored sample (all rows of 95 columns, all columns of 32 rows), bitwise."""
    from pyjac_b200 import synth
    path = str(tmp_path / 'nc7.inp')
    synth.write('nc7', path, seed=0)
    mech = Mechanism.from_chemkin(path)
    g = np.load(os.path.join(golden_dir, 'nc7_syn.npz'))
    ora = Oracle(mech)
    nsp = mech.NSP
    assert (ora.NSP, ora.NR) == (654, 2827) and g['y'].shape == (4, nsp)
    for key, arr in zip(KEYS, ora.rates(g['P'], g['y'])):
        assert np.array_equal(arr, g[key]), key
    assert np.array_equal(ora.dydt(g['P'], g['y']), g['dydt'])
    jac = ora.eval_jacob(g['P'], g['y']).reshape(4, nsp, nsp)               # [state, col, row]
    assert np.array_equal(jac[:, g['cols'], :], g['jac_cols'])
    assert np.array_equal(jac[:, :, g['rows']], g['jac_rows'])
    assert np.array_equal(np.abs(jac).max(axis=2), g['jac_colmax'])
