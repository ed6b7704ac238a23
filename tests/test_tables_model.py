"""The product's mechanism tables + the NumPy model of the kernel algorithm (tests/kernel_model.py)
reproduce the reference's generated C (golden vectors) within the parity gates -- on the CPU,
so table/algebra mistakes are caught without a GPU."""
import os

import numpy as np
import pytest

import gates
import kernel_model
from pyjac_b200 import blob, tables
from pyjac_b200.mechanism import Mechanism

CASES = [('h2o2_n2.inp', 'h2o2_pasr.npz', slice(None, None, 5)),
         ('torture.inp', 'torture_pasr.npz', slice(None)),
         ('gri30_syn.inp', 'gri30_syn.npz', slice(None)),
         ('usc2_syn.inp', 'usc2_syn.npz', slice(None)),
         ('plog.inp', 'plog_syn.npz', slice(None)),
         ('cheb.inp', 'cheb_syn.npz', slice(None)),
         ('nega.inp', 'nega_pasr.npz', slice(None)),
         ('mini.inp', 'mini_syn.npz', slice(None))]


@pytest.mark.parametrize('mech_file,npz,sl', CASES)
def test_model_matches_reference_golden(golden_dir, mech_file, npz, sl):
    mech = Mechanism.from_chemkin(os.path.join(golden_dir, mech_file))
    T = blob.unpack(blob.pack(tables.build(mech)))       # through the container, as the GPU sees it
    g = {k: v[sl] for k, v in np.load(os.path.join(golden_dir, npz)).items()}
    out = kernel_model.evaluate(T, g['P'], g['y'])
    gates.check_rates(mech, g['P'], g['y'], out, g, mech_file)
    worst, frac = gates.check_jac(out['jac'], g['jac'], mech.NSP, mech_file, mech, g['y'])
    gates.check_case(mech_file, worst, frac)


@pytest.mark.parametrize('mech_file,npz,sl', [c for c in CASES if c[0] != 'usc2_syn.inp'])
def test_stream_model_matches_reference_golden(golden_dir, mech_file, npz, sl):
    """The record streams of k_jac6 (plan6.py), interpreted record by record."""
    mech = Mechanism.from_chemkin(os.path.join(golden_dir, mech_file))
    T = blob.unpack(blob.pack(tables.build(mech, streams=True)))
    assert 'p6_str' in T
    g = {k: v[sl] for k, v in np.load(os.path.join(golden_dir, npz)).items()}
    out = kernel_model.evaluate(T, g['P'], g['y'], plan=6)
    worst, frac = gates.check_jac(out['jac'], g['jac'], mech.NSP, mech_file, mech, g['y'])
    gates.check_case(mech_file, worst, frac)
    ref = kernel_model.evaluate(T, g['P'], g['y'])
    assert np.abs(out['jac'] - ref['jac']).max() <= 1e-12 * np.abs(ref['jac']).max()


@pytest.mark.parametrize('mech_file,npz,gs,threads', [('h2o2_n2.inp', 'h2o2_pasr.npz', 4, 128), ('h2o2_n2.inp', 'h2o2_pasr.npz', 8, 512),
                                                      ('torture.inp', 'torture_pasr.npz', 16, 64), ('gri30_syn.inp', 'gri30_syn.npz', 4, 256),
                                                      ('gri30_syn.inp', 'gri30_syn.npz', 8, 256)])
def test_stream_model_other_shapes(golden_dir, mech_file, npz, gs, threads):
    """Stream plans for other states-per-block / block sizes than the automatic choice."""
    mech = Mechanism.from_chemkin(os.path.join(golden_dir, mech_file))
    T = tables.build(mech, gs=gs, threads=threads, streams=True)
    assert 'p6_str' in T and tuple(T['p6_cfg'][:2]) == (gs, threads)
    g = {k: v[:16] for k, v in np.load(os.path.join(golden_dir, npz)).items()}
    out = kernel_model.evaluate(T, g['P'], g['y'], plan=6)
    gates.check_jac(out['jac'], g['jac'], mech.NSP, mech_file, mech, g['y'])


def test_streams_can_be_left_out(golden_dir):
    mech = Mechanism.from_chemkin(os.path.join(golden_dir, 'h2o2_n2.inp'))
    assert 'p6_str' not in tables.build(mech) and 'p6_str' in tables.build(mech, streams=True)
    big = Mechanism.from_chemkin(os.path.join(golden_dir, 'usc2_syn.inp'))
    assert 'p6_str' not in tables.build(big, streams=True)          # working set in global memory: schedule tables only


def test_tables_reject_unsupported(golden_dir):
    mech = Mechanism.from_chemkin(os.path.join(golden_dir, 'h2o2_n2.inp'))
    mech.reacs[3].reac_nu = [1.5 for _ in mech.reacs[3].reac_nu]
    with pytest.raises(tables.UnsupportedMechanism):
        tables.build(mech)


def test_table_structure(golden_dir):
    mech = Mechanism.from_chemkin(os.path.join(golden_dir, 'gri30_syn.inp'))
    T = tables.build(mech)
    nsp, nr, nrev, npd, nraw = (int(v) for v in T['dims'][:5])
    assert (nsp, nr, nrev, npd) == (mech.NSP, mech.FWD_RATES, mech.REV_RATES, mech.PRES_MOD_RATES)
    assert sorted(T['rx_orig']) == list(range(nr))
    # pressure-modified reactions are the tail of the kernel order
    fl = T['rx_flags']
    pm = (fl & (tables.F_THD | tables.F_PDEP)) != 0
    assert not pm[:int(T['dims'][8])].any() and pm[int(T['dims'][8]):].all()
    # every raw row is written by exactly one reaction slot / collider / pres_mod_temp value
    dst = T['rx_dst'].reshape(nr, 8)
    rows = sorted(int(v) for v in dst[:, :6].ravel() if v != 0xFFFF)
    assert len(set(rows)) == len(rows) and (not rows or rows[-1] < nraw)
    assert T['red_off'][-1] == len(T['red_rx']) == len(T['red_nu'])


@pytest.mark.parametrize('mech_file,npz,sl', CASES)
def test_factored_record_expands_to_the_jacobian(golden_dir, mech_file, npz, sl):
    """the record k_eval<.., M_FACT> writes (energy row, T column, rank-2 factors, sparse block in a
    fixed pattern) expands to the dense Jacobian, and J v computed from the record equals (dense J) v."""
    from pyjac_b200 import factored
    mech = Mechanism.from_chemkin(os.path.join(golden_dir, mech_file))
    T = blob.unpack(blob.pack(tables.build(mech)))
    g = {k: v[sl][:32] for k, v in np.load(os.path.join(golden_dir, npz)).items()}
    out = kernel_model.evaluate(T, g['P'], g['y'])
    nsp = mech.NSP
    rows, cols, ca, cb = factored.pattern_from_tables(T)
    assert len(rows) == int(T['p5_cfg'][15]) and out['fac'].shape[1] == nsp + 3 * (nsp - 1) + len(rows)
    assert len(set(zip(rows.tolist(), cols.tolist()))) == len(rows) and rows.min() >= 1 and cols.min() >= 1
    dense = factored.expand(out['fac'], nsp, rows, cols, ca, cb)
    scale = np.abs(out['jac']).max(axis=1, keepdims=True)
    assert (np.abs(dense - out['jac']) <= 4e-16 * scale).all()
    gates.check_jac(dense, g['jac'], nsp, mech_file, mech, g['y'])
    rng = np.random.default_rng(11)
    v = rng.standard_normal((dense.shape[0], nsp))
    ref = np.einsum('ncr,nc->nr', g['jac'].reshape(-1, nsp, nsp), v)          # golden J (column-major) times v
    got = kernel_model.factored_jvp(out['fac'], v, nsp, rows, cols, ca, cb)
    # J v sums NSP products whose magnitudes differ by many decades: compare against the sum of magnitudes
    mag = np.einsum('ncr,nc->nr', np.abs(g['jac'].reshape(-1, nsp, nsp)), np.abs(v))
    assert (np.abs(got - ref) <= 1e-9 * mag + 1e-300).all()
    # the record is what makes the host round trip cheaper (the synthetic mechanisms draw their species at random, so
    # their sparse blocks are denser than those of real mechanisms of the same size)
    assert out['fac'].shape[1] < 0.75 * nsp * nsp or nsp <= 10
