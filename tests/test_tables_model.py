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
         ('gri30_syn.inp', 'gri30_syn.npz', slice(None))]


@pytest.mark.parametrize('mech_file,npz,sl', CASES)
def test_model_matches_reference_golden(golden_dir, mech_file, npz, sl):
    mech = Mechanism.from_chemkin(os.path.join(golden_dir, mech_file))
    T = blob.unpack(blob.pack(tables.build(mech)))       # through the container, as the GPU sees it
    g = {k: v[sl] for k, v in np.load(os.path.join(golden_dir, npz)).items()}
    out = kernel_model.evaluate(T, g['P'], g['y'])
    gates.check_rates(mech, g['P'], g['y'], out, g, mech_file)
    worst, frac = gates.check_jac(out['jac'], g['jac'], mech.NSP, mech_file)
    assert frac > 0.97


def test_tables_reject_unsupported(golden_dir):
    mech = Mechanism.from_chemkin(os.path.join(golden_dir, 'h2o2_n2.inp'))
    mech.reacs[3].reac_nu = [1.5 for _ in mech.reacs[3].reac_nu]
    with pytest.raises(tables.UnsupportedMechanism):
        tables.build(mech)


def test_table_structure(golden_dir):
    mech = Mechanism.from_chemkin(os.path.join(golden_dir, 'gri30_syn.inp'))
    T = tables.build(mech)
    nsp, nr, nrev, npd, nraw, nnz, ncon = (int(v) for v in T['dims'][:7])
    assert (nsp, nr, nrev, npd) == (mech.NSP, mech.FWD_RATES, mech.REV_RATES, mech.PRES_MOD_RATES)
    assert sorted(T['rx_orig']) == list(range(nr))
    # pressure-modified reactions are the tail of the kernel order
    fl = T['rx_flags']
    pm = (fl & (tables.F_THD | tables.F_PDEP)) != 0
    assert not pm[:int(T['dims'][8])].any() and pm[int(T['dims'][8]):].all()
    # sparse entries sorted by decreasing work, every contribution points at a raw slot
    cnt = np.diff(T['ent_off'])
    assert (cnt[:-1] >= cnt[1:]).all() and T['ent_off'][-1] == ncon
    assert ((T['con'] & 0xFFFF) < nraw).all() and ((T['con'] >> 16) < int(T['dims'][7])).all()
    assert T['jmap'].max() == nnz and (np.sort(T['jmap'][T['jmap'] < nnz]) == np.arange(nnz)).all()
