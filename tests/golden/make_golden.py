#!/usr/bin/env python
"""Generate the committed golden vectors from the REAL reference (run in the build
container only -- needs /root/reference).

For each mechanism it runs the reference's own generator + gcc (oracle/build_ref.py),
evaluates the emitted C on fixed states and stores inputs and outputs:

    tests/golden/h2o2_pasr.npz     H2/O2 10-species mechanism, all 1020 bundled PaSR states
                                   (data/h2_pasr_output.npy, renormalised as
                                   functional_tester/test.py:1254-1258)
    tests/golden/torture_pasr.npz  branch-coverage mechanism, every 4th of those states
    tests/golden/gri30_syn.npz     GRI-3.0-shaped synthetic mechanism, 48 synthetic states
    tests/golden/usc2_syn.npz      USC-Mech-II-shaped synthetic mechanism (111 sp / 784 rxn, species with
                                   different T_mid), 64 synthetic states

    tests/golden/plog_syn.npz      PLOG coverage mechanism over the H2/O2 species, 192 synthetic states of
                                   which 32 below and 32 above every pressure table
    tests/golden/cheb_syn.npz      Chebyshev coverage mechanism over the H2/O2 species, 160 synthetic states of
                                   which 32 outside the fitted pressure ranges

    tests/golden/mini_syn.npz      the mechanism that also exists as mini.cti / mini.yaml (third body, Troe, SRI, Lindemann with a
                                   specific collider, chemically activated, PLOG, Chebyshev, duplicates), 96 synthetic states

    tests/golden/nc7_syn.npz       n-heptane-sized synthetic mechanism (654 sp / 2827 rxn; the file is synth.write('nc7', seed=0)),
                                   4 synthetic states; rates and dydt in full, of the Jacobian all rows of 96 columns and
                                   all columns of 32 rows (+ the per-column maxima the gate scales with)

    tests/golden/nega_pasr.npz     negative pre-exponential factors (duplicate pairs, every A < 0 branch of rs:108-141),
                                   every 4th PaSR state

    tests/golden/h2o2_conv.npz     constant-volume dydt (the reference's CONV branch, its two syntax slips repaired in
                                   the emitted copy: oracle/build_ref.py conv=True) on every 4th PaSR state:
                                   rho (the density of the state at its pressure), y, dydt
    tests/golden/gri30_conv.npz    the same for the GRI-3.0-shaped mechanism, 48 synthetic states

Arrays are in pyJac's internal (moved-last) species order, row-major per state:
P[n], y[n,NSP] = [T, Y_0..Y_{NSP-2}], conc, fwd, rev, pres_mod, spec_rates, dydt, jac[n,NSP*NSP]
(column-major inside a state).
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import build_ref                                   # noqa: E402
from oracle.oracle import RefLib                               # noqa: E402
from pyjac_b200 import synth                                   # noqa: E402
from pyjac_b200.mechanism import Mechanism                     # noqa: E402
from pyjac_b200.states import pasr_states, synthetic_states    # noqa: E402


def dump_conv(name, mech_file, P, y, out_name):
    """Constant-volume dydt of the reference at the densities the states have at their pressures."""
    import ctypes
    build_ref.build(name, mech_file)
    conc = RefLib(name).rates(P, y)[0]
    w = np.array([sp.mw for sp in Mechanism.from_chemkin(mech_file).specs])
    rho = np.ascontiguousarray((conc * w[None, :]).sum(axis=1))
    L = ctypes.CDLL(build_ref.build(name + '_conv', mech_file, conv=True))
    dp = ctypes.POINTER(ctypes.c_double)
    L.ref_dydt_batch.argtypes = [ctypes.c_int, dp, dp, dp, ctypes.c_int]
    y = np.ascontiguousarray(y)
    dy = np.zeros_like(y)
    L.ref_dydt_batch(len(rho), rho.ctypes.data_as(dp), y.ctypes.data_as(dp), dy.ctypes.data_as(dp), 1)
    np.savez_compressed(os.path.join(HERE, out_name), rho=rho, y=y, dydt=dy)
    print(name + '_conv', y.shape, 'dT/dt in [%.3g, %.3g]' % (dy[:, 0].min(), dy[:, 0].max()))


def dump(name, mech_file, P, y, out_name):
    build_ref.build(name, mech_file)
    ref = RefLib(name)
    conc, fwd, rev, pm, sr = ref.rates(P, y)
    out = dict(P=P, y=y, conc=conc, fwd=fwd, rev=rev, pres_mod=pm, spec_rates=sr,
               dydt=ref.dydt(P, y, 1), jac=ref.eval_jacob(P, y, 1))
    np.savez_compressed(os.path.join(HERE, out_name), **out)
    print(name, y.shape, 'jac nnz frac %.3f' % (out['jac'] != 0).mean())


if __name__ == '__main__':
    pasr = os.path.join(HERE, 'h2_pasr_output.npy')
    if not os.path.exists(pasr):
        shutil.copy('/root/reference/data/h2_pasr_output.npy', pasr)
    only = set(sys.argv[1:])               # e.g. `make_golden.py conv` regenerates only the constant-volume vectors

    if 'mini' in only:
        mech = Mechanism.from_chemkin(os.path.join(HERE, 'mini.inp'))
        P, y = synthetic_states(mech.NSP, 96, seed=8)
        P = P.copy()
        P[64:80] *= 0.01          # below / above the PLOG table and the Chebyshev pressure range
        P[80:96] *= 20.0
        dump('mini', os.path.join(HERE, 'mini.inp'), P, y, 'mini_syn.npz')
        sys.exit(0)
    if 'nc7' in only:
        # 654 species / 2827 reactions: the generator emits 7.5 M lines of C (609 MB); built outside the repo with -O0
        # (no value-changing optimisation is enabled at -O3 -mtune=native either: the other fixtures are bit-identical
        # at both levels), and only a sample of the Jacobian is kept: all rows of 96 columns, all columns of 32 rows
        build_ref.OUT_ROOT = '/tmp/pyjac_ref_big'
        nc7 = '/tmp/pyjac_ref_big/nc7_syn.inp'
        os.makedirs(build_ref.OUT_ROOT, exist_ok=True)
        synth.write('nc7', nc7, seed=0)
        mech = Mechanism.from_chemkin(nc7)
        P, y = synthetic_states(mech.NSP, 4, seed=21)
        lib = build_ref.build('nc7', nc7, opt='-O0', jobs=8)
        ref = RefLib(lib)
        conc, fwd, rev, pm, sr = ref.rates(P, y)
        jac = ref.eval_jacob(P, y, 1).reshape(4, mech.NSP, mech.NSP)            # [state, col, row]
        rng = np.random.default_rng(0)
        cols = np.unique(np.concatenate([[0, 1, mech.NSP - 1], rng.choice(mech.NSP, 93, replace=False)]))
        rows = np.unique(np.concatenate([[0, 1, mech.NSP - 1], rng.choice(mech.NSP, 29, replace=False)]))
        np.savez_compressed(os.path.join(HERE, 'nc7_syn.npz'), P=P, y=y, conc=conc, fwd=fwd, rev=rev, pres_mod=pm,
                            spec_rates=sr, dydt=ref.dydt(P, y, 1), cols=cols, rows=rows,
                            jac_cols=jac[:, cols, :], jac_rows=jac[:, :, rows],
                            jac_colmax=np.abs(jac).max(axis=2))
        print('nc7', y.shape, 'sampled jac', jac[:, cols, :].shape, jac[:, :, rows].shape)
        sys.exit(0)
    if 'usc2' in only:
        usc = os.path.join(HERE, 'usc2_syn.inp')
        mech = Mechanism.from_chemkin(usc)
        P, y = synthetic_states(mech.NSP, 64, seed=7)
        dump('usc2', usc, P, y, 'usc2_syn.npz')
        sys.exit(0)
    if 'nega' in only:
        mech = Mechanism.from_chemkin(os.path.join(HERE, 'nega.inp'))
        P, y = pasr_states(pasr, mech)
        dump('nega', os.path.join(HERE, 'nega.inp'), P[::4], y[::4], 'nega_pasr.npz')
        sys.exit(0)

    mech = Mechanism.from_chemkin(os.path.join(HERE, 'h2o2_n2.inp'))
    P, y = pasr_states(pasr, mech)
    dump_conv('h2o2', os.path.join(HERE, 'h2o2_n2.inp'), P[::4], y[::4], 'h2o2_conv.npz')
    mech = Mechanism.from_chemkin(os.path.join(HERE, 'gri30_syn.inp'))
    P, y = synthetic_states(mech.NSP, 48, seed=0)
    dump_conv('gri30', os.path.join(HERE, 'gri30_syn.inp'), P, y, 'gri30_conv.npz')
    if only == {'conv'}:
        sys.exit(0)

    mech = Mechanism.from_chemkin(os.path.join(HERE, 'h2o2_n2.inp'))
    P, y = pasr_states(pasr, mech)
    dump('h2o2', os.path.join(HERE, 'h2o2_n2.inp'), P, y, 'h2o2_pasr.npz')

    mech = Mechanism.from_chemkin(os.path.join(HERE, 'torture.inp'))
    P, y = pasr_states(pasr, mech)
    dump('torture', os.path.join(HERE, 'torture.inp'), P[::4], y[::4], 'torture_pasr.npz')

    mech = Mechanism.from_chemkin(os.path.join(HERE, 'nega.inp'))
    P, y = pasr_states(pasr, mech)
    dump('nega', os.path.join(HERE, 'nega.inp'), P[::4], y[::4], 'nega_pasr.npz')

    gri = os.path.join(HERE, 'gri30_syn.inp')
    synth.write('gri30', gri, seed=0)
    mech = Mechanism.from_chemkin(gri)
    P, y = synthetic_states(mech.NSP, 48, seed=0)
    dump('gri30', gri, P, y, 'gri30_syn.npz')

    usc = os.path.join(HERE, 'usc2_syn.inp')
    synth.write('usc2', usc, seed=0)
    mech = Mechanism.from_chemkin(usc)
    P, y = synthetic_states(mech.NSP, 64, seed=7)
    dump('usc2', usc, P, y, 'usc2_syn.npz')

    plog = os.path.join(HERE, 'plog.inp')
    mech = Mechanism.from_chemkin(plog)
    P, y = synthetic_states(mech.NSP, 192, seed=3)
    P = P.copy()
    P[128:160] *= 0.01          # 0.005 - 0.25 atm: below the first pressure of most tables
    P[160:192] *= 20.0          # 10 - 500 atm: above the last pressure
    dump('plog', plog, P, y, 'plog_syn.npz')

    cheb = os.path.join(HERE, 'cheb.inp')
    mech = Mechanism.from_chemkin(cheb)
    P, y = synthetic_states(mech.NSP, 160, seed=4)
    P = P.copy()
    P[128:144] *= 0.01
    P[144:160] *= 20.0
    dump('cheb', cheb, P, y, 'cheb_syn.npz')
