"""GPU tests of the drop-in boundary: the reference-named entry points that round 1 lacked --
eval_h / eval_u / eval_cv / eval_cp, apply_mask, the C++-linkage init / run / cleanup of pyjacob.cuh --, the
re-entrancy of the scalar API, the importable pyjacob / cu_pyjacob modules, and the reference's own
performance harness (tester.c.in + read_initial_conditions.c + timer.h, built by oracle/build_ref.py where
those sources lie) linked against this library."""
import ctypes
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

import gates
from pyjac_b200.mechanism import Mechanism

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def torch():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch


def _build(golden_dir, tmp_path, mech_file='h2o2_n2.inp'):
    from pyjac_b200.create_jacobian import create_jacobian
    out = str(tmp_path / 'out')
    mech = create_jacobian('cuda', os.path.join(golden_dir, mech_file), build_path=out)
    return mech, out


def _ref_lib(name):
    path = os.path.join(ROOT, 'oracle', '_ref', name, 'libc_pyjac.so')
    if not os.path.exists(path):
        pytest.skip('oracle/_ref/%s is not built (run __graft_entry__.build() where /root/reference exists)' % name)
    return ctypes.CDLL(path)


@pytest.mark.parametrize('mech_file,ref', [('h2o2_n2.inp', 'h2o2'), ('gri30_syn.inp', 'gri30')])
def test_thermo_entry_points_vs_reference(torch, golden_dir, tmp_path, mech_file, ref):
    """eval_h / eval_u / eval_cv / eval_cp (rate_subs.py:1806-2086) against the reference's own generated C."""
    from pyjac_b200.pywrap import generate_wrapper
    mech, out = _build(golden_dir, tmp_path, mech_file)
    mod = generate_wrapper('c', out, str(tmp_path / 'mods'))
    R = _ref_lib(ref)
    nsp = mech.NSP
    temps = [300.0, 999.9999, 1000.0, 1000.0001, 1537.25, 3500.0] + [sp.Trange[1] for sp in mech.specs[:3]]
    for which in ('h', 'u', 'cv', 'cp'):
        fn = getattr(R, 'eval_' + which)
        fn.argtypes = [ctypes.c_double, ctypes.c_void_p]
        fn.restype = None
        for T in temps:
            want = np.empty(nsp)
            fn(T, want.ctypes.data)
            got = mod.eval_thermo(which, T)
            assert np.all(np.abs(got - want) <= 1e-13 * np.abs(want) + 1e-300), (which, T, np.abs(got / want - 1).max())
    mod.close()


def test_apply_mask_round_trip(torch, golden_dir, tmp_path):
    """apply_mask moves the last species to the end, apply_reverse_mask undoes it (mech_auxiliary.py:188-206);
    checked against the reference's generated functions."""
    from pyjac_b200 import lib
    from pyjac_b200.pywrap import generate_wrapper
    mech, out = _build(golden_dir, tmp_path)
    mod = generate_wrapper('c', out, str(tmp_path / 'mods'))
    mod._select()
    L = lib.load()
    R = _ref_lib('h2o2')
    y = np.arange(1.0, mech.NSP + 1.0)
    a, b = y.copy(), y.copy()
    L.apply_mask(a.ctypes.data)
    R.apply_mask.argtypes = [ctypes.c_void_p]
    R.apply_mask(b.ctypes.data)
    assert np.array_equal(a, b) and np.array_equal(a, y[mech.fwd_spec_map])
    L.apply_reverse_mask(a.ctypes.data)
    assert np.array_equal(a, y)
    mod.close()


def test_scalar_api_is_reentrant(torch, golden_dir, tmp_path):
    """eval_jacob / dydt from eight host threads at once (the reference's harness calls eval_jacob inside
    an OpenMP loop, tester.c.in:24-29): every result equals the single-threaded one bitwise."""
    from pyjac_b200.pywrap import generate_wrapper
    mech, out = _build(golden_dir, tmp_path)
    mod = generate_wrapper('c', out, str(tmp_path / 'mods'))
    g = dict(np.load(os.path.join(golden_dir, 'h2o2_pasr.npz')))
    nsp, n = mech.NSP, 160
    ref_j, ref_d = np.zeros((n, nsp * nsp)), np.zeros((n, nsp))
    for s in range(n):
        mod.py_eval_jacobian(0.0, float(g['P'][s]), np.ascontiguousarray(g['y'][s]), ref_j[s])
        mod.py_dydt(0.0, float(g['P'][s]), np.ascontiguousarray(g['y'][s]), ref_d[s])
    gates.check_jac(ref_j, g['jac'][:n], nsp, 'scalar API', mech, g['y'][:n])
    out_j, out_d = np.zeros_like(ref_j), np.zeros_like(ref_d)
    errs = []

    def work(t):
        try:
            for s in range(t, n, 8):
                mod.py_eval_jacobian(0.0, float(g['P'][s]), np.ascontiguousarray(g['y'][s]), out_j[s])
                mod.py_dydt(0.0, float(g['P'][s]), np.ascontiguousarray(g['y'][s]), out_d[s])
        except Exception as exc:           # pragma: no cover
            errs.append(exc)
    threads = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs
    assert np.array_equal(out_j, ref_j) and np.array_equal(out_d, ref_d)
    mod.close()


def test_cxx_init_run_cleanup(torch, golden_dir, tmp_path):
    """int init(int) / void run(...) / void cleanup() with the C++ linkage of pyjac/pywrap/pyjacob.cuh:6-10,
    called the way pyjacob_cuda_wrapper.pyx:13-34 does."""
    from pyjac_b200 import lib
    from pyjac_b200.pywrap import generate_wrapper
    mech, out = _build(golden_dir, tmp_path)
    mod = generate_wrapper('cuda', out, str(tmp_path / 'mods'))
    mod._select()
    L = lib.load()
    g = dict(np.load(os.path.join(golden_dir, 'h2o2_pasr.npz')))
    num, nsp = 333, mech.NSP
    padded = getattr(L, '_Z4initi')(num)
    assert padded >= num
    y = np.ascontiguousarray(g['y'][:num].T).ravel()
    dy, jac = np.zeros(num * nsp), np.zeros(num * nsp * nsp)
    getattr(L, '_Z3runiiPKdS0_PdS1_S1_S1_S1_S1_S1_')(num, padded, np.ascontiguousarray(g['P'][:num]).ctypes.data, y.ctypes.data,
                                                   None, None, None, None, None, dy.ctypes.data, jac.ctypes.data)
    getattr(L, '_Z7cleanupv')()
    sub = {k: v[:num] for k, v in g.items()}
    gates.check_dydt(mech, sub['y'], dy.reshape(nsp, num).T, sub, 'run()')
    gates.check_jac(np.ascontiguousarray(jac.reshape(nsp * nsp, num).T), sub['jac'], nsp, 'run()', mech, sub['y'])
    mod.close()


def test_modules_import_by_name(torch, golden_dir, tmp_path):
    """`import pyjacob` / `import cu_pyjacob` after generate_wrapper, as functional_tester/test.py:432,740 do."""
    from pyjac_b200.pywrap import generate_wrapper
    mech, out = _build(golden_dir, tmp_path)
    mods = str(tmp_path / 'mods')
    generate_wrapper('c', out, mods).close()
    generate_wrapper('cuda', out, mods).close()
    code = ('import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r); import pyjacob, cu_pyjacob\n'
            'g = np.load(%r); y = np.ascontiguousarray(g["y"][5]); jac = np.zeros(pyjacob.NSP ** 2)\n'
            'pyjacob.py_eval_jacobian(0.0, float(g["P"][5]), y, jac)\n'
            'print(float(np.abs(jac - g["jac"][5]).max() / np.abs(g["jac"][5]).max()))\n'
            'p = cu_pyjacob.py_cuinit(4); cu_pyjacob.py_cuclean(); print(p)\n'
            % (mods, ROOT, os.path.join(golden_dir, 'h2o2_pasr.npz')))
    res = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    err, padded = res.stdout.split()
    assert float(err) < 1e-12 and int(padded) >= 4


def test_reference_harness_links_and_runs(torch, golden_dir, tmp_path):
    """`speedtest <num_odes> <num_threads>` = the reference's tester.c.in + read_initial_conditions.c + timer.h,
    compiled unmodified against the headers create_jacobian writes and linked with -lc_pyjac
    (oracle/build_ref.py build_speedtest, prebuilt under oracle/_ref/): what performance_tester.py:500-508 runs."""
    from pyjac_b200 import speedtest
    exe = os.path.join(ROOT, 'oracle', '_ref', 'h2o2_speedtest', 'speedtest')
    if not os.path.exists(exe):
        pytest.skip('oracle/_ref/h2o2_speedtest is not built')
    mech = Mechanism.from_chemkin(os.path.join(golden_dir, 'h2o2_n2.inp'))
    g = dict(np.load(os.path.join(golden_dir, 'h2o2_pasr.npz')))
    Y_int = np.concatenate([g['y'][:, 1:], 1.0 - g['y'][:, 1:].sum(axis=1, keepdims=True)], axis=1)
    Y_orig = np.empty_like(Y_int)
    Y_orig[:, mech.fwd_spec_map] = Y_int
    speedtest.write_data_bin(str(tmp_path / 'data.bin'), g['y'][:, 0], g['P'], Y_orig)
    for threads in (1, 4):
        res = subprocess.run([exe, '1020', str(threads)], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, res.stderr
        num, ms = res.stdout.strip().split(',')
        assert int(num) == 1020 and float(ms) > 0.0
        print('reference harness on pyjac_b200: %s states, %d threads: %s ms' % (num, threads, ms))


def test_constant_volume_build(torch, golden_dir, tmp_path):
    """create_jacobian(conp=False): header.h says `#define CONV`, the reference-named dydt takes the density,
    eval_jacob refuses (no such form upstream either)."""
    from pyjac_b200 import lib
    from pyjac_b200.create_jacobian import create_jacobian
    from pyjac_b200.pywrap import generate_wrapper
    out = str(tmp_path / 'out')
    mech = create_jacobian('cuda', os.path.join(golden_dir, 'h2o2_n2.inp'), build_path=out, conp=False)
    assert '\n#define CONV' in open(os.path.join(out, 'header.h')).read()
    mod = generate_wrapper('c', out, str(tmp_path / 'mods'))
    g = dict(np.load(os.path.join(golden_dir, 'h2o2_conv.npz')))
    dy = np.zeros(mech.NSP)
    # reacting states (at near-equilibrium ones dydt is the rounding residue of cancelling rates: the gated
    # comparison of those is tests/test_gpu_parity.py::test_constant_volume_dydt_vs_reference_golden)
    for s in np.argsort(-np.abs(g['dydt'][:, 0]))[[0, 20, 60]]:
        mod.py_dydt(0.0, float(g['rho'][s]), np.ascontiguousarray(g['y'][s]), dy)
        ref = g['dydt'][s]
        assert np.abs(dy - ref).max() <= 1e-9 * np.abs(ref).max()
    mod.close()
