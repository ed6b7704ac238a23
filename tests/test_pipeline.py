"""create_jacobian / generate_library / generate_wrapper keep the reference's pipeline shape:
the build directory holds a mechanism.h that the reference's own regex consumers can read
(functional_tester/test.py:311-318,358; libgen/libgen.py:385) plus the table blob."""
import os
import re

import numpy as np
import pytest

from pyjac_b200 import blob, libgen
from pyjac_b200.create_jacobian import create_jacobian, load_tables


def test_create_jacobian_writes_header_and_tables(golden_dir, tmp_path):
    out = str(tmp_path / 'out')
    mech = create_jacobian('cuda', os.path.join(golden_dir, 'h2o2_n2.inp'), build_path=out)
    lines = open(os.path.join(out, 'mechanism.h')).read().splitlines()
    found = {}
    for line in lines:
        for key in ('NSP', 'NN', 'FWD_RATES', 'REV_RATES'):
            m = re.search(r'^#define %s (\d+)$' % key, line)
            if m:
                found[key] = int(m.group(1))
        m = re.search(r'\s*#define PRES_MOD_RATES (\d+)', line.strip())
        if m:
            found['PRES_MOD_RATES'] = int(m.group(1))
        m = re.search(r'^//last_spec (\d+)$', line)
        if m:
            found['last_spec'] = int(m.group(1))
    assert found == {'NSP': 10, 'NN': 11, 'FWD_RATES': 28, 'REV_RATES': 28, 'PRES_MOD_RATES': 6,
                     'last_spec': mech.last_spec_original}
    T = blob.unpack(load_tables(out))
    assert int(T['dims'][0]) == 10 and 'p5_cfg' in T
    # what a program written against the emitted library includes (tester.c.in, read_initial_conditions.c, the .pyx files)
    for name in ('header.h', 'jacob.h', 'dydt.h', 'rates.h', 'chem_utils.h', 'pyjacob.cuh'):
        assert os.path.isfile(os.path.join(out, name)), name
    assert 'void apply_mask(double*);' in open(os.path.join(out, 'mechanism.h')).read()


def test_generate_library_builds_the_mechanism_stub(golden_dir, tmp_path):
    """generate_library: a per-mechanism library under the reference's name (libgen.py:170-186) that embeds
    the tables, registers them with the CUDA library when loaded and forwards the emitted library's
    entry points -- what `-lc_pyjac` links.  Loading it needs no GPU."""
    import ctypes
    import subprocess
    out = str(tmp_path / 'out')
    create_jacobian('cuda', os.path.join(golden_dir, 'h2o2_n2.inp'), build_path=out)
    for lang, name in (('c', 'libc_pyjac.so'), ('cuda', 'libcu_pyjac.so')):
        path = libgen.generate_library(lang, out)
        assert os.path.basename(path) == name and os.path.dirname(path) == out
        syms = subprocess.run(['nm', '-D', '--defined-only', path], capture_output=True, text=True).stdout
        for fn in ('eval_jacob', 'dydt', 'eval_conc', 'eval_rxn_rates', 'get_rxn_pres_mod', 'eval_spec_rates',
                   'eval_h', 'eval_u', 'eval_cv', 'eval_cp', 'apply_mask', 'apply_reverse_mask'):
            assert re.search(r' T %s$' % fn, syms, flags=re.M), fn
        needed = subprocess.run(['readelf', '-d', path], capture_output=True, text=True).stdout
        assert 'libpyjac_b200.so' in needed
        ctypes.CDLL(path)                     # the constructor registers the tables: host only
    with pytest.raises(ValueError):
        libgen.generate_library('fortran', out)


def test_generate_wrapper_writes_importable_modules(golden_dir, tmp_path):
    """generate_wrapper writes `pyjacob` / `cu_pyjacob` modules that the reference's callers import by name
    (functional_tester/test.py:432,740).  Importing them needs a GPU (GPU test); here: they exist and name
    the reference's functions."""
    from pyjac_b200 import lib, pywrap
    out = str(tmp_path / 'out')
    create_jacobian('cuda', os.path.join(golden_dir, 'h2o2_n2.inp'), build_path=out)
    mods = str(tmp_path / 'mods')
    for lang, name, fns in (('c', 'pyjacob', pywrap._PYJACOB), ('cuda', 'cu_pyjacob', pywrap._CU_PYJACOB)):
        try:
            pywrap.generate_wrapper(lang, out, mods)
        except lib.PyjacError:
            pass                              # no device here: the module is written before the mechanism is loaded
        text = open(os.path.join(mods, name + '.py')).read()
        for fn in fns:
            assert '%s = _mod.%s' % (fn, fn) in text


def test_create_jacobian_rejects_what_it_cannot_do(golden_dir, tmp_path):
    mech = os.path.join(golden_dir, 'h2o2_n2.inp')
    with pytest.raises(ValueError):
        create_jacobian('c', mech, build_path=str(tmp_path))          # no CPU back end
    with pytest.raises(NotImplementedError):
        create_jacobian('cuda', mech, build_path=str(tmp_path), auto_diff=True)
    with pytest.raises(FileNotFoundError):
        libgen.generate_library('cuda', str(tmp_path))


def test_plan_tables_are_consistent(golden_dir, tmp_path):
    """Every species, reaction and Jacobian element is scheduled exactly once for several
    (states per block, block size) choices."""
    from pyjac_b200 import tables
    from pyjac_b200.mechanism import Mechanism
    mech = Mechanism.from_chemkin(os.path.join(golden_dir, 'gri30_syn.inp'))
    for gs, nt in ((8, 512), (4, 256), (2, 128), (8, 64), (32, 512)):
        if gs == 32:
            mech_ = Mechanism.from_chemkin(os.path.join(golden_dir, 'h2o2_n2.inp'))
        else:
            mech_ = mech
        T = tables.build(mech_, gs=gs, threads=nt)
        nsp, nr = int(T['dims'][0]), int(T['dims'][1])
        nw, nsub = nt // 32, 64 // gs
        items = T['p5_b_item'][:int(T['p5_b_off'][nw]) * nsub]
        assert sorted(items[items >= 0]) == list(range(nr))
        c_item = T['p5_c_item'].reshape(-1, 4)[:int(T['p5_c_off'][nw])]
        c_str = T['p5_c_str'].view(np.uint32).reshape(-1, nsub, 2)
        heads = [int(c_str[int(u), sb, 0]) for u in c_item[:, 0] for sb in range(nsub) if c_str[int(u), sb, 1]]
        assert sorted(hd // (gs * 8 * 8) for hd in heads) == list(range(nsp))
        d_str = T['p5_d_str'].view(np.uint32).reshape(-1, nsub, 2)
        d_elems = []
        for u, ncol in T['p5_d_item'].reshape(-1, 2)[:int(T['p5_d_off'][nw])]:
            for sb in range(nsub):
                if d_str[u, sb, 0] != 0xFFFFFFFF:
                    if d_str[u, sb, 1] != 0x3FFFFF:
                        d_elems.append(int(d_str[u, sb, 1]))
                    d_elems += [int(e) for e in d_str[u + 1:u + 1 + ncol, sb, 0] if e != 0x3FFFFF]
        s = T['p5_s_str'].view(np.uint32).reshape(-1, nsub, 4)[:int(T['p5_s_off'][nw]) * 2][0::2, :, 0].ravel() & 0x3FFFFF
        elems = sorted(d_elems + [int(e) for e in s if e != 0x3FFFFF])
        want = sorted(c * nsp + r for c in range(nsp) for r in range(1, nsp))
        assert elems == want
        t_str = T['p5_t_str'].view(np.uint32).reshape(-1, nsub, 2)
        t_cols = [int(t_str[u, sb, 0]) & 0xFFFF for u, _ in T['p5_t_item'].reshape(-1, 2)[:int(T['p5_t_off'][nw])]
                  for sb in range(nsub) if t_str[u, sb, 0] >> 16]
        assert sorted(t_cols) == list(range(1, nsp))
        assert int(T['p5_cfg'][9]) * 8 <= 232448


def test_data_bin_round_trip(golden_dir, tmp_path):
    """data.bin rows [t, T, P, Y...] in the original species order come back masked and
    state-fastest, the way read_initial_conditions.cu:10-59 hands them to the device."""
    from pyjac_b200 import speedtest
    from pyjac_b200.states import pasr_states
    out = str(tmp_path / 'out')
    mech = create_jacobian('cuda', os.path.join(golden_dir, 'h2o2_n2.inp'), build_path=out, skip_jac=True)
    raw = np.load(os.path.join(golden_dir, 'h2_pasr_output.npy'))
    raw = raw.reshape(-1, raw.shape[-1])[:50]
    data = str(tmp_path / 'data.bin')
    speedtest.write_data_bin(data, raw[:, 1], raw[:, 2], raw[:, 3:], t=raw[:, 0])
    assert os.path.getsize(data) == 50 * (mech.NSP + 3) * 8
    fwd = speedtest._fwd_spec_map(out, mech.NSP)
    assert fwd == mech.fwd_spec_map
    y, pres = speedtest.read_initial_conditions(data, 50, fwd)
    assert y.shape == (mech.NSP, 50) and np.array_equal(pres, raw[:, 2]) and np.array_equal(y[0], raw[:, 1])
    assert np.array_equal(y[1:], raw[:, 3:][:, fwd][:, :-1].T)


def test_plan_choice_by_mechanism_size(golden_dir, tmp_path):
    """Where the working set of a block lives (plan.py: SMEM_MIN_GS, wsg_shape): shared memory while
    at least 8 states fit, else global memory; both can be forced."""
    from pyjac_b200 import synth, tables
    from pyjac_b200.mechanism import Mechanism
    cases = [(os.path.join(golden_dir, 'h2o2_n2.inp'), 32, 0), (os.path.join(golden_dir, 'gri30_syn.inp'), 8, 0),
             (os.path.join(golden_dir, 'usc2_syn.inp'), 8, 1)]
    nc7 = str(tmp_path / 'nc7.inp')
    synth.write('nc7', nc7)
    cases.append((nc7, 16, 1))
    for path, gs, wsg in cases:
        cfg = tables.build(Mechanism.from_chemkin(path))['p5_cfg']
        assert (int(cfg[0]), int(cfg[14])) == (gs, wsg), path
    usc = Mechanism.from_chemkin(os.path.join(golden_dir, 'usc2_syn.inp'))
    cfg = tables.build(usc, ws_global=False)['p5_cfg']
    assert (int(cfg[0]), int(cfg[14])) == (2, 0)
    cfg = tables.build(Mechanism.from_chemkin(os.path.join(golden_dir, 'gri30_syn.inp')), ws_global=True)['p5_cfg']
    assert (int(cfg[0]), int(cfg[14])) == (8, 1)
    with pytest.raises(tables.UnsupportedMechanism):
        tables.build(Mechanism.from_chemkin(nc7), ws_global=False)


def _build_in_child(build_dir, counter, q):
    """Child of test_library_build_is_safe_under_concurrent_ranks: libgen redirected to a scratch directory and a
    fake compiler that takes a while and counts its runs."""
    import stat
    from pyjac_b200 import libgen
    fake = os.path.join(build_dir, 'fake_nvcc.sh')
    if not os.path.exists(fake):
        with open(fake + '.tmp%d' % os.getpid(), 'w') as fh:
            fh.write('#!/bin/bash\nout=""\nwhile [ $# -gt 0 ]; do if [ "$1" = "-o" ]; then out="$2"; fi; shift; done\n'
                     'echo run >> %s\nsleep 0.5\nprintf "x%%.0s" $(seq 1 200000) > "$out"\n' % counter)
        os.chmod(fake + '.tmp%d' % os.getpid(), stat.S_IRWXU)
        os.replace(fake + '.tmp%d' % os.getpid(), fake)
    libgen.BUILD = build_dir
    libgen.LIB_PATH = os.path.join(build_dir, libgen.LIB_NAME)
    libgen._nvcc = lambda: fake
    path = libgen.build_library()
    q.put(os.path.getsize(path))


def test_library_build_is_safe_under_concurrent_ranks(tmp_path):
    """One process per GPU may find the in-tree library stale at the same moment (torchrun): the build must run once,
    under a lock, and nobody may see a half-written file (this corrupted an 8-rank bench run once)."""
    import multiprocessing as mp
    ctx = mp.get_context('spawn')
    build_dir, counter = str(tmp_path / 'b'), str(tmp_path / 'runs')
    os.makedirs(build_dir)
    q = ctx.Queue()
    procs = [ctx.Process(target=_build_in_child, args=(build_dir, counter, q)) for _ in range(4)]
    for p in procs:
        p.start()
    sizes = [q.get(timeout=60) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert sizes == [200000] * 4                                   # everybody loaded the complete file
    assert open(counter).read().count('run') == 1                 # ... which was built once
    assert not [f for f in os.listdir(build_dir) if '.tmp' in f]
