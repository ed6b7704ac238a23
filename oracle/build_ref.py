#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- build the *real* reference for a mechanism into oracle/_ref/.

Recipe:
  1. run the unmodified reference generator where it lies,
         PYTHONPATH=/root/reference python -m pyjac --lang c --input MECH -b oracle/_ref/<name>/src
     (pyjac/__main__.py:7-26 -> create_jacobian, create_jacobian.py:3407)
  2. compile the emitted C with the reference's own flags
         gcc -std=c99 -O3 -mtune=native            (pyjac/libgen/libgen.py:43)
     plus -fPIC -fopenmp, together with oracle/ref_batch.c (an OpenMP loop over states,
     as pyjac/performance_tester/tester.c.in:23-31), into
         oracle/_ref/<name>/libc_pyjac.so
Nothing is copied out of /root/reference; generated sources and binaries stay under
oracle/_ref/ (git-ignored, but they travel to the GPU box with the repo snapshot).

Usage:  python oracle/build_ref.py NAME MECH.inp [--last-spec SP] [--opt -O3] [--jobs 8]
"""
from __future__ import annotations

import argparse
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = '/root/reference'
OUT_ROOT = os.path.join(HERE, '_ref')


def ref_dir(name: str) -> str:
    return os.path.join(OUT_ROOT, name)


def lib_path(name: str) -> str:
    return os.path.join(ref_dir(name), 'libc_pyjac.so')


def have_reference() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, 'pyjac'))


def build(name: str, mech: str, last_spec=None, opt: str = '-O3', jobs: int = 8,
          force: bool = False, quiet: bool = True, conv: bool = False) -> str:
    """conv: the constant-volume build -- the generator hard-wires `#define CONP` into the header.h it
    emits (mech_auxiliary.py:464-466); the emitted copy under oracle/_ref/ is switched to `#define CONV`
    before compiling, which is how a user of the reference selects it.  `name` should differ from the
    constant-pressure build's."""
    out = ref_dir(name)
    src = os.path.join(out, 'src')
    lib = lib_path(name)
    stamp = os.path.join(out, 'mech.inp')
    mech_txt = open(mech).read()
    if (not force and os.path.exists(lib) and os.path.exists(stamp)
            and open(stamp).read() == mech_txt):
        return lib
    if not have_reference():
        raise RuntimeError('reference sources not present at %s' % REF_ROOT)
    os.makedirs(src, exist_ok=True)
    for f in glob.glob(os.path.join(src, '**', '*.[cho]'), recursive=True):
        os.remove(f)

    env = dict(os.environ, PYTHONPATH=REF_ROOT)
    cmd = [sys.executable, '-W', 'ignore', '-m', 'pyjac', '--lang', 'c',
           '--input', os.path.abspath(mech), '-b', src]
    if last_spec:
        cmd += ['-ls', last_spec]
    res = subprocess.run(cmd, env=env, cwd=out, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('reference codegen failed:\n' + res.stdout + res.stderr)

    if conv:
        hdr = os.path.join(src, 'header.h')
        txt = open(hdr).read()
        assert '#define CONP\n//#define CONV' in txt
        with open(hdr, 'w') as fh:
            fh.write(txt.replace('#define CONP\n//#define CONV', '//#define CONP\n#define CONV'))
        # The constant-volume branch is dead code upstream and does not compile as emitted: the generator
        # leaves out a comma (rate_subs.py:2361-2363 -> "eval_conc_rho (y[0]rho, ...") and a plus sign
        # (rate_subs.py:2428-2430 -> "(cv[8] * y[9])(cv[9] * y_N)").  The two characters are put back in the
        # emitted copy; everything else is the generator's text.
        dy_c = os.path.join(src, 'dydt.c')
        txt = open(dy_c).read()
        assert 'eval_conc_rho (y[0]rho,' in txt
        txt = txt.replace('eval_conc_rho (y[0]rho,', 'eval_conc_rho (y[0], rho,')
        import re as _re
        txt, nfix = _re.subn(r'\)\(cv\[(\d+)\] \* y_N\)', r') + (cv[\1] * y_N)', txt)
        assert nfix == 1
        with open(dy_c, 'w') as fh:
            fh.write(txt)
    cfiles = [f for f in glob.glob(os.path.join(src, '**', '*.c'), recursive=True)]
    if conv:
        # eval_jacob has no constant-volume form upstream (its dydt prototype no longer matches)
        cfiles = [f for f in cfiles if 'jacob' not in os.path.basename(f) and os.sep + 'jacobs' + os.sep not in f]
    cfiles.append(os.path.join(HERE, 'ref_batch.c'))
    inc = ['-I', src, '-I', os.path.join(src, 'jacobs'), '-I', os.path.join(src, 'rates')]
    flags = ['-std=c99', opt, '-mtune=native', '-fPIC', '-fopenmp', '-D_DEFAULT_SOURCE'] + (['-DREF_NO_JACOB'] if conv else [])

    def cc(f):
        o = os.path.join(out, 'obj_' + os.path.relpath(f, '/').replace('/', '_')[:-2] + '.o')
        r = subprocess.run(['gcc'] + flags + inc + ['-c', f, '-o', o],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('gcc failed on %s:\n%s' % (f, r.stderr))
        return o

    with ThreadPoolExecutor(jobs) as ex:
        objs = list(ex.map(cc, cfiles))
    r = subprocess.run(['gcc', '-shared', '-fopenmp', '-o', lib] + objs + ['-lm'],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stderr)
    for o in objs:
        os.remove(o)
    with open(stamp, 'w') as fh:
        fh.write(mech_txt)
    if not quiet:
        print('built', lib)
    return lib


def cuda_lib_path(name: str) -> str:
    return os.path.join(ref_dir(name + '_cuda'), 'libcu_pyjac.so')


def build_cuda(name: str, mech: str, jobs: int = 8, force: bool = False, quiet: bool = True) -> str:
    """The reference's *generated CUDA* for a mechanism, rebuilt for sm_100a:
    `python -m pyjac --lang cuda` where the reference lies, then nvcc on the emitted files
    (the reference's own flags use -arch=sm_20, libgen.py:45, and helper_cuda.h from the CUDA
    samples, libgen.py:38-40, absent in CUDA 12.9: an empty stand-in header is used), linked
    with oracle/ref_cuda_driver.cu into oracle/_ref/<name>_cuda/libcu_pyjac.so."""
    out = ref_dir(name + '_cuda')
    src = os.path.join(out, 'src')
    lib = cuda_lib_path(name)
    stamp = os.path.join(out, 'mech.inp')
    mech_txt = open(mech).read()
    if (not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read() == mech_txt):
        return lib
    if not have_reference():
        raise RuntimeError('reference sources not present at %s' % REF_ROOT)
    os.makedirs(src, exist_ok=True)
    env = dict(os.environ, PYTHONPATH=REF_ROOT)
    res = subprocess.run([sys.executable, '-W', 'ignore', '-m', 'pyjac', '--lang', 'cuda',
                          '--input', os.path.abspath(mech), '-b', src], env=env, cwd=out,
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('reference codegen failed:\n' + res.stdout + res.stderr)
    stub = os.path.join(out, 'stub')
    os.makedirs(stub, exist_ok=True)
    with open(os.path.join(stub, 'helper_cuda.h'), 'w') as fh:
        fh.write('/* empty stand-in: CUDA 12.9 ships no samples/common/inc/helper_cuda.h */\n')
    files = [f for f in glob.glob(os.path.join(src, '**', '*.cu'), recursive=True)
             if os.path.basename(f) != 'sparse_multiplier.cu']      # broken upstream
    files.append(os.path.join(HERE, 'ref_cuda_driver.cu'))
    flags = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-rdc=true', '-Xcompiler', '-fPIC',
             '-I', stub, '-I', src, '-I', os.path.join(src, 'jacobs'), '-I', os.path.join(src, 'rates')]

    def cc(f):
        o = os.path.join(out, 'obj_' + os.path.basename(f)[:-3] + '.o')
        r = subprocess.run(['nvcc'] + flags + ['-c', f, '-o', o], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed on %s:\n%s' % (f, r.stderr))
        return o

    with ThreadPoolExecutor(jobs) as ex:
        objs = list(ex.map(cc, files))
    r = subprocess.run(['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', lib] + objs,
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stderr)
    for o in objs:
        os.remove(o)
    with open(stamp, 'w') as fh:
        fh.write(mech_txt)
    if not quiet:
        print('built', lib)
    return lib


def harness_path(name: str) -> str:
    return os.path.join(ref_dir(name), 'speedtest')


def build_harness(name: str, jobs: int = 8, force: bool = False) -> str:
    """The reference's own `speedtest` for mechanism `name` (built by :func:`build` before): its generated
    C, tester.c.in (datafile "data.bin"), read_initial_conditions.c and timer.h compiled where they lie with
    the reference's flags (libgen.py:43, performance_tester.py:480-482) -> oracle/_ref/<name>/speedtest.
    `speedtest <num_odes> <num_threads>` in a directory holding data.bin prints "num_odes,ms": the CPU arm
    of bench.py times the reference through it."""
    from string import Template
    out = ref_dir(name)
    src = os.path.join(out, 'src')
    exe = harness_path(name)
    if not force and os.path.exists(exe) and os.path.getmtime(exe) >= os.path.getmtime(lib_path(name)):
        return exe
    if not have_reference():
        raise RuntimeError('reference sources not present at %s' % REF_ROOT)
    home = os.path.join(REF_ROOT, 'pyjac', 'performance_tester')
    with open(os.path.join(home, 'tester.c.in')) as fh:
        main_c = Template(fh.read()).substitute(datafile='data.bin')
    test_c = os.path.join(out, 'test.c')
    with open(test_c, 'w') as fh:
        fh.write(main_c)
    cfiles = glob.glob(os.path.join(src, '**', '*.c'), recursive=True) + [test_c, os.path.join(home, 'read_initial_conditions.c')]
    inc = ['-I', src, '-I', os.path.join(src, 'jacobs'), '-I', os.path.join(src, 'rates'), '-I', home]
    flags = ['-std=c99', '-O3', '-mtune=native', '-fopenmp', '-D_DEFAULT_SOURCE']

    def cc(f):
        o = os.path.join(out, 'hobj_' + os.path.relpath(f, '/').replace('/', '_')[:-2] + '.o')
        r = subprocess.run(['gcc'] + flags + inc + ['-c', f, '-o', o], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('gcc failed on %s:\n%s' % (f, r.stderr))
        return o

    with ThreadPoolExecutor(jobs) as ex:
        objs = list(ex.map(cc, cfiles))
    r = subprocess.run(['gcc', '-fopenmp', '-o', exe] + objs + ['-lm'], capture_output=True, text=True)
    for o in objs:
        os.remove(o)
    os.remove(test_c)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stderr)
    return exe


def speedtest_path(name: str) -> str:
    return os.path.join(ref_dir(name + '_speedtest'), 'speedtest')


def build_speedtest(name: str, mech: str, force: bool = False) -> str:
    """The reference's own performance harness -- pyjac/performance_tester/tester.c.in (its $datafile
    filled in with "data.bin", as performance_tester.py:430-445 does with string.Template),
    read_initial_conditions.c and timer.h, compiled where they lie with the reference's flags
    (performance_tester.py:480-482: -std=c99 -O3 -mtune=native -fopenmp; -D_DEFAULT_SOURCE for timersub
    on current glibc) -- against the headers and the per-mechanism stub library that pyjac_b200 writes
    (create_jacobian / libgen.generate_library('c', ...)) instead of the generated sources:
    the drop-in proof of the scalar C API.  -> oracle/_ref/<name>_speedtest/speedtest, run as
    ``speedtest <num_odes> <num_threads>`` in a directory holding data.bin."""
    from string import Template
    sys.path.insert(0, os.path.dirname(HERE))
    from pyjac_b200 import create_jacobian as cj, libgen
    out = ref_dir(name + '_speedtest')
    exe = speedtest_path(name)
    stamp = os.path.join(out, 'mech.inp')
    mech_txt = open(mech).read()
    if not force and os.path.exists(exe) and os.path.exists(stamp) and open(stamp).read() == mech_txt \
            and os.path.getmtime(exe) >= os.path.getmtime(libgen.LIB_PATH):
        return exe
    if not have_reference():
        raise RuntimeError('reference sources not present at %s' % REF_ROOT)
    os.makedirs(out, exist_ok=True)
    cj.create_jacobian('cuda', mech, build_path=out)
    libgen.generate_library('c', out, out_dir=out)
    home = os.path.join(REF_ROOT, 'pyjac', 'performance_tester')
    with open(os.path.join(home, 'tester.c.in')) as fh:
        main_c = Template(fh.read()).substitute(datafile='data.bin')
    test_c = os.path.join(out, 'test.c')
    with open(test_c, 'w') as fh:                       # generated from the template, like the reference's build dir
        fh.write(main_c)
    cmd = ['gcc', '-std=c99', '-O3', '-mtune=native', '-fopenmp', '-D_DEFAULT_SOURCE', '-I', out, '-I', home,
           test_c, os.path.join(home, 'read_initial_conditions.c'), '-o', exe, '-L', out, '-lc_pyjac',
           '-L', libgen.BUILD, '-lpyjac_b200', '-lm',
           '-Wl,-rpath,$ORIGIN', '-Wl,-rpath,$ORIGIN/../../../pyjac_b200/_build']
    r = subprocess.run(cmd, capture_output=True, text=True)
    os.remove(test_c)
    if r.returncode != 0:
        raise RuntimeError('building the reference harness against pyjac_b200 failed:\n' + r.stderr)
    with open(stamp, 'w') as fh:
        fh.write(mech_txt)
    return exe


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('name')
    ap.add_argument('mech')
    ap.add_argument('--last-spec', default=None)
    ap.add_argument('--opt', default='-O3')
    ap.add_argument('--jobs', type=int, default=os.cpu_count() or 4)
    ap.add_argument('--force', action='store_true')
    a = ap.parse_args()
    print(build(a.name, a.mech, a.last_spec, a.opt, a.jobs, a.force, quiet=False))
