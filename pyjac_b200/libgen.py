"""Build of the fixed CUDA library -- the ``pyjac.libgen`` stage of the pipeline.

The reference compiles the *generated* per-mechanism sources, one ``gcc`` / ``nvcc -arch=sm_20
-dc`` per file, then archives them into ``libc_pyjac`` / ``libcu_pyjac``
(pyjac/libgen/libgen.py:43-46,149-215,322-411).  Here there is nothing to generate: one
translation unit (csrc/pyjac_b200.cu) is compiled once for sm_100a into
``pyjac_b200/_build/libpyjac_b200.so`` and serves every mechanism; ``generate_library``
keeps the reference's signature and hands that library back.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
BUILD = os.path.join(HERE, '_build')
LIB_NAME = 'libpyjac_b200.so'
LIB_PATH = os.path.join(BUILD, LIB_NAME)

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-shared', '-Xcompiler', '-fPIC', '-Xcompiler', '-O2']


def _sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))
            if f.endswith(('.cu', '.cuh', '.h'))] + [os.path.join(ROOT, 'include', 'pyjac_b200.h')]


def _nvcc() -> str:
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found; the CUDA library cannot be built (there is no CPU fallback)')
    return exe


def _fresh() -> bool:
    if not os.path.exists(LIB_PATH):
        return False
    built = os.path.getmtime(LIB_PATH)
    return all(os.path.getmtime(s) <= built for s in _sources())


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/pyjac_b200.cu for sm_100a if the in-tree library is missing or stale.  Safe when several
    processes (one per GPU under torchrun) get here together: the build runs under an exclusive file lock, into a
    temporary file that replaces the library atomically, and whoever waited for the lock finds it fresh."""
    import fcntl
    if not force and _fresh():
        return LIB_PATH
    if not force and os.path.exists(LIB_PATH) and not (shutil.which('nvcc') or os.path.exists('/usr/local/cuda/bin/nvcc')):
        return LIB_PATH            # a box without nvcc uses the library it was shipped
    os.makedirs(BUILD, exist_ok=True)
    with open(os.path.join(BUILD, '.build.lock'), 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and _fresh():
                return LIB_PATH
            tmp = LIB_PATH + '.tmp%d' % os.getpid()
            cmd = [_nvcc()] + NVCC_FLAGS + ['-I', os.path.join(ROOT, 'include'), '-I', CSRC,
                                            '-o', tmp, os.path.join(CSRC, 'pyjac_b200.cu')]
            if verbose:
                cmd.insert(1, '-Xptxas')
                cmd.insert(2, '-v')
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError('nvcc failed:\n' + res.stdout + res.stderr)
            os.replace(tmp, LIB_PATH)
            if verbose:
                print(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


_STUB_C = '''/* written by pyjac_b200.libgen.generate_library: the mechanism's tables, registered with the fixed
   CUDA library when this stub is loaded (include/pyjac_b200.h: pyjac_register_tables) */
#define _GNU_SOURCE
#include <stddef.h>
#include <dlfcn.h>
extern int pyjac_register_tables(const void* blob, size_t len);
__asm__(".section .rodata\\n.balign 16\\n.global pyjac_b200_tables_start\\npyjac_b200_tables_start:\\n"
        ".incbin \\"%(pjt)s\\"\\n.global pyjac_b200_tables_end\\npyjac_b200_tables_end:\\n.previous\\n");
extern const char pyjac_b200_tables_start[], pyjac_b200_tables_end[];
__attribute__((constructor)) static void pyjac_b200_register(void)
{
    pyjac_register_tables(pyjac_b200_tables_start, (size_t)(pyjac_b200_tables_end - pyjac_b200_tables_start));
}

/* The entry points of the emitted library, forwarded to the CUDA library: a program that links
   -lc_pyjac alone (performance_tester.py:466-470) resolves them here, so the linker keeps this stub
   -- and with it the mechanism -- even under --as-needed. */
#define FWD(ret, name, decl, call)                                         \\
    ret name decl                                                          \\
    {                                                                      \\
        static ret (*fn) decl;                                             \\
        if (!fn) fn = (ret (*) decl)dlsym(RTLD_NEXT, #name);               \\
        return fn call;                                                    \\
    }
FWD(void, eval_jacob, (const double t, const double p, const double* y, double* jac), (t, p, y, jac))
FWD(void, dydt, (const double t, const double p, const double* y, double* dy), (t, p, y, dy))
FWD(void, eval_conc, (const double T, const double p, const double* mf, double* yN, double* mw, double* rho, double* c), (T, p, mf, yN, mw, rho, c))
FWD(void, eval_rxn_rates, (const double T, const double p, const double* C, double* f, double* r), (T, p, C, f, r))
FWD(void, get_rxn_pres_mod, (const double T, const double p, const double* C, double* pm), (T, p, C, pm))
FWD(void, eval_spec_rates, (const double* f, const double* r, const double* pm, double* sp, double* dyN), (f, r, pm, sp, dyN))
FWD(void, eval_h, (const double T, double* o), (T, o))
FWD(void, eval_u, (const double T, double* o), (T, o))
FWD(void, eval_cv, (const double T, double* o), (T, o))
FWD(void, eval_cp, (const double T, double* o), (T, o))
FWD(void, apply_mask, (double* y), (y))
FWD(void, apply_reverse_mask, (double* y), (y))
'''


def lib_name(lang: str, shared: bool = True) -> str:
    """libc_pyjac / libcu_pyjac, the names of pyjac/libgen/libgen.py:170-186."""
    return 'lib%s_pyjac%s' % ('cu' if lang == 'cuda' else 'c', '.so' if shared else '.a')


def generate_library(lang: str, source_dir: str, obj_dir: Optional[str] = None,
                     out_dir: Optional[str] = None, shared: Optional[bool] = None,
                     finite_difference: bool = False, auto_diff: bool = False) -> str:
    """Signature of pyjac/libgen/libgen.py:322.  ``source_dir`` is a directory written by
    :func:`pyjac_b200.create_jacobian.create_jacobian` (headers + mechanism tables).

    The reference compiles the generated sources into ``libc_pyjac`` / ``libcu_pyjac``; here the compute
    library is fixed (``libpyjac_b200.so``, built for sm_100a on first use) and what is produced per
    mechanism is a *stub* of that name: a small shared library that embeds the table blob, registers
    it with the CUDA library when loaded and depends on it -- so ``-lc_pyjac`` (what the reference's
    harness links, performance_tester.py:466-470) brings the entry points of ``jacob.h`` & co. with the
    mechanism baked in, as the reference's library does.  ``lang`` only picks the name: 'c' =
    ``libc_pyjac.so`` (the host-callable C entry points), 'cuda' = ``libcu_pyjac.so``; both run on the
    GPU -- there is no CPU back end.  Returns the stub's path (in ``out_dir``, default ``source_dir``)."""
    if lang not in ('c', 'cuda'):
        raise ValueError("lang must be 'c' or 'cuda' (both name the sm_100a library); got %r" % (lang,))
    if finite_difference or auto_diff:
        raise NotImplementedError('finite-difference / autodiff comparison libraries are out of scope')
    if shared is False:
        raise NotImplementedError('only shared libraries are built (the stub must run a constructor)')
    pjt = os.path.abspath(os.path.join(source_dir, 'mechanism.pjt'))
    if not os.path.isfile(os.path.join(source_dir, 'mechanism.h')) or not os.path.isfile(pjt):
        raise FileNotFoundError('%s holds no mechanism.h / mechanism.pjt; run create_jacobian first' % source_dir)
    lib = build_library()
    out_dir = os.path.abspath(out_dir or source_dir)
    obj_dir = os.path.abspath(obj_dir or out_dir)
    os.makedirs(out_dir, exist_ok=True)
    os.makedirs(obj_dir, exist_ok=True)
    stub_c = os.path.join(obj_dir, 'pyjac_b200_stub.c')
    with open(stub_c, 'w') as fh:
        fh.write(_STUB_C % {'pjt': pjt})
    out = os.path.join(out_dir, lib_name(lang))
    gcc = shutil.which('gcc') or shutil.which('cc')
    if not gcc:
        raise RuntimeError('gcc not found: cannot build the per-mechanism stub library')
    # the stub finds the CUDA library next to itself or where it was built
    cmd = [gcc, '-shared', '-fPIC', '-O1', stub_c, '-o', out, '-L', BUILD, '-lpyjac_b200',
           '-Wl,-rpath,$ORIGIN', '-Wl,-rpath,' + BUILD]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('building %s failed:\n%s' % (out, res.stderr))
    return out
