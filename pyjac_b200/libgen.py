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


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/pyjac_b200.cu for sm_100a if the in-tree library is missing or stale."""
    if not force and os.path.exists(LIB_PATH):
        built = os.path.getmtime(LIB_PATH)
        if all(os.path.getmtime(s) <= built for s in _sources()):
            return LIB_PATH
        if not (shutil.which('nvcc') or os.path.exists('/usr/local/cuda/bin/nvcc')):
            return LIB_PATH        # a box without nvcc uses the library it was shipped
    os.makedirs(BUILD, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + ['-I', os.path.join(ROOT, 'include'), '-I', CSRC,
                                    '-o', LIB_PATH, os.path.join(CSRC, 'pyjac_b200.cu')]
    cmd[1:1] = os.environ.get('PYJAC_B200_NVCC_EXTRA', '').split()      # development builds
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


def generate_library(lang: str, source_dir: str, obj_dir: Optional[str] = None,
                     out_dir: Optional[str] = None, shared: Optional[bool] = None,
                     finite_difference: bool = False, auto_diff: bool = False) -> str:
    """Signature of pyjac/libgen/libgen.py:322.  ``source_dir`` is a directory written by
    :func:`pyjac_b200.create_jacobian.create_jacobian` (mechanism.h + mechanism tables);
    the returned path is the sm_100a library, copied into ``out_dir`` when one is given.
    Only ``lang='cuda'`` exists -- there is no C (CPU) back end to fall back to."""
    if lang != 'cuda':
        raise ValueError("pyjac_b200 only builds the CUDA (sm_100a) library; lang=%r" % (lang,))
    if finite_difference or auto_diff:
        raise NotImplementedError('finite-difference / autodiff comparison libraries are out of scope')
    if not os.path.isfile(os.path.join(source_dir, 'mechanism.h')):
        raise FileNotFoundError('%s holds no mechanism.h; run create_jacobian first' % source_dir)
    lib = build_library()
    if out_dir and os.path.abspath(out_dir) != BUILD:
        os.makedirs(out_dir, exist_ok=True)
        dst = os.path.join(out_dir, LIB_NAME)
        shutil.copy2(lib, dst)
        return dst
    return lib
