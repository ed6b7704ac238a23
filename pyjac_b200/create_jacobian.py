"""``create_jacobian`` -- the build entry of the pipeline, with the reference's signature
(pyjac/core/create_jacobian.py:3407-3412).

The reference writes ~18 generated source files into ``build_path``.  Here the mechanism is
exported as data: ``mechanism.h`` keeps the macros and the ``//last_spec`` comment that other
tools parse by regex (functional_tester/test.py:311-318,358; libgen.py:385), and
``mechanism.pjt`` holds the device tables (pyjac_b200/tables.py, blob.py) that the fixed
sm_100a library interprets.  ``libgen.generate_library`` then hands back that library.
"""
from __future__ import annotations

import os
from typing import Optional

from . import blob, tables
from .mechanism import Mechanism

TABLE_FILE = 'mechanism.pjt'
HEADER_FILE = 'mechanism.h'

# The headers a program written against the emitted library includes (create_jacobian.py:2226-2248,
# rate_subs.py:292-323,1581-1608,2130-2150, mech_auxiliary.py:109-206,440-478): same names, same
# prototypes; the definitions live in the fixed CUDA library (include/pyjac_b200.h surface 3).
_HEADERS = {
    'header.h': '''#ifndef HEAD
#define HEAD
#include <stdlib.h>
#include <math.h>

/** Constant pressure or volume. */
#define CONP
//#define CONV

/** Include mechanism header to get NSP and NN **/
#include "mechanism.h"
// OpenMP
#ifdef _OPENMP
 #include <omp.h>
#else
 #define omp_get_max_threads() 1
 #define omp_get_num_threads() 1
#endif
#endif
''',
    'jacob.h': '''#ifndef JACOB_HEAD
#define JACOB_HEAD

#include "header.h"
#include "chem_utils.h"
#include "rates.h"

#ifdef __cplusplus
extern "C" {
#endif
void eval_jacob (const double, const double, const double * __restrict__, double * __restrict__);
#ifdef __cplusplus
}
#endif

#endif
''',
    'dydt.h': '''#ifndef DYDT_HEAD
#define DYDT_HEAD

#include "header.h"

#ifdef __cplusplus
extern "C" {
#endif
void dydt (const double, const double, const double * __restrict__, double * __restrict__);
#ifdef __cplusplus
}
#endif

#endif
''',
    'rates.h': '''#ifndef RATES_HEAD
#define RATES_HEAD

#include "header.h"

#ifdef __cplusplus
extern "C" {
#endif
void eval_rxn_rates (const double, const double, const double * __restrict__, double * __restrict__, double * __restrict__);
void eval_spec_rates (const double * __restrict__, const double * __restrict__, const double * __restrict__, double * __restrict__, double * __restrict__);
void get_rxn_pres_mod (const double, const double, const double * __restrict__, double * __restrict__);
#ifdef __cplusplus
}
#endif

#endif
''',
    'chem_utils.h': '''#ifndef CHEM_UTILS_HEAD
#define CHEM_UTILS_HEAD

#include "header.h"

#ifdef __cplusplus
extern "C" {
#endif
void eval_conc (const double, const double, const double * __restrict__, double * __restrict__, double * __restrict__, double * __restrict__, double * __restrict__);
void eval_h (const double, double * __restrict__);
void eval_u (const double, double * __restrict__);
void eval_cv (const double, double * __restrict__);
void eval_cp (const double, double * __restrict__);
#ifdef __cplusplus
}
#endif

#endif
''',
    'pyjacob.cuh': '''/* wrapper to translate to cuda arrays */

#ifndef CU_PYJAC_HEAD
#define CU_PYJAC_HEAD

void run(int, int, const double*, const double*,
			double*, double*, double*,
			double*, double*, double*, double*);
int init(int);
void cleanup();

#endif
''',
}


def _header(mech: Mechanism) -> str:
    lines = ['#ifndef MECHANISM_H', '#define MECHANISM_H', '',
             '/* written by pyjac_b200.create_jacobian: metadata only, the mechanism itself is in',
             '   %s (device tables) */' % TABLE_FILE, '',
             '//last_spec %d' % mech.last_spec_original,
             '/* Species Indexes']
    lines += ['%d  %s' % (i, sp.name) for i, sp in enumerate(mech.specs)]
    lines += ['*/', '',
              '/* Number of species */', '#define NSP %d' % mech.NSP,
              '/* Number of variables. NN = NSP + 1 (temperature) */', '#define NN %d' % (mech.NSP + 1),
              '/* Number of forward reactions */', '#define FWD_RATES %d' % mech.FWD_RATES,
              '/* Number of reversible reactions */', '#define REV_RATES %d' % mech.REV_RATES,
              '/* Number of reactions with pressure modified rates */',
              '#define PRES_MOD_RATES %d' % mech.PRES_MOD_RATES, '',
              '#ifdef __cplusplus', 'extern "C" {', '#endif',
              '//apply masking of ICs for cache optimized mechanisms', 'void apply_mask(double*);',
              'void apply_reverse_mask(double*);',
              '#ifdef __cplusplus', '}', '#endif', '', '#endif', '']
    return '\n'.join(lines)


def create_jacobian(lang, mech_name=None, therm_name=None, gas=None, optimize_cache=False,
                    initial_state='', num_blocks=8, num_threads=64, no_shared=False,
                    L1_preferred=True, multi_thread=None, force_optimize=False,
                    build_path='./out/', last_spec=None, skip_jac=False, auto_diff=False,
                    gs: int = 0, threads: int = 0, conp: bool = True) -> Mechanism:
    """Export the mechanism ``mech_name`` (Chemkin format with optional ``therm_name``, or a Cantera ``.cti``
    file -- read without Cantera, :mod:`pyjac_b200.cti_interpret`) for the
    B200 library into ``build_path``.

    Only ``lang='cuda'`` exists (there is no CPU back end).  The code-generation tuning knobs
    of the reference (``optimize_cache``, ``num_blocks``, ``num_threads``, ``no_shared``,
    ``L1_preferred``, ``multi_thread``, ``force_optimize``) have nothing to act on and are
    accepted and ignored; ``gas`` (a Cantera object), ``auto_diff`` and ``initial_state`` are
    rejected.  ``gs`` / ``threads`` choose the Jacobian kernel's plan (0 = automatic).
    ``conp=False`` exports the constant-volume problem: header.h then says ``#define CONV`` (the reference
    hard-wires CONP there, mech_auxiliary.py:464-466, and a user edits the file) and ``dydt(t, rho, y, dy)``
    of a library built from this directory takes the density; ``eval_jacob`` has no such form.
    """
    if lang != 'cuda':
        raise ValueError("pyjac_b200 only targets CUDA (sm_100a); lang=%r" % (lang,))
    if gas is not None:
        raise NotImplementedError('Cantera input is not supported (Chemkin files only)')
    if auto_diff or initial_state:
        raise NotImplementedError('auto_diff / initial_state are outside the hot path')
    if mech_name is None:
        raise ValueError('mech_name is required')
    mech = Mechanism.from_file(mech_name, therm_name, last_spec)
    os.makedirs(build_path, exist_ok=True)
    with open(os.path.join(build_path, HEADER_FILE), 'w') as fh:
        fh.write(_header(mech))
    for name, text in _HEADERS.items():
        if name == 'header.h' and not conp:
            text = text.replace('#define CONP\n//#define CONV', '//#define CONP\n#define CONV')
        with open(os.path.join(build_path, name), 'w') as fh:
            fh.write(text)
    if not skip_jac:
        T = tables.build(mech, gs=gs, threads=threads, conv=not conp)
        with open(os.path.join(build_path, TABLE_FILE), 'wb') as fh:
            fh.write(blob.pack(T))
    return mech


def load_tables(build_path: str) -> bytes:
    """The table blob written by :func:`create_jacobian`."""
    with open(os.path.join(build_path, TABLE_FILE), 'rb') as fh:
        return fh.read()
