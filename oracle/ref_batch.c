/* TEST INFRASTRUCTURE -- not part of the shipped product path.
 *
 * Batch driver linked into oracle/_ref/<mech>/libc_pyjac.so next to the reference's own
 * generated C (emitted by /root/reference/pyjac from a Chemkin file, see build_ref.py).
 * It only loops the reference's scalar entry points over a state batch the way the
 * reference's timing harness does (pyjac/performance_tester/tester.c.in:23-31:
 * `#pragma omp parallel for`, a zeroed stack `jac[NSP*NSP]` per state).
 *
 * States are row-major [n][NSP]: T, Y_0 .. Y_{NSP-2} (the layout eval_jacob takes).
 */
#include <string.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "mechanism.h"

void eval_jacob(const double t, const double pres, const double* y, double* jac);
void dydt(const double t, const double pres, const double* y, double* dy);

int ref_nsp(void) { return NSP; }
int ref_fwd_rates(void) { return FWD_RATES; }
int ref_rev_rates(void) { return REV_RATES; }
int ref_pres_mod_rates(void) { return PRES_MOD_RATES; }

#ifndef REF_NO_JACOB          /* the constant-volume build (build_ref.py conv=True) has no eval_jacob */
/* jac may be NULL: results are then discarded exactly like tester.c.in does. */
void ref_eval_jacob_batch(int n, const double* pres, const double* y, double* jac, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    #pragma omp parallel for
    for (int s = 0; s < n; ++s) {
        if (jac) {
            double* out = jac + (size_t)s * NSP * NSP;
            memset(out, 0, sizeof(double) * NSP * NSP);
            eval_jacob(0.0, pres[s], y + (size_t)s * NSP, out);
        } else {
            double* local = (double*)calloc((size_t)NSP * NSP, sizeof(double));
            eval_jacob(0.0, pres[s], y + (size_t)s * NSP, local);
            free(local);
        }
    }
}

#endif

void ref_dydt_batch(int n, const double* pres, const double* y, double* dy, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    #pragma omp parallel for
    for (int s = 0; s < n; ++s)
        dydt(0.0, pres[s], y + (size_t)s * NSP, dy + (size_t)s * NSP);
}
