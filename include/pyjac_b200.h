/* pyjac_b200 -- C ABI of the B200-native batched kinetics evaluator.
 *
 * One shared library (pyjac_b200/_build/libpyjac_b200.so) replaces the library that the
 * reference *generates and compiles per mechanism* (libc_pyjac / libcu_pyjac,
 * pyjac/libgen/libgen.py:170-186).  The mechanism arrives at run time as a table blob
 * (pyjac_b200/tables.py + pyjac_b200/blob.py) instead of being unrolled into source.
 *
 * Three surfaces, each citing the reference interface it stands in for:
 *
 *  1. device-pointer batch API (new; the reference has no equivalent -- its GPU entry
 *     point pyjac/pywrap/pyjacob.cu:134-188 always round-trips through host memory),
 *  2. host-pointer batch API = pyjac/pywrap/pyjacob.cuh:6-10 (init / run / cleanup bound by
 *     pyjac/pywrap/pyjacob_cuda_wrapper.pyx:5-34),
 *  3. scalar API with the emitted library's own names and signatures
 *     (headers emitted at pyjac/core/create_jacobian.py:2226-2248,
 *     pyjac/core/rate_subs.py:292-323,1581-1608,2130-2150; bound by
 *     pyjac/pywrap/pyjacob_wrapper.pyx:4-16 and linked by
 *     pyjac/performance_tester/tester.c.in:2,28).
 *
 * Conventions: plain pointers and sizes only.  Species / reaction order is pyJac's internal
 * order (last species moved to the end, utils.py:55-91).  A state is y = [T, Y_0..Y_{NSP-2}]
 * plus a pressure; a Jacobian is NSP x NSP column-major, jac[i + NSP*j] = d f_i / d y_j
 * (docs/faqs.rst:82-87).  Unlike the reference (all entry points void, exit() on error) the
 * pyjac_* functions return 0 on success or a negative PYJAC_E* code and keep a message
 * retrievable with pyjac_last_error(); the reference-named scalar functions keep the
 * reference's behaviour (message on stderr, exit(1)).  There is NO CPU fallback: without a
 * CUDA device every compute entry point fails with PYJAC_ENODEVICE.
 */
#ifndef PYJAC_B200_H
#define PYJAC_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pyjac_mech pyjac_mech;

enum {
    PYJAC_OK = 0,
    PYJAC_EINVAL = -1,      /* bad argument / malformed table blob */
    PYJAC_ENODEVICE = -2,   /* no usable CUDA device */
    PYJAC_ECUDA = -3,       /* CUDA runtime error (see pyjac_last_error) */
    PYJAC_ENOMEM = -4,      /* device or host allocation failed */
    PYJAC_ETOOBIG = -5      /* mechanism does not fit the kernel's on-chip working set */
};

/* Jacobian output layouts */
enum {
    PYJAC_JAC_STATE_MAJOR = 0,  /* jac[s*NSP*NSP + i + NSP*j]: one contiguous column-major
                                   Jacobian per state -- what eval_jacob() writes for one state */
    PYJAC_JAC_STATE_FASTEST = 1 /* jac[(i + NSP*j)*ld + s]: the reference GPU layout
                                   (mech_auxiliary.py:418-420 INDEX(); docs/faqs.rst:163-172) */
};

const char* pyjac_last_error(void);
int pyjac_device_count(void);

/* ---- mechanism handle ------------------------------------------------------------- */
/* Parses a PJB200T1 table blob, uploads the tables to `device` (-1 = current device). */
int pyjac_mech_create(const void* blob, size_t len, int device, pyjac_mech** out);
void pyjac_mech_destroy(pyjac_mech* m);
/* dims[0..3] = NSP, FWD_RATES, REV_RATES, PRES_MOD_RATES (the mechanism.h macros,
 * mech_auxiliary.py:109-176) */
int pyjac_mech_dims(const pyjac_mech* m, int dims[4]);
/* Launch tuning: cap on resident thread blocks per SM (0 = automatic).  States per block and
 * block size belong to the plan inside the table blob (pyjac_b200/plan.py).  Results do not
 * depend on this. */
int pyjac_mech_tune(pyjac_mech* m, int blocks_per_sm);
/* (mangled) symbol of the kernel that a call of mode 0 = eval_jacob, 1 = dydt, 2 = the rate routines,
 * 3 = the factored Jacobian launches for this mechanism's plan -- what a profiler lists */
int pyjac_mech_kernel_name(const pyjac_mech* m, int mode, char* buf, size_t len);
/* number of kernels launched through this handle since creation */
long long pyjac_mech_launches(const pyjac_mech* m);

/* ---- 1. device-pointer batch API -------------------------------------------------- */
/* All pointers are device pointers on the handle's device; `stream` is a cudaStream_t
 * (NULL = default stream); calls are asynchronous.
 * State input: element v (0 = T, 1.. = Y_{v-1}) of state s is d_y[s*y_ss + v*y_sv]:
 *   state-fastest ("SoA", the reference GPU layout):  y_ss = 1,   y_sv = ld
 *   one row per state (what the scalar API takes):    y_ss = NSP, y_sv = 1            */
int pyjac_eval_jacob_dev(pyjac_mech* m, int n, const double* d_pres, const double* d_y,
                         long long y_ss, long long y_sv, double* d_jac, int jac_layout,
                         long long jac_ld, void* stream);
/* dy element v of state s -> d_dy[s*o_ss + v*o_sv] */
int pyjac_dydt_dev(pyjac_mech* m, int n, const double* d_pres, const double* d_y,
                   long long y_ss, long long y_sv, double* d_dy, long long o_ss,
                   long long o_sv, void* stream);
/* Constant-volume dydt (the reference's `#define CONV` form, pyjac/core/rate_subs.py:2340-2485; compiled
 * out upstream by header.h, mech_auxiliary.py:464-466, and not compilable as emitted -- DESIGN.md): d_rho holds
 * one density per state in place of the pressure, dT/dt = -sum_k wdot_k u_k W_k / (rho cv_avg). */
int pyjac_dydt_conv_dev(pyjac_mech* m, int n, const double* d_rho, const double* d_y,
                        long long y_ss, long long y_sv, double* d_dy, long long o_ss,
                        long long o_sv, void* stream);
/* conv != 0: the reference-named dydt(t, rho, y, dy) of this handle is the constant-volume one -- what
 * editing header.h to `#define CONV` selects in the reference; eval_jacob then fails (no such form exists). */
int pyjac_mech_set_conv(pyjac_mech* m, int conv);
/* Finite-difference Jacobian of dydt, the independent on-device check of eval_jacob; replaces
 * the reference's finite-difference comparison build (pyjac/performance_tester/fd_jacob.cu:23-95:
 * same CVODE-style increments, ATOL 1e-15, RTOL 1e-8).  order 1 = the reference's default forward
 * difference; 2, 4, 6 = central differences (the reference's FD_ORD > 1 branch sums y_temp instead
 * of dy, fd_jacob.cu:84 -- the intended formula is used here).  r_cap = 0: the reference's
 * increments; r_cap > 0: increments kept within [r_cap / 100, r_cap] * max(|y_j|, 1) (far from
 * equilibrium the reference's r0 term exceeds the mass fractions, at equilibrium it vanishes).  d_y[NSP][n] and d_jac[NSP*NSP][n]
 * state-fastest with leading dimension n; synchronous (scratch memory is allocated and freed). */
int pyjac_fd_jacob_dev(pyjac_mech* m, int n, const double* d_pres, const double* d_y, double* d_jac,
                       int order, double r_cap, void* stream);
/* ---- factored Jacobian and its consumers ---------------------------------
 * The dense NSP x NSP Jacobian is the expansion of a much smaller record; the reference only gestures
 * at a sparse form (create_jacobian.py:3301-3404, `sparse_multiplier`, broken at :3322).  Per state the
 * record holds NF = NSP + 3 (NSP - 1) + NNZ doubles:
 *     fac[0 .. NSP)              J[0][j]                      the energy-equation row
 *     fac[NSP + k]               J[k+1][0]                    the temperature column, k = 0 .. NSP-2
 *     fac[NSP + (NSP-1) + k]     WA_k
 *     fac[NSP + 2 (NSP-1) + k]   WB_k
 *     fac[NSP + 3 (NSP-1) + p]   S_p, entry (rows[p], cols[p]) of the sparse block, p = 0 .. NNZ-1
 * and for i, j >= 1:   J[i][j] = ca[j] * WA_{i-1} + cb[j] * WB_{i-1} + sum_{p: (rows[p], cols[p]) = (i, j)} S_p.
 * The pattern (rows, cols; column-major order) and the column factors ca, cb are fixed per mechanism.
 * Layouts as for the Jacobian: PYJAC_JAC_STATE_MAJOR fac[s*NF + e], PYJAC_JAC_STATE_FASTEST fac[e*ld + s]. */
int pyjac_factored_size(const pyjac_mech* m, int* nf, int* nnz);
/* rows[NNZ], cols[NNZ] (indices into the NSP x NSP Jacobian), ca[NSP], cb[NSP] (entry 0 unused); any may be NULL */
int pyjac_factored_pattern(const pyjac_mech* m, int* rows, int* cols, double* ca, double* cb);
int pyjac_eval_jacob_factored_dev(pyjac_mech* m, int n, const double* d_pres, const double* d_y,
                                  long long y_ss, long long y_sv, double* d_fac, int fac_layout,
                                  long long fac_ld, void* stream);
/* host rows in (as pyjac_eval_jacob_host), one record of NF doubles per state out */
int pyjac_eval_jacob_factored_host(pyjac_mech* m, int n, const double* pres, const double* y, double* fac);
/* out = J v per state straight from the records (J is never formed): element i of state s of v is
 * d_v[s*v_ss + i*v_sv], of the product d_out[s*o_ss + i*o_sv] */
int pyjac_jvp_dev(pyjac_mech* m, int n, const double* d_fac, int fac_layout, long long fac_ld,
                  const double* d_v, long long v_ss, long long v_sv,
                  double* d_out, long long o_ss, long long o_sv, void* stream);
/* The consumer step of an implicit integrator (docs/faqs.rst:113-117): x = (I - gamma J)^-1 rhs per state,
 * the matrix expanded from the record into shared memory, LU with partial pivoting -- the dense Jacobian
 * never reaches HBM.  gamma: one value for all states, or d_gamma[n] (device) when not NULL.  d_info[n]
 * (device, may be NULL): 0, or c + 1 when column c had no pivot (x of that state is left untouched).
 * PYJAC_ETOOBIG when (NSP|1) * NSP + 3 NSP doubles exceed the shared memory of a block (NSP > ~165). */
int pyjac_newton_solve_dev(pyjac_mech* m, int n, const double* d_fac, int fac_layout, long long fac_ld,
                           double gamma, const double* d_gamma,
                           const double* d_rhs, long long r_ss, long long r_sv,
                           double* d_x, long long x_ss, long long x_sv, int* d_info, void* stream);
/* conc[NSP], fwd[FWD_RATES], rev[REV_RATES], pres_mod[PRES_MOD_RATES], spec_rates[NSP] and
 * dy[NSP] per state; any output pointer may be NULL.  Element v of state s of every output
 * goes to out[s*o_ss_mult*width + v] when o_state_fastest == 0 (rows), or out[v*o_ld + s]
 * when o_state_fastest != 0. */
int pyjac_rates_dev(pyjac_mech* m, int n, const double* d_pres, const double* d_y,
                    long long y_ss, long long y_sv, double* d_conc, double* d_fwd,
                    double* d_rev, double* d_pres_mod, double* d_spec_rates, double* d_dy,
                    int o_state_fastest, long long o_ld, void* stream);

/* ---- 2. host-pointer batch API (pyjac/pywrap/pyjacob.cuh:6-10) --------------------- */
/* Selects the mechanism used by surfaces 2 and 3 (the reference bakes it in at build time). */
int pyjac_set_mechanism(pyjac_mech* m);
/* init(num): returns `padded` >= num (the reference pads to its 64-thread block and may
 * cap at 80 % of free memory, pyjacob.cu:97-121; here padded == num rounded up to 8 and
 * large batches are streamed in chunks instead of being capped).  Deviation: does NOT call
 * cudaDeviceReset() (pyjacob.cu:88), which would destroy the caller's CUDA context. */
int pyjac_cu_init(int num);
/* run(): host arrays, flattened Fortran order (variable-major / state-fastest), pitch `num`
 * (functional_tester/test.py:656-660,732-733): mass_frac is NSP x num ([T; Y_0..Y_{NSP-2}]),
 * conc NSP x num, fwd FWD_RATES x num, rev REV_RATES x num, pres_mod PRES_MOD_RATES x num,
 * spec_rates NSP x num, dy NSP x num, jac NSP*NSP x num. */
void pyjac_cu_run(int num, int padded, const double* pres, const double* mass_frac,
                  double* conc, double* fwd_rxn_rates, double* rev_rxn_rates,
                  double* pres_mod, double* spec_rates, double* dy, double* jac);
void pyjac_cu_cleanup(void);
/* Host batch of row-major states (n x NSP) -> row-major Jacobians (n x NSP*NSP), streamed
 * through pinned staging buffers; the path bench.py's e2e figure times. */
int pyjac_eval_jacob_host(pyjac_mech* m, int n, const double* pres, const double* y,
                          double* jac);
int pyjac_dydt_host(pyjac_mech* m, int n, const double* pres, const double* y, double* dy);

/* Registers the table blob of a mechanism as the one surfaces 2 and 3 use when none was selected
 * with pyjac_set_mechanism: the per-mechanism stub library that pyjac_b200.libgen.generate_library
 * writes (libcu_pyjac.so, the name of pyjac/libgen/libgen.py:170-186) calls this from a constructor, so a
 * program linked against it -- the reference's tester.c.in -- needs no set-up call.  The blob is
 * not copied; the mechanism is loaded on the current device at first use. */
int pyjac_register_tables(const void* blob, size_t len);

/* ---- 3. scalar API, reference names ------------------------------------------------ */
/* Re-entrant and safe to call from concurrent host threads (every call stages through its own
 * slot), like the reference's stack-only functions (tester.c.in:24-29 calls eval_jacob inside
 * an OpenMP loop). */
void eval_jacob(const double t, const double pres, const double* y, double* jac);
void dydt(const double t, const double pres, const double* y, double* dy);
void eval_conc(const double T, const double pres, const double* mass_frac, double* y_N,
               double* mw_avg, double* rho, double* conc);
void eval_rxn_rates(const double T, const double pres, const double* C, double* fwd_rxn_rates,
                    double* rev_rxn_rates);
void get_rxn_pres_mod(const double T, const double pres, const double* C, double* pres_mod);
void eval_spec_rates(const double* fwd_rates, const double* rev_rates, const double* pres_mod,
                     double* sp_rates, double* dy_N);
/* chem_utils.h of the emitted library (rate_subs.py:1581-1608; bodies 1806-2086): NSP mass-based
 * enthalpies [J/kg], internal energies, and heat capacities at constant volume / pressure [J/kg/K] */
void eval_h(const double T, double* h);
void eval_u(const double T, double* u);
void eval_cv(const double T, double* cv);
void eval_cp(const double T, double* cp);
/* mass_mole.h / mechanism.h (mech_auxiliary.py:188-206), called by read_initial_conditions.c:29:
 * NSP mass fractions from the mechanism file's species order to pyJac's internal order (last
 * species moved to the end) and back, in place */
void apply_mask(double* y_specs);
void apply_reverse_mask(double* y_specs);

#ifdef __cplusplus
}

/* ---- 2'. pyjac/pywrap/pyjacob.cuh:6-10 verbatim: C++ linkage, as the reference's pyjacob.cu defines them
 * and pyjacob_cuda_wrapper.pyx:5-10 binds them.  init = pyjac_cu_init (exits on error like the
 * reference's cudaErrorCheck), run = pyjac_cu_run, cleanup = pyjac_cu_cleanup. */
void run(int, int, const double*, const double*, double*, double*, double*, double*, double*, double*, double*);
int init(int);
void cleanup();
#endif
#endif
