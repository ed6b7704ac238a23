// Shared definitions of the sm_100a kernels: mechanism-table handles, the I/O descriptor of one
// launch, reaction flag bits (pyjac_b200/tables.py holds the same values) and the exponential.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pj {

enum : int {
    F_REV = 1, F_THD = 2, F_PDEP = 4, F_LOW = 8, F_TROE = 16, F_SRI = 32, F_PMT = 64,
    F_PMT_INJ = 128, F_TROE_T2 = 256, F_SRI5 = 512, F_SRI5_DT = 1024, F_NO_T = 2048,
    F_EFFN1 = 4096, F_HAS_LAST = 1 << 13, F_NEGA = 1 << 14, F_WANT_PMT = 1 << 16, F_EFF_SLOTS = 1 << 17,
    F_PLOG = 1 << 18, F_CHEB = 1 << 19,
    NRE_SHIFT = 20, NPR_SHIFT = 24, NPAR = 32
};

// what a launch of k_eval produces
enum : int { M_JAC = 0,     // eval_jacob: the Jacobian
             M_DYDT = 1,    // dydt
             M_RATES = 2,   // conc / fwd / rev / pres_mod / spec_rates (/ dydt): the rate routines
             M_FACT = 3 };  // the Jacobian in factored form: energy row, T column, the rank-2 factors W_k a_k, W_k b_k
                            // and the sparse block in a fixed pattern, through IO::jac

struct Tables {
    int nsp, nr, nrev, npd, nraw, first_pm, npm;
    double ru;
    const double *sp_w, *sp_iw, *sp_ruw, *sp_tmid, *sp_nasa;
    const double* pm_par;            // [npm][NPAR]
    const int* pm_sp;                // specific collider of a fall-off reaction or -1
    // eval_spec_rates entry point only: per species CSR of (reaction, nu), positions of a
    // reaction in the reference's fwd / rev / pres_mod arrays
    const int *red_off, *red_rx;
    const double* red_nu;
    const int4* rx_out;              // {fwd index, rev index or -1, pres_mod index or -1, 0}
    // PLOG reactions: entries plog_off[p] .. plog_off[p + 1] of plog_par[][8] = {threshold (Pa),
    // ln A, b, Ta, ln P, 1 / (ln P' - ln P), b' - b, Ta' - Ta}; nplog = 0: no such reaction
    int nplog;
    const int* plog_off;
    const double* plog_par;
    // Chebyshev reactions: cheb_par + cheb_off[p] = {n_T, n_P, (tsum, tsub, psum, psub) of the rate,
    // the same of the Jacobian, -2 ln10 / tsub, 0}, n_T x n_P rate coefficients, n_T x n_P
    // coefficients of the temperature derivative (row i times i); ncheb = 0: no such reaction
    int ncheb;
    const int* cheb_off;
    const double* cheb_par;
};

struct IO {
    int n;
    const double* pres;
    const double* y;
    long long y_ss, y_sv;
    int in_conc;     // 1: the input row is [T, C_0 .. C_{NSP-1}] (concentrations given)
    int conv;        // 1 (dydt / rate routines): constant volume -- `pres` holds densities, the energy
                     // equation uses u and cv (rate_subs.py:2340-2485)
    double* jac;
    int jac_layout;
    long long jac_ld;
    double* dy;
    long long dy_ss, dy_sv;
    double *conc, *fwd, *rev, *pm, *sr;     // M_RATES outputs (nullable)
    double* scal3;   // M_RATES: y_N, mw_avg, rho per state (3 doubles, rows), nullable
    int o_sf;
    long long o_ld;
    const char* ws;  // working sets in global memory (one per block) for plans that ask for it, or NULL
    int dbg_skip;    // development only (PYJAC_DEBUG_SKIP): phases to skip when timing
    long long* dbg_clk;   // development only: per-phase cycle counts of block 0 or NULL
};

// out element v of state s for the M_RATES outputs
__device__ __forceinline__ void put(double* base, const IO& io, int width, long long s, int v, double x)
{
    if (io.o_sf) base[(long long)v * io.o_ld + s] = x;
    else base[s * (long long)width + v] = x;
}

__device__ __forceinline__ double log10_clamped(double x) { return log10(fmax(x, 1.0e-300)); }

// exp for |x| <= 708 without the range handling of the library version: Cody-Waite reduction
// to |r| <= ln2/2, degree-12 Taylor polynomial (truncation 1.7e-16 relative), exponent added
// to the high word.  Anything else (overflow, underflow, NaN) takes the library path.
__device__ __forceinline__ double exp_fast(double x)
{
    if (!(fabs(x) <= 708.0)) return exp(x);
    const double t = fma(x, 1.4426950408889634074, 6755399441055744.0);
    const int k = __double2loint(t);
    const double kd = t - 6755399441055744.0;
    double r = fma(kd, -6.93147180369123816490e-01, x);
    r = fma(kd, -1.90821492927058770002e-10, r);
    double p = 2.08767569878680989792e-09;               // 1/12!
    p = fma(p, r, 2.50521083854417187751e-08);
    p = fma(p, r, 2.75573192239858906526e-07);
    p = fma(p, r, 2.75573192239858906526e-06);
    p = fma(p, r, 2.48015873015873015873e-05);
    p = fma(p, r, 1.98412698412698412698e-04);
    p = fma(p, r, 1.38888888888888888889e-03);
    p = fma(p, r, 8.33333333333333333333e-03);
    p = fma(p, r, 4.16666666666666666667e-02);
    p = fma(p, r, 1.66666666666666666667e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// eval_spec_rates (rate_subs.py:1425-1527) from caller-supplied rate arrays: one thread per species.
__global__ void k_spec_rates(const __grid_constant__ Tables tb, const double* fwd, const double* rev,
                             const double* pm, double* sp_rates)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= tb.nsp) return;
    double acc = 0.0;
    for (int e = tb.red_off[k]; e < tb.red_off[k + 1]; ++e) {
        const int4 d = tb.rx_out[tb.red_rx[e]];
        double rate = fwd[d.x];
        if (d.y >= 0) rate -= rev[d.y];
        if (d.z >= 0) rate *= pm[d.z];
        acc += tb.red_nu[e] * rate;
    }
    sp_rates[k] = acc;
}

// ---- finite-difference Jacobian (the reference's comparison build, performance_tester/fd_jacob.cu:23-95)
// All arrays state-fastest with leading dimension n.  k_fd_step: CVODE-style increment of column j,
// r = max(sqrt(eps) |y_j|, r0 / ewt_j) with ewt = ATOL + RTOL |y|, r0 = 1000 RTOL eps NSP fac,
// fac = rms(ewt * dy0), optionally capped; writes the increment and y_tmp = y with y_j + c * r.
__global__ void k_fd_step(int n, int nsp, int j, double c, double r_cap, const double* y, const double* dy0, double* ytmp, double* r_out)
{
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s >= n) return;
    const double atol = 1e-15, rtol = 1e-8, eps = 2.2204460492503131e-16;
    double sum = 0.0;
    for (int i = 0; i < nsp; ++i) {
        const double e = (atol + rtol * fabs(y[(long long)i * n + s])) * dy0[(long long)i * n + s];
        sum += e * e;
    }
    const double r0 = 1000.0 * rtol * eps * nsp * sqrt(sum / nsp);
    const double yj = y[(long long)j * n + s];
    double r = fmax(sqrt(eps) * fabs(yj), r0 / (atol + rtol * fabs(yj)));
    // r0 grows with |dy/dt|: far from equilibrium the reference's increment can exceed the mass
    // fractions themselves, and at equilibrium it vanishes (the difference quotient then amplifies
    // the round-off of dy); r_cap > 0 keeps r within [r_cap / 100, r_cap] * max(|y_j|, 1)
    if (r_cap > 0.0) r = fmax(fmin(r, r_cap * fmax(fabs(yj), 1.0)), 0.01 * r_cap * fmax(fabs(yj), 1.0));
    r_out[s] = r;
    // only row j differs from y: the caller keeps ytmp == y in every other row
    ytmp[(long long)j * n + s] = yj + c * r;
}

// jac[:, j] (+)= w * dy / r   (first = 1: assign)
__global__ void k_fd_accum(int n, int nsp, int j, double w, int first, const double* dy, const double* r, double* jac)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)n * nsp) return;
    const long long i = t / n, s = t % n;
    const double v = w * dy[i * n + s] / r[s];
    double* o = jac + ((long long)j * nsp + i) * n + s;
    *o = first ? v : *o + v;
}

// eval_h / eval_u / eval_cv / eval_cp of the emitted chem_utils (rate_subs.py:1806-1874 h, 1876-1945 u,
// 1947-2019 cv, 2021-2086 cp): NASA-7 polynomials per species, mass based, one thread per species.
// sp_nasa[k][branch] = {a0..a4, a5, a1/2, a2/3, a3/4, a4/5, ., a0 - 1, ...} (tables._nasa_row).
enum : int { TH_H = 0, TH_U = 1, TH_CV = 2, TH_CP = 3 };
__global__ void k_thermo(const __grid_constant__ Tables tb, double T, int what, double* out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= tb.nsp) return;
    const double* c = tb.sp_nasa + (k * 2 + (T <= tb.sp_tmid[k] ? 0 : 1)) * 16;
    const double ruw = tb.sp_ruw[k];
    double v;
    if (what == TH_H) v = ruw * (c[5] + T * (c[0] + T * (c[6] + T * (c[7] + T * (c[8] + c[9] * T)))));
    else if (what == TH_U) v = ruw * (c[5] + T * (c[11] + T * (c[6] + T * (c[7] + T * (c[8] + c[9] * T)))));
    else if (what == TH_CV) v = ruw * (c[11] + T * (c[1] + T * (c[2] + T * (c[3] + c[4] * T))));
    else v = ruw * (c[0] + T * (c[1] + T * (c[2] + T * (c[3] + c[4] * T))));
    out[k] = v;
}

}  // namespace pj
