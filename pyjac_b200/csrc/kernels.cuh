// Data-driven sm_100a kernels: species rates, dydt and the analytical Jacobian of a
// gas-phase kinetic mechanism, evaluated for a batch of states from mechanism *tables*
// (pyjac_b200/tables.py) instead of per-mechanism generated code.
//
// Replaces the emitted library of the reference: eval_conc (rate_subs.py:1626-1706),
// eval_rxn_rates (:254-876), get_rxn_pres_mod (:879-1294), eval_spec_rates (:1297-1542),
// eval_h / eval_cp (:1806-2086), dydt (:2171-2335) and eval_jacob
// (create_jacobian.py:2189-3298).  The arithmetic is regrouped (see tables.py) so that
// each Jacobian element is produced and stored exactly once.
//
// A persistent thread block walks over groups of G states (G = 1 or 2).  Every per-state
// quantity in shared memory is interleaved over the G states of the group (slot-major,
// state-minor), so that one 16-byte shared-memory access serves both states of a group and
// the index decoding of the mechanism tables is done once per group.  Per group, separated by
// three block barriers:
//
//   A   one warp per state: mass fractions -> concentrations and NASA-7 thermo per species
//       (C_k, B_k = Gibbs term of Kc, dB_k/dT, h_k W_k).  Runs for the *next* group while
//       phase E stores the current one.
//   B   one thread per reaction, both states of the group in registers: kf, kr, rates of
//       progress, third-body / fall-off factors and their derivatives -> four scalars per
//       reaction (net rate, T-column term, X1, X2), the reaction enthalpy dH and the non-zero
//       d(rate)/dC values ("raw")
//   C   four lanes per species: sum over its reactions of nu * (the four scalars) -> wdot_k
//       and the dense rank-2 part of the Jacobian (a_k, b_k, T column)
//   D   sparse gather into a dense NSP x NSP tile in shared memory: one thread per
//       structurally non-zero element (1, 2, 4 or 8 contributions, unrolled), four lanes for
//       long lists and for the dH-weighted energy-equation row
//   E   one warp per Jacobian column, lanes = rows: tile + rank-2 part -> HBM, coalesced;
//       warp g < G instead finishes the energy-equation row of state g (five dot products) and
//       then runs phase A of the next group
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pj {

enum : int {
    F_REV = 1, F_THD = 2, F_PDEP = 4, F_LOW = 8, F_TROE = 16, F_SRI = 32, F_PMT = 64,
    F_PMT_INJ = 128, F_TROE_T2 = 256, F_SRI5 = 512, F_SRI5_DT = 1024, F_NO_T = 2048,
    F_EFFN1 = 4096, F_WANT_PMT = 1 << 16, F_EFF_SLOTS = 1 << 17, NRE_SHIFT = 20, NPR_SHIFT = 24,
    NPAR = 32, NONE16 = 0xFFFF
};

enum : int { M_JAC = 1, M_DYDT = 2, M_RATES = 4 };

struct Tables {
    int nsp, nr, nrev, npd, nraw, first_pm, npm;
    int nfix, nq, nq_j;              // sparse entries: fixed-length classes, quad entries
    int d_cls[5], d_ccon[4];
    double ru;
    const double *sp_w, *sp_iw, *sp_ruw, *sp_tmid, *sp_mwf, *sp_nasa;
    const int4* rx_rec;
    const uint4* rx_dst;             // 8 x u16 per reaction
    const double* pm_par;
    const int *pm_sp, *pm_eff_off, *pm_eff_sp;
    const double* pm_eff_am1;
    const int* red_off;
    const unsigned* red_pk;
    const unsigned short* d_dst;
    const unsigned* d_con;
    const unsigned short* q_dst;
    const int* q_off;
    const unsigned* q_con;
    // eval_spec_rates entry point only
    const int* red_rx;
    const double* red_nu;
};

struct IO {
    int n;
    const double* pres;
    const double* y;
    long long y_ss, y_sv;
    int in_conc;     // 1: the input row is [T, C_0 .. C_{NSP-1}] (concentrations given)
    double* jac;
    int jac_layout;
    long long jac_ld;
    double* dy;
    long long dy_ss, dy_sv;
    double *conc, *fwd, *rev, *pm, *sr;     // M_RATES outputs (nullable)
    double* scal3;   // M_RATES: y_N, mw_avg, rho per state (3 doubles, rows), nullable
    int o_sf;
    long long o_ld;
    int dbg_skip;    // development only (PYJAC_DEBUG_SKIP): phases of k_jacobian to skip when timing
    long long* dbg_clk;   // development only: per-phase cycle counts of block 0 (8 slots) or NULL
};

// shared-memory carve-up: offsets in doubles, every array interleaved over the G states
struct Layout {
    int nsp1;                 // padded species count (>= nsp + 1)
    int off_C, off_B, off_dB, off_hW;     // [nsp1][G], slot nsp = empty reaction slot
    int off_wdot, off_sT, off_a, off_b;   // [nsp1][G] per species
    int off_cp;               // [2][nsp1][G]
    int off_y;                // [nsp1][G]
    int off_scal;             // [2][NSCAL][G]
    int off_net, off_tT, off_X1, off_X2, off_rh;   // [nr][G]
    int off_raw;              // [nraw + 2][G]: nraw = zero slot
    int off_tile;             // [nsp*nsp][G]
    int total;
};

enum : int { S_T = 0, S_LOGT, S_IT, S_RHO, S_RHOINV, S_MW, S_M, S_CPAVG, S_WDCP, S_P,
             S_H1, S_NWT, S_NMWR, NSCAL = 16 };

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
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

// G interleaved doubles at p (16-byte vector access for G = 2)
template <int G>
__device__ __forceinline__ void ldv(const double* p, double (&o)[G])
{
    if (G == 2) {
        const double2 t = *reinterpret_cast<const double2*>(p);
        o[0] = t.x;
        o[G - 1] = t.y;
    } else {
        o[0] = *p;
    }
}

template <int G>
__device__ __forceinline__ void stv(double* p, const double (&v)[G])
{
    if (G == 2) *reinterpret_cast<double2*>(p) = make_double2(v[0], v[G - 1]);
    else *p = v[0];
}

// the double whose top 16 bits are the upper half of w (small integers: low 48 bits zero)
__device__ __forceinline__ double coef_of(unsigned w) { return __hiloint2double((int)(w & 0xFFFF0000u), 0); }

// out element v of state s for the M_RATES outputs
__device__ __forceinline__ void put(double* base, const IO& io, int width, long long s, int v, double x)
{
    if (io.o_sf) base[(long long)v * io.o_ld + s] = x;
    else base[s * (long long)width + v] = x;
}

struct RxP {
    double lnA, b, Ta, lnKc;
    int fl, s0, s1, s2, s3, s4, s5, rev_idx, pm_idx, orig;
};

__device__ __forceinline__ RxP load_rx(const int4* rec, int p)
{
    const int4 a = __ldg(rec + p * 4), b = __ldg(rec + p * 4 + 1), c = __ldg(rec + p * 4 + 2),
               d = __ldg(rec + p * 4 + 3);
    RxP r;
    r.lnA = __hiloint2double(a.y, a.x);
    r.b = __hiloint2double(a.w, a.z);
    r.Ta = __hiloint2double(b.y, b.x);
    r.lnKc = __hiloint2double(b.w, b.z);
    r.fl = c.x;
    r.s0 = c.z & 0xFFFF; r.s1 = (unsigned)c.z >> 16;
    r.s2 = c.w & 0xFFFF; r.s3 = (unsigned)c.w >> 16;
    r.s4 = d.x & 0xFFFF; r.s5 = (unsigned)d.x >> 16;
    r.rev_idx = d.y; r.pm_idx = d.z; r.orig = d.w;
    return r;
}

// Everything phase B does for one reaction and the G states of the group.  PM selects the
// third-body / fall-off code.  THREE: some lane of the warp has a third molecule on a side.
template <int G, bool PM, bool JAC, bool RATES>
__device__ __forceinline__ void reaction(const Tables& tb, const IO& io, const Layout& L,
                                         double* __restrict__ smem, int buf, int p, bool valid,
                                         bool three, long long s0)
{
    const int nsp = tb.nsp, last = tb.nsp - 1;
    const RxP rx = load_rx(tb.rx_rec, p);
    const int fl = rx.fl;
    const bool isrev = fl & F_REV;
    const double* sc = smem + L.off_scal + buf * NSCAL * G;
    const double* Cv = smem + L.off_C;
    double T[G], logT[G], iT[G];
    ldv<G>(sc + S_T * G, T);
    ldv<G>(sc + S_LOGT * G, logT);
    ldv<G>(sc + S_IT * G, iT);

    // ---- pressure modification first: PM_, and for the Jacobian gg, Xd, e1Fi
    double PM_[G], gg[G], Xd[G], e1Fi[G];
#pragma unroll
    for (int g = 0; g < G; ++g) { PM_[g] = 1.0; gg[g] = 1.0; Xd[g] = 0.0; e1Fi[g] = 0.0; }
    const int mi = PM ? p - tb.first_pm : 0;
    const double* par = tb.pm_par + mi * NPAR;
    if (PM) {
        double thd[G];
        ldv<G>(sc + S_M * G, thd);
        const int e0 = tb.pm_eff_off[mi], e1_ = tb.pm_eff_off[mi + 1];
        for (int e = e0; e < e1_; ++e) {
            double ce[G];
            ldv<G>(Cv + tb.pm_eff_sp[e] * G, ce);
            const double am1 = tb.pm_eff_am1[e];
#pragma unroll
            for (int g = 0; g < G; ++g) thd[g] += am1 * ce[g];
        }
        if (fl & F_PDEP) {
            const int csp = tb.pm_sp[mi];
            double ct[G];
            if (csp >= 0) ldv<G>(Cv + csp * G, ct);
            else {
#pragma unroll
                for (int g = 0; g < G; ++g) ct[g] = thd[g];
            }
            const bool low = fl & F_LOW;
            const double p0 = par[0], p1 = par[1], p2 = par[2], p3 = par[3];
            double Pr[G], i1p[G], F[G], dpr[G];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const double e1 = exp_fast(p0 + p1 * logT[g] - p2 * iT[g]);
                Pr[g] = ct[g] * e1;
                const double dpr4 = p3 + p2 * iT[g] - 1.0;
                dpr[g] = p1 + p2 * iT[g] - 1.0;
                i1p[g] = 1.0 / (1.0 + Pr[g]);
                if (low) { Xd[g] = dpr4 * iT[g] * i1p[g]; gg[g] = i1p[g]; }
                else { Xd[g] = -Pr[g] * dpr4 * iT[g] * i1p[g]; gg[g] = -Pr[g] * i1p[g]; }
                F[g] = 1.0;
                e1Fi[g] = e1;
            }
            if (fl & F_TROE) {
                const double iln10 = 0.43429448190325182765;
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const double e3 = exp_fast(T[g] / par[7]), e1t = exp_fast(T[g] / par[9]);
                    double Fc = par[6] * e3 + par[8] * e1t;
                    double dF = par[11] * e3 - par[12] * e1t;
                    if (fl & F_TROE_T2) {
                        const double e2 = exp_fast(par[10] * iT[g]);
                        Fc += e2;
                        dF += par[13] * iT[g] * iT[g] * e2;
                    }
                    const double lnFc = log(fmax(Fc, 1.0e-300));
                    const double lF = lnFc * iln10, lP = log10_clamped(Pr[g]);
                    const double A = lP - 0.67 * lF - 0.4;
                    const double Bq = 0.806 - 1.1762 * lF - 0.14 * lP;
                    const double q1 = 1.0 + A * A / (Bq * Bq);
                    const double lnF_AB = 2.0 * lnFc * A / (Bq * Bq * Bq * q1 * q1);
                    F[g] = exp_fast(lnFc / q1);
                    if (JAC) {
                        Xd[g] += (1.0 / (Fc * q1) - lnF_AB * (-0.67 * iln10 * Bq + 1.1762 * iln10 * A) / Fc) * dF
                                 - lnF_AB * (Bq * iln10 + 0.14 * iln10 * A) * dpr[g] * iT[g];
                        gg[g] -= lnF_AB * (Bq * iln10 + A * 0.14 * iln10);
                    }
                }
            } else if (fl & F_SRI) {
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const double lP = log10_clamped(Pr[g]);
                    const double X = 1.0 / (1.0 + lP * lP);
                    F[g] = pow(par[14] * exp(-par[15] * iT[g]) + exp(-T[g] / par[16]), X);
                    if (fl & F_SRI5) F[g] *= par[17] * pow(T[g], par[18]);
                    if (JAC) {
                        const double two_iln10 = 0.86858896380650365530;
                        const double eb = exp(par[23] * iT[g]), ec = exp(T[g] / par[25]);
                        const double den = par[26] * eb + ec;
                        Xd[g] += X * ((par[22] * iT[g] * iT[g] * eb - par[24] * ec) / den
                                      - X * two_iln10 * lP * dpr[g] * log(den) * iT[g]);
                        if (fl & F_SRI5_DT) Xd[g] += par[27] * iT[g];
                        gg[g] -= X * X * two_iln10 * lP * log(par[19] * exp(par[20] * iT[g]) + exp(T[g] / par[21]));
                    }
                }
            }
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const double Fi = F[g] * i1p[g];
                PM_[g] = low ? Fi * Pr[g] : Fi;
                e1Fi[g] *= Fi;
            }
        } else {
#pragma unroll
            for (int g = 0; g < G; ++g) PM_[g] = thd[g];
        }
    }

    // ---- rate constants and rates of progress
    double c0[G], c1[G], c2[G], c3[G], c4[G], c5[G];
    ldv<G>(Cv + rx.s0 * G, c0);
    ldv<G>(Cv + rx.s1 * G, c1);
    ldv<G>(Cv + rx.s3 * G, c3);
    ldv<G>(Cv + rx.s4 * G, c4);
    if (three) { ldv<G>(Cv + rx.s2 * G, c2); ldv<G>(Cv + rx.s5 * G, c5); }
    else {
#pragma unroll
        for (int g = 0; g < G; ++g) { c2[g] = 1.0; c5[g] = 1.0; }
    }
    double kf[G], kr[G], f[G], r[G], net[G];
    {
        const double* Bv = smem + L.off_B;
        double b0[G], b1[G], b3[G], b4[G];
        ldv<G>(Bv + rx.s0 * G, b0);
        ldv<G>(Bv + rx.s1 * G, b1);
        ldv<G>(Bv + rx.s3 * G, b3);
        ldv<G>(Bv + rx.s4 * G, b4);
        double sB[G];
#pragma unroll
        for (int g = 0; g < G; ++g) sB[g] = (b3[g] + b4[g]) - (b0[g] + b1[g]);
        if (three) {
            double b2[G], b5[G];
            ldv<G>(Bv + rx.s2 * G, b2);
            ldv<G>(Bv + rx.s5 * G, b5);
#pragma unroll
            for (int g = 0; g < G; ++g) sB[g] += b5[g] - b2[g];
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const double lnkf = rx.lnA + rx.b * logT[g] - rx.Ta * iT[g];
            kf[g] = exp_fast(lnkf);
            const double krv = exp_fast(lnkf - sB[g] - rx.lnKc);
            kr[g] = isrev ? krv : 0.0;
            f[g] = kf[g] * c0[g] * c1[g] * c2[g];
            r[g] = kr[g] * c3[g] * c4[g] * c5[g];
            net[g] = f[g] - r[g];
        }
    }
    double pmt[G];
#pragma unroll
    for (int g = 0; g < G; ++g) pmt[g] = (PM && (fl & F_PMT)) ? gg[g] * net[g] : 0.0;

    if (RATES && valid) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const long long s = s0 + g;
            if (s >= io.n) continue;
            if (io.fwd) put(io.fwd, io, tb.nr, s, rx.orig, f[g]);
            if (io.rev && isrev) put(io.rev, io, tb.nrev, s, rx.rev_idx, r[g]);
            if (PM && io.pm) put(io.pm, io, tb.npd, s, rx.pm_idx, PM_[g]);
        }
    }
    if (!JAC) {
        if (valid) {
            double v[G];
#pragma unroll
            for (int g = 0; g < G; ++g) v[g] = net[g] * PM_[g];
            stv<G>(smem + L.off_net + p * G, v);
        }
        return;
    }

    // ---- Jacobian scalars
    const double nre = (double)((fl >> NRE_SHIFT) & 15), npr = (double)((fl >> NPR_SHIFT) & 15);
    double rho_inv[G], nmwr[G];
    ldv<G>(sc + S_RHOINV * G, rho_inv);
    ldv<G>(sc + S_NMWR * G, nmwr);
    double sdB[G], dH[G];
    {
        const double* Dv = smem + L.off_dB;
        const double* Hv = smem + L.off_hW;
        double a0[G], a1[G], a3[G], a4[G];
        ldv<G>(Dv + rx.s0 * G, a0);
        ldv<G>(Dv + rx.s1 * G, a1);
        ldv<G>(Dv + rx.s3 * G, a3);
        ldv<G>(Dv + rx.s4 * G, a4);
#pragma unroll
        for (int g = 0; g < G; ++g) sdB[g] = (a3[g] + a4[g]) - (a0[g] + a1[g]);
        ldv<G>(Hv + rx.s0 * G, a0);
        ldv<G>(Hv + rx.s1 * G, a1);
        ldv<G>(Hv + rx.s3 * G, a3);
        ldv<G>(Hv + rx.s4 * G, a4);
#pragma unroll
        for (int g = 0; g < G; ++g) dH[g] = (a3[g] + a4[g]) - (a0[g] + a1[g]);
        if (three) {
            ldv<G>(Dv + rx.s2 * G, a0);
            ldv<G>(Dv + rx.s5 * G, a1);
            ldv<G>(Hv + rx.s2 * G, a3);
            ldv<G>(Hv + rx.s5 * G, a4);
#pragma unroll
            for (int g = 0; g < G; ++g) { sdB[g] += a1[g] - a0[g]; dH[g] += a4[g] - a3[g]; }
        }
    }
    const double extra = (PM && (fl & F_EFFN1)) ? 1.0 : 0.0;
    double tT[G], X1[G], X2[G], pk[G], pr[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const double dk = rx.b + rx.Ta * iT[g];
        // irreversible: r = 0 makes this f * (dk + 1 - nre)          (cj:1461-1523)
        const double elem = net[g] * dk + f[g] * (1.0 - nre) - r[g] * ((1.0 - npr) - T[g] * sdB[g]);
        double t;
        if (PM) {
            if (fl & F_PDEP) t = (PM_[g] * Xd[g] * net[g] + PM_[g] * iT[g] * elem) * rho_inv[g];
            else t = (-PM_[g] * net[g] * iT[g] + PM_[g] * iT[g] * elem) * rho_inv[g];
        } else {
            t = iT[g] * elem * rho_inv[g];
        }
        tT[g] = (fl & F_NO_T) ? 0.0 : t;
        double inner = (nre + extra) * f[g] - (npr + extra) * r[g];
        if (PM && (fl & F_PMT_INJ)) inner += pmt[g];
        const double jy = nmwr[g] * PM_[g] * inner;
        if (PM && (fl & F_PMT_INJ)) pmt[g] *= e1Fi[g];
        X1[g] = jy;
        X2[g] = -jy;
        if (PM) { X1[g] += par[5] * pmt[g]; X2[g] -= par[4] * pmt[g]; }
        pk[g] = PM_[g] * kf[g];
        pr[g] = -PM_[g] * kr[g];
    }

    // ---- d(rate)/dC values: to their raw slots, or folded into X2 for the last species
    const uint4 dq = __ldg(tb.rx_dst + p);
    double* raw = smem + L.off_raw;
    double d[G];
#define PJ_EMIT(SLOT, DST, EXPR)                                               \
    if ((SLOT) != nsp) {                                                       \
        _Pragma("unroll") for (int g = 0; g < G; ++g) d[g] = (EXPR);           \
        if ((SLOT) == last) {                                                  \
            _Pragma("unroll") for (int g = 0; g < G; ++g) X2[g] -= d[g];       \
        } else if (valid) {                                                    \
            stv<G>(raw + (int)(DST) * G, d);                                   \
        }                                                                      \
    }
    PJ_EMIT(rx.s0, dq.x & 0xFFFFu, pk[g] * c1[g] * c2[g])
    PJ_EMIT(rx.s1, dq.x >> 16, pk[g] * c0[g] * c2[g])
    if (three) { PJ_EMIT(rx.s2, dq.y & 0xFFFFu, pk[g] * c0[g] * c1[g]) }
    if (isrev) {
        PJ_EMIT(rx.s3, dq.y >> 16, pr[g] * c4[g] * c5[g])
        PJ_EMIT(rx.s4, dq.z & 0xFFFFu, pr[g] * c3[g] * c5[g])
        if (three) { PJ_EMIT(rx.s5, dq.z >> 16, pr[g] * c3[g] * c4[g]) }
    }
#undef PJ_EMIT
    if (PM && valid) {
        if (fl & F_EFF_SLOTS) {
            int rb = dq.w & 0xFFFFu;
            const int e0 = tb.pm_eff_off[mi], e1_ = tb.pm_eff_off[mi + 1];
            for (int e = e0; e < e1_; ++e)
                if (tb.pm_eff_sp[e] != last) {
                    const double am1 = tb.pm_eff_am1[e];
#pragma unroll
                    for (int g = 0; g < G; ++g) d[g] = pmt[g] * am1;
                    stv<G>(raw + rb * G, d);
                    ++rb;
                }
        }
        if (fl & F_WANT_PMT) stv<G>(raw + (int)(dq.w >> 16) * G, pmt);
    }
    if (valid) {
        double v[G];
#pragma unroll
        for (int g = 0; g < G; ++g) v[g] = net[g] * PM_[g];
        stv<G>(smem + L.off_net + p * G, v);
        stv<G>(smem + L.off_tT + p * G, tT);
        stv<G>(smem + L.off_X1 + p * G, X1);
        stv<G>(smem + L.off_X2 + p * G, X2);
        stv<G>(smem + L.off_rh + p * G, dH);
    }
}

// Phase D, fixed-length class: LEN contributions of one tile element for the G states.
template <int G, int LEN>
__device__ __forceinline__ void gather_fixed(const unsigned* __restrict__ cw, const double* __restrict__ raw,
                                             double* __restrict__ dst)
{
    unsigned w[LEN];
    if (LEN == 8) {
        const uint4 a = __ldg((const uint4*)cw), b = __ldg((const uint4*)cw + 1);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
        w[LEN > 4 ? 4 : 0] = b.x; w[LEN > 4 ? 5 : 0] = b.y; w[LEN > 4 ? 6 : 0] = b.z; w[LEN > 4 ? 7 : 0] = b.w;
    } else if (LEN == 4) {
        const uint4 a = __ldg((const uint4*)cw);
        w[0] = a.x; w[LEN > 1 ? 1 : 0] = a.y; w[LEN > 2 ? 2 : 0] = a.z; w[LEN > 2 ? 3 : 0] = a.w;
    } else if (LEN == 2) {
        const uint2 a = __ldg((const uint2*)cw);
        w[0] = a.x; w[LEN > 1 ? 1 : 0] = a.y;
    } else {
        w[0] = __ldg(cw);
    }
    double acc[G];
#pragma unroll
    for (int g = 0; g < G; ++g) acc[g] = 0.0;
#pragma unroll
    for (int i = 0; i < LEN; ++i) {
        double v[G];
        ldv<G>(raw + (w[i] & 0xFFFFu) * G, v);
        const double cf = coef_of(w[i]);
#pragma unroll
        for (int g = 0; g < G; ++g) acc[g] = fma(cf, v[g], acc[g]);
    }
    stv<G>(dst, acc);
}

// MINB = 1: up to 512 threads and 128 registers per thread; MINB = 2: up to 384 threads and
// 80 registers, so that two blocks share an SM.
template <int G, int MODE, int MINB>
__global__ void __launch_bounds__(MINB == 1 ? 512 : 384, MINB)
k_eval(const __grid_constant__ Tables tb, const __grid_constant__ IO io,
       const __grid_constant__ Layout L)
{
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int nsp = tb.nsp, last = tb.nsp - 1;
    constexpr bool JAC = (MODE & M_JAC) != 0;
    constexpr bool RATES = (MODE & M_RATES) != 0;

    // ---- phase A for state g of the group starting at s0, into buffer `buf` (one warp)
    auto phase_a = [&](long long s0, int buf, int g) {
        const bool live = s0 + g < io.n;
        const long long s = live ? s0 + g : (long long)io.n - 1;
        const double* ys = io.y + s * io.y_ss;
        const double T = ys[0];
        const double P = io.pres[s];
        double* Yv = smem + L.off_y;
        double sumY = 0.0, sumYW = 0.0;
        double mw_avg, rho;
        if (io.in_conc) {
            // concentrations supplied (eval_rxn_rates / get_rxn_pres_mod entry points):
            // rho = sum C_k W_k, Y_k = C_k W_k / rho
            for (int k = lane; k < nsp; k += 32) {
                const double Ck = ys[(long long)(k + 1) * io.y_sv];
                sumY += Ck;
                sumYW += Ck * tb.sp_w[k];
            }
            sumY = warp_sum(sumY);
            rho = warp_sum(sumYW);
            mw_avg = rho / sumY;
            for (int k = lane; k < nsp; k += 32)
                Yv[k * G + g] = ys[(long long)(k + 1) * io.y_sv] * tb.sp_w[k] / rho;
        } else {
            for (int k = lane; k < last; k += 32) {
                const double Yk = ys[(long long)(k + 1) * io.y_sv];
                Yv[k * G + g] = Yk;
                sumY += Yk;
                sumYW += Yk * tb.sp_iw[k];
            }
            sumY = warp_sum(sumY);
            sumYW = warp_sum(sumYW);
            const double yN = 1.0 - sumY;
            sumYW += yN * tb.sp_iw[last];
            mw_avg = 1.0 / sumYW;
            rho = P * mw_avg / (tb.ru * T);
            if (lane == 0) Yv[last * G + g] = yN;
        }
        __syncwarp();
        const double logT = log(T), iT = 1.0 / T;
        double cpavg = 0.0, wdcp = 0.0;
        double* cpv = smem + L.off_cp + buf * L.nsp1 * G;
        for (int k = lane; k < nsp; k += 32) {
            const double Yk = Yv[k * G + g];
            const double ck = io.in_conc ? ys[(long long)(k + 1) * io.y_sv] : rho * Yk * tb.sp_iw[k];
            if (RATES && io.conc && live) put(io.conc, io, nsp, s, k, ck);
            const double* c = tb.sp_nasa + (k * 2 + (T <= tb.sp_tmid[k] ? 0 : 1)) * 16;
            const double ruw = tb.sp_ruw[k];
            const double cp = ruw * (c[0] + T * (c[1] + T * (c[2] + T * (c[3] + c[4] * T))));
            const double hh = c[6] + T * (c[7] + T * (c[8] + c[9] * T));
            const double h = ruw * (c[5] + T * (c[0] + T * hh));
            cpv[k * G + g] = cp;
            cpavg += Yk * cp;
            double dB = 0.0;
            if (JAC) {
                const double dcp = ruw * (c[1] + T * (2.0 * c[2] + T * (3.0 * c[3] + 4.0 * c[4] * T)));
                wdcp += Yk * dcp;
                dB = (c[11] + c[5] * iT) * iT + hh;
            }
            const double Bk = c[10] + c[11] * logT + T * (c[6] + T * (c[12] + T * (c[13] + c[14] * T))) - c[5] * iT;
            smem[L.off_C + k * G + g] = ck;
            smem[L.off_B + k * G + g] = Bk;
            smem[L.off_dB + k * G + g] = dB;
            smem[L.off_hW + k * G + g] = h * tb.sp_w[k];
        }
        cpavg = warp_sum(cpavg);
        if (JAC) wdcp = warp_sum(wdcp);
        if (lane == 0) {
            smem[L.off_C + nsp * G + g] = 1.0;          // empty reaction slot
            smem[L.off_B + nsp * G + g] = 0.0;
            smem[L.off_dB + nsp * G + g] = 0.0;
            smem[L.off_hW + nsp * G + g] = 0.0;
            double* sc = smem + L.off_scal + buf * NSCAL * G + g;
            const double rho_inv = 1.0 / rho;
            sc[S_T * G] = T; sc[S_LOGT * G] = logT; sc[S_IT * G] = iT; sc[S_RHO * G] = rho;
            sc[S_RHOINV * G] = rho_inv; sc[S_MW * G] = mw_avg; sc[S_M * G] = P / (tb.ru * T);
            sc[S_CPAVG * G] = cpavg; sc[S_WDCP * G] = wdcp; sc[S_P * G] = P; sc[S_NWT * G] = -1.0 / cpavg;
            sc[S_NMWR * G] = -mw_avg * rho_inv;
            if (RATES && io.scal3 && live) {
                double* o = io.scal3 + s * 3;
                o[0] = Yv[last * G + g]; o[1] = mw_avg; o[2] = rho;
            }
        }
    };

    if (JAC) {
        // structurally zero tile elements are never written again
        for (int e = tid; e < nsp * nsp * G; e += blockDim.x) smem[L.off_tile + e] = 0.0;
        for (int e = tid; e < 2 * G; e += blockDim.x) smem[L.off_raw + tb.nraw * G + e] = 0.0;
    }

    const long long ngroups = ((long long)io.n + G - 1) / G;
    int buf = 0;
    if ((long long)blockIdx.x < ngroups)
        for (int g = warp; g < G; g += nwarps) phase_a((long long)blockIdx.x * G, 0, g);
    __syncthreads();

    for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x, buf ^= 1) {
        const long long s0 = grp * G;
        const double* sc = smem + L.off_scal + buf * NSCAL * G;

        // ------------------------------------------------------------ phase B
        // pressure-modified reactions first (longest), then the plain ones; whole warps
        {
            const int npm_pad = (tb.npm + 31) & ~31, n_plain = tb.first_pm;
            const int items = npm_pad + ((n_plain + 31) & ~31);
            for (int base = warp * 32; base < items; base += blockDim.x) {
                const int it = base + lane;
                if (base < npm_pad) {
                    const bool valid = it < tb.npm;
                    reaction<G, true, JAC, RATES>(tb, io, L, smem, buf, tb.first_pm + (valid ? it : 0), valid, true, s0);
                } else {
                    const int q = it - npm_pad;
                    const bool valid = q < n_plain;
                    const int p = valid ? q : 0;
                    const int4 c = __ldg(tb.rx_rec + p * 4 + 2);
                    const int4 dd = __ldg(tb.rx_rec + p * 4 + 3);
                    const bool has3 = ((c.w & 0xFFFF) != nsp) || (((unsigned)dd.x >> 16) != (unsigned)nsp);
                    const bool three = __any_sync(0xffffffffu, has3);
                    reaction<G, false, JAC, RATES>(tb, io, L, smem, buf, p, valid, three, s0);
                }
            }
        }
        __syncthreads();

        // ------------------------------------------------------------ phase C: four lanes per species
        {
            const int q = lane & 3;
            for (int kb = (tid >> 2); kb < ((nsp + 7) & ~7); kb += blockDim.x >> 2) {
                const int k = kb < nsp ? kb : nsp - 1;
                const int e0 = __ldg(tb.red_off + k), e1 = __ldg(tb.red_off + k + 1);
                const int len = kb < nsp ? e1 - e0 : 0;
                int mx = len;
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                double aN[G], aT[G], a1[G], a2[G];
#pragma unroll
                for (int g = 0; g < G; ++g) { aN[g] = 0.0; aT[g] = 0.0; a1[g] = 0.0; a2[g] = 0.0; }
                for (int i = q; i < mx; i += 4) {
                    if (i < len) {
                        const unsigned w = __ldg(tb.red_pk + e0 + i);
                        const int p = w & 0xFFFFu;
                        const double cf = coef_of(w);
                        double v[G];
                        ldv<G>(smem + L.off_net + p * G, v);
#pragma unroll
                        for (int g = 0; g < G; ++g) aN[g] = fma(cf, v[g], aN[g]);
                        if (JAC) {
                            ldv<G>(smem + L.off_tT + p * G, v);
#pragma unroll
                            for (int g = 0; g < G; ++g) aT[g] = fma(cf, v[g], aT[g]);
                            ldv<G>(smem + L.off_X1 + p * G, v);
#pragma unroll
                            for (int g = 0; g < G; ++g) a1[g] = fma(cf, v[g], a1[g]);
                            ldv<G>(smem + L.off_X2 + p * G, v);
#pragma unroll
                            for (int g = 0; g < G; ++g) a2[g] = fma(cf, v[g], a2[g]);
                        }
                    }
                }
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    aN[g] += __shfl_xor_sync(0xffffffffu, aN[g], 1);
                    aN[g] += __shfl_xor_sync(0xffffffffu, aN[g], 2);
                    if (JAC) {
                        aT[g] += __shfl_xor_sync(0xffffffffu, aT[g], 1);
                        aT[g] += __shfl_xor_sync(0xffffffffu, aT[g], 2);
                        a1[g] += __shfl_xor_sync(0xffffffffu, a1[g], 1);
                        a1[g] += __shfl_xor_sync(0xffffffffu, a1[g], 2);
                        a2[g] += __shfl_xor_sync(0xffffffffu, a2[g], 1);
                        a2[g] += __shfl_xor_sync(0xffffffffu, a2[g], 2);
                    }
                }
                if (q == 0 && kb < nsp) {
                    stv<G>(smem + L.off_wdot + k * G, aN);
                    if (JAC) {
                        double mw[G], ri[G];
                        ldv<G>(sc + S_MW * G, mw);
                        ldv<G>(sc + S_RHOINV * G, ri);
#pragma unroll
                        for (int g = 0; g < G; ++g) {
                            const double comp = aN[g] * (mw[g] * ri[g]);
                            a1[g] += comp;
                            a2[g] -= comp;
                        }
                        stv<G>(smem + L.off_sT + k * G, aT);
                        stv<G>(smem + L.off_a + k * G, a1);
                        stv<G>(smem + L.off_b + k * G, a2);
                    }
                }
            }
        }

        // ------------------------------------------------------------ phase D
        if (JAC) {
            const double* raw = smem + L.off_raw;
            double* tile = smem + L.off_tile;
            for (int e = tid; e < tb.d_cls[1]; e += blockDim.x)
                gather_fixed<G, 8>(tb.d_con + tb.d_ccon[0] + e * 8, raw, tile + (int)__ldg(tb.d_dst + e) * G);
            for (int e = tb.d_cls[1] + tid; e < tb.d_cls[2]; e += blockDim.x)
                gather_fixed<G, 4>(tb.d_con + tb.d_ccon[1] + (e - tb.d_cls[1]) * 4, raw, tile + (int)__ldg(tb.d_dst + e) * G);
            for (int e = tb.d_cls[2] + tid; e < tb.d_cls[3]; e += blockDim.x)
                gather_fixed<G, 2>(tb.d_con + tb.d_ccon[2] + (e - tb.d_cls[2]) * 2, raw, tile + (int)__ldg(tb.d_dst + e) * G);
            for (int e = tb.d_cls[3] + tid; e < tb.d_cls[4]; e += blockDim.x)
                gather_fixed<G, 1>(tb.d_con + tb.d_ccon[3] + (e - tb.d_cls[3]), raw, tile + (int)__ldg(tb.d_dst + e) * G);
            // quad entries: four lanes per element, lists padded to multiples of 16 words
            const int q = lane & 3;
            const double* rh = smem + L.off_rh;
            for (int eb = (tid >> 2); eb < ((tb.nq + 7) & ~7); eb += blockDim.x >> 2) {
                const bool on = eb < tb.nq;
                const int e = on ? eb : 0;
                const int o0 = __ldg(tb.q_off + e), o1 = on ? __ldg(tb.q_off + e + 1) : o0;
                const bool trow = e >= tb.nq_j;
                double acc[G];
#pragma unroll
                for (int g = 0; g < G; ++g) acc[g] = 0.0;
                for (int o = o0 + q * 4; o < o1; o += 16) {
                    const uint4 ww = __ldg((const uint4*)(tb.q_con + o));
                    const unsigned w[4] = {ww.x, ww.y, ww.z, ww.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        double v[G], cf[G];
                        ldv<G>(raw + (w[i] & 0xFFFFu) * G, v);
                        if (trow) ldv<G>(rh + (w[i] >> 16) * G, cf);
                        else {
#pragma unroll
                            for (int g = 0; g < G; ++g) cf[g] = coef_of(w[i]);
                        }
#pragma unroll
                        for (int g = 0; g < G; ++g) acc[g] = fma(cf[g], v[g], acc[g]);
                    }
                }
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    acc[g] += __shfl_xor_sync(0xffffffffu, acc[g], 1);
                    acc[g] += __shfl_xor_sync(0xffffffffu, acc[g], 2);
                }
                if (on && q == 0) stv<G>(tile + (int)__ldg(tb.q_dst + e) * G, acc);
            }
        }
        __syncthreads();

        // ------------------------------------------------------------ phase E  (|| row 0, A of next group)
        const long long next = grp + gridDim.x;
        if (warp < G) {
            const int g = warp;
            const long long s = s0 + g;
            // five dot products over the species: H1 (dydt[0]) and the energy-equation row
            double H1 = 0.0, HA = 0.0, HB = 0.0, HT = 0.0, SCP = 0.0;
            const double* cpv = smem + L.off_cp + buf * L.nsp1 * G;
            for (int k = lane; k < nsp; k += 32) {
                const double hW = smem[L.off_hW + k * G + g], wd = smem[L.off_wdot + k * G + g];
                H1 += hW * wd;
                if (JAC) {
                    HT += hW * smem[L.off_sT + k * G + g];
                    HA += hW * smem[L.off_a + k * G + g];
                    HB += hW * smem[L.off_b + k * G + g];
                    SCP += cpv[k * G + g] * __ldg(tb.sp_w + k) * wd;
                }
            }
            H1 = warp_sum(H1);
            const double rho = sc[S_RHO * G + g], cpavg = sc[S_CPAVG * G + g], rho_inv = sc[S_RHOINV * G + g];
            if (JAC) {
                HA = warp_sum(HA); HB = warp_sum(HB); HT = warp_sum(HT); SCP = warp_sum(SCP);
                if (s < io.n) {
                    // energy-equation row (cj:3095-3254) and jac[0] (cj:1853-1905)
                    const double nwt = sc[S_NWT * G + g];
                    const double XT = H1 / (rho * cpavg * cpavg);
                    const double A0 = nwt * HA, B0 = nwt * HB;
                    const bool sf = io.jac_layout != 0;
                    double* base = sf ? io.jac + s : io.jac + s * (long long)(nsp * nsp);
                    const long long es = sf ? io.jac_ld : 1;
                    const double* tile = smem + L.off_tile;
                    const double cpl = cpv[last * G + g];
                    for (int col = lane; col < nsp; col += 32) {
                        double v;
                        if (col == 0) {
                            v = -(-sc[S_WDCP * G + g] / cpavg * H1 + SCP + HT * rho) / (rho * cpavg);
                        } else {
                            const int j = col - 1;
                            const double iwj = __ldg(tb.sp_iw + j), mwfj = __ldg(tb.sp_mwf + j);
                            v = iwj * (A0 + B0 * mwfj + nwt * tile[(col * nsp) * G + g])
                                + XT * (cpv[j * G + g] - cpl);
                        }
                        base[(long long)(col * nsp) * es] = v;
                    }
                }
            }
            // rates / dydt outputs of this state
            if ((RATES || (MODE & M_DYDT)) && s < io.n) {
                for (int k = lane; k < nsp; k += 32) {
                    const double wd = smem[L.off_wdot + k * G + g];
                    if (RATES && io.sr) put(io.sr, io, nsp, s, k, wd);
                    if (io.dy && k < last) {
                        const double v = wd * __ldg(tb.sp_w + k) * rho_inv;
                        if (RATES) put(io.dy, io, nsp, s, k + 1, v);
                        else io.dy[s * io.dy_ss + (long long)(k + 1) * io.dy_sv] = v;
                    }
                }
                if (lane == 0 && io.dy) {
                    const double v = -1.0 / (rho * cpavg) * H1;
                    if (RATES) put(io.dy, io, nsp, s, 0, v);
                    else io.dy[s * io.dy_ss] = v;
                }
            }
            if (next < ngroups) phase_a(next * G, buf ^ 1, g);
        } else if (JAC) {
            // columns: warp -> (row chunk, column subset); lanes = species rows k (output row k+1)
            const int nchunks = (last + 31) >> 5;
            const int ew = warp - G, enw = nwarps - G;
            const bool sf = io.jac_layout != 0;
            const double* tile = smem + L.off_tile;
            auto columns = [&](int ch, int c0, int cstep) {
                const int k = ch * 32 + lane;
                const bool on = k < last;
                const int kk = on ? k : 0;
                const double wk = __ldg(tb.sp_w + kk);
                double WA[G], WB[G], WT[G];
                ldv<G>(smem + L.off_a + kk * G, WA);
                ldv<G>(smem + L.off_b + kk * G, WB);
                ldv<G>(smem + L.off_sT + kk * G, WT);
#pragma unroll
                for (int g = 0; g < G; ++g) { WA[g] *= wk; WB[g] *= wk; WT[g] *= wk; }
                for (int col = c0; col < nsp; col += cstep) {
                    double v[G];
                    if (col == 0) {
#pragma unroll
                        for (int g = 0; g < G; ++g) v[g] = WT[g];
                    } else {
                        const int j = col - 1;
                        const double iwj = __ldg(tb.sp_iw + j), mwfj = __ldg(tb.sp_mwf + j);
                        double S[G];
                        ldv<G>(tile + (col * nsp + kk + 1) * G, S);
#pragma unroll
                        for (int g = 0; g < G; ++g) v[g] = iwj * (WA[g] + WB[g] * mwfj + wk * S[g]);
                    }
                    if (on) {
#pragma unroll
                        for (int g = 0; g < G; ++g) {
                            const long long s = s0 + g;
                            if (s < io.n) {
                                if (sf) io.jac[(long long)(col * nsp + k + 1) * io.jac_ld + s] = v[g];
                                else io.jac[s * (long long)(nsp * nsp) + col * nsp + k + 1] = v[g];
                            }
                        }
                    }
                }
            };
            if (enw >= nchunks) {
                const int ch = ew % nchunks;
                columns(ch, ew / nchunks, (enw - ch + nchunks - 1) / nchunks);
            } else {
                for (int ch = ew; ch < nchunks; ch += enw) columns(ch, 0, 1);
            }
        }
        __syncthreads();
    }
}

// ---- small kernels behind the reference-named scalar entry points ---------------------

// eval_spec_rates (rate_subs.py:1425-1527): one thread per species.
__global__ void k_spec_rates(const __grid_constant__ Tables tb, const double* fwd, const double* rev,
                             const double* pm, double* sp_rates)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= tb.nsp) return;
    double acc = 0.0;
    for (int e = tb.red_off[k]; e < tb.red_off[k + 1]; ++e) {
        const int p = tb.red_rx[e];
        const int4 d = tb.rx_rec[p * 4 + 3];
        double rate = fwd[d.w];
        if (d.y >= 0) rate -= rev[d.y];
        if (d.z >= 0) rate *= pm[d.z];
        acc += tb.red_nu[e] * rate;
    }
    sp_rates[k] = acc;
}

}  // namespace pj
