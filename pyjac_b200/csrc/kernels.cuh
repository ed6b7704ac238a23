// Data-driven sm_100a kernels: species rates, dydt and the analytical Jacobian of a
// gas-phase kinetic mechanism, evaluated for a batch of states from mechanism *tables*
// (pyjac_b200/tables.py) instead of per-mechanism generated code.
//
// Replaces the emitted library of the reference: eval_conc (rate_subs.py:1626-1706),
// eval_rxn_rates (:254-876), get_rxn_pres_mod (:879-1294), eval_spec_rates (:1297-1542),
// eval_h / eval_cp (:1806-2086), dydt (:2171-2335) and eval_jacob
// (create_jacobian.py:2189-3298).  The arithmetic is regrouped (see tables.py) so that
// each Jacobian element is produced and stored exactly once.  A persistent thread block
// walks over groups of G states; per group, separated by block barriers:
//
//   A   one warp per state: mass fractions -> concentrations and NASA-7 thermo; per species
//       {C_k, B_k (Gibbs term of Kc), dB_k/dT, h_k W_k} as one 32-byte record in shared
//       memory (runs for the *next* group while phase E stores the current one)
//   B   one thread per reaction (x G states; one thread per (reaction, state) for the
//       pressure-dependent ones): kf, kr = exp(ln kf - sum nu B - ...), rates of progress,
//       third-body / fall-off factors and derivatives -> 4 scalars per reaction (R4), the
//       reaction enthalpy dH, and the non-zero d(rate)/dC values ("raw")
//   C1  species reductions, first level: chunks of 8 (reaction, nu) pairs -> partial sums
//   D   sparse gather: one thread per sub-entry (1, 2, 4 or 8 contributions, unrolled) of a
//       structurally non-zero Jacobian element (scaled by W_k), and of the energy row's
//       sparse part (dH-weighted, scaled by -1/cp_avg)
//   C2  one warp per state: per species sum of its chunk partials -> wdot_k and the row
//       vectors rowT / rowA / rowB (row 0 = energy equation, from five dot products: the only
//       warp-shuffle reductions left); other warps sum the entries cut into sub-entries
//   E   one warp per Jacobian column, lanes = rows:
//       out[r] = (rowA[r] + rowB[r] W_j/W_N + S[jmap[j][r]]) / W_j, coalesced stores to HBM
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pj {

enum : int {
    F_REV = 1, F_THD = 2, F_PDEP = 4, F_LOW = 8, F_TROE = 16, F_SRI = 32, F_PMT = 64,
    F_PMT_INJ = 128, F_TROE_T2 = 256, F_SRI5 = 512, F_SRI5_DT = 1024, F_NO_T = 2048,
    F_EFFN1 = 4096, F_WANT_PMT = 1 << 16, F_EFF_SLOTS = 1 << 17, NRE_SHIFT = 20, NPR_SHIFT = 24,
    NPAR = 32, RCH = 8
};

enum : int { M_JAC = 1, M_DYDT = 2, M_RATES = 4 };

struct Tables {
    int nsp, nr, nrev, npd, nraw, nsub, ncon, ncoef, first_pm, npm, nsub_j, nsplit, nchunk, zero_slot;
    double ru;
    const double *sp_w, *sp_iw, *sp_ruw, *sp_tmid, *sp_mwf, *sp_nasa;
    const int4* rx_rec;
    const double* pm_par;
    const int *pm_sp, *pm_eff_off, *pm_eff_sp;
    const double* pm_eff_am1;
    const int* chk_rx;
    const double* chk_nu;
    const int* sp_chk_off;
    int cls_sub[9], cls_con[8];   // sparse sub-entry classes J8 J4 J2 J1 T8 T4 T2 T1
    const unsigned* con;
    const double* sub_w;
    const int *cmb_off, *cmb_idx;
    const unsigned short* jmap;
    // eval_spec_rates entry point only
    const int *red_off, *red_rx;
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
};

// shared-memory carve-up (offsets and per-state strides in doubles), filled on the host
struct Layout {
    int nsp1;                 // padded species count (>= nsp + 1, even)
    int off_spv, st_spv;      // double4 {conc, B, dB, hW} per species
    int off_vec, st_vec;      // wdot, tcol, Ap, Bp
    int off_cp, st_cp;        // [buf][g][nsp1]
    int off_y, st_y;
    int off_scal;             // [buf][g][NSCAL]
    int off_r4, st_r4;        // double4 per reaction
    int off_rh, st_rh;
    int off_raw, st_raw;
    int off_part, st_part;    // double4 per chunk
    int off_sval, st_sval;
    int total;
};

enum : int { V_WDOT = 0, V_ROWT, V_ROWA, V_ROWB, NVEC };   // ROW*: indexed by output row r
enum : int { S_T = 0, S_LOGT, S_IT, S_RHO, S_RHOINV, S_MW, S_M, S_CPAVG, S_WDCP, S_P,
             S_H1, S_XT, S_NWT, NSCAL = 16 };

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double log10_clamped(double x) { return log10(fmax(x, 1.0e-300)); }

// out element v of state s for the M_RATES outputs
__device__ __forceinline__ void put(double* base, const IO& io, int width, long long s, int v, double x)
{
    if (io.o_sf) base[(long long)v * io.o_ld + s] = x;
    else base[s * (long long)width + v] = x;
}

struct RxP {
    double lnA, b, Ta, lnKc;
    int fl, rbase, s0, s1, s2, s3, s4, s5, rev_idx, pm_idx, orig;
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
    r.fl = c.x; r.rbase = c.y;
    r.s0 = c.z & 0xFFFF; r.s1 = (unsigned)c.z >> 16;
    r.s2 = c.w & 0xFFFF; r.s3 = (unsigned)c.w >> 16;
    r.s4 = d.x & 0xFFFF; r.s5 = (unsigned)d.x >> 16;
    r.rev_idx = d.y; r.pm_idx = d.z; r.orig = d.w;
    return r;
}

// Everything phase B does for one (reaction, state).  PM selects the third-body / fall-off
// code; plain reactions compile without it.
template <bool PM, bool JAC, bool RATES>
__device__ __forceinline__ void reaction(const Tables& tb, const IO& io, const RxP& rx, int p,
                                         const double4* __restrict__ sv, const double* __restrict__ sc,
                                         double4* __restrict__ r4, double* __restrict__ rh,
                                         double* __restrict__ raw, long long s_out)
{
    const int nsp = tb.nsp, last = tb.nsp - 1;
    const int fl = rx.fl;
    const double T = sc[S_T], logT = sc[S_LOGT], iT = sc[S_IT];
    const double4 v0 = sv[rx.s0], v1 = sv[rx.s1], v2 = sv[rx.s2];
    const double4 v3 = sv[rx.s3], v4 = sv[rx.s4], v5 = sv[rx.s5];
    const double lnkf = rx.lnA + rx.b * logT - rx.Ta * iT;
    const double kf = exp(lnkf);
    const double f = kf * v0.x * v1.x * v2.x;
    const bool isrev = fl & F_REV;
    double kr = 0.0, r = 0.0;
    if (isrev) {
        const double sB = (v3.y + v4.y + v5.y) - (v0.y + v1.y + v2.y);
        kr = exp(lnkf - sB - rx.lnKc);
        r = kr * v3.x * v4.x * v5.x;
    }
    const double net = f - r;

    double PM_ = 1.0, pmt = 0.0, Xd = 0.0, e1Fi = 0.0;
    const int mi = p - tb.first_pm;
    const double* par = tb.pm_par + (PM ? mi : 0) * NPAR;
    if (PM) {
        double thd = sc[S_M];
        const int e0 = tb.pm_eff_off[mi], e1_ = tb.pm_eff_off[mi + 1];
        for (int e = e0; e < e1_; ++e) thd += tb.pm_eff_am1[e] * sv[tb.pm_eff_sp[e]].x;
        if (fl & F_PDEP) {
            const int csp = tb.pm_sp[mi];
            const double ct = csp >= 0 ? sv[csp].x : thd;
            const double e1 = exp(par[0] + par[1] * logT - par[2] * iT);
            const double Pr = ct * e1;
            const double dpr4 = par[3] + par[2] * iT - 1.0;
            const double dpr = par[1] + par[2] * iT - 1.0;
            const double i1p = 1.0 / (1.0 + Pr);
            const bool low = fl & F_LOW;
            double gg;
            if (low) { Xd = dpr4 * iT * i1p; gg = i1p; }
            else { Xd = -Pr * dpr4 * iT * i1p; gg = -Pr * i1p; }
            double F = 1.0;
            if (fl & F_TROE) {
                const double e3 = exp(T / par[7]), e1t = exp(T / par[9]);
                double Fc = par[6] * e3 + par[8] * e1t;
                double dF = par[11] * e3 - par[12] * e1t;
                if (fl & F_TROE_T2) {
                    const double e2 = exp(par[10] * iT);
                    Fc += e2;
                    dF += par[13] * iT * iT * e2;
                }
                const double lnFc = log(fmax(Fc, 1.0e-300));
                const double iln10 = 0.43429448190325182765;
                const double lF = lnFc * iln10, lP = log10_clamped(Pr);
                const double A = lP - 0.67 * lF - 0.4;
                const double Bq = 0.806 - 1.1762 * lF - 0.14 * lP;
                const double q1 = 1.0 + A * A / (Bq * Bq);
                const double lnF_AB = 2.0 * lnFc * A / (Bq * Bq * Bq * q1 * q1);
                F = exp(lnFc / q1);
                if (JAC) {
                    Xd += (1.0 / (Fc * q1) - lnF_AB * (-0.67 * iln10 * Bq + 1.1762 * iln10 * A) / Fc) * dF
                          - lnF_AB * (Bq * iln10 + 0.14 * iln10 * A) * dpr * iT;
                    gg -= lnF_AB * (Bq * iln10 + A * 0.14 * iln10);
                }
            } else if (fl & F_SRI) {
                const double lP = log10_clamped(Pr);
                const double X = 1.0 / (1.0 + lP * lP);
                F = pow(par[14] * exp(-par[15] * iT) + exp(-T / par[16]), X);
                if (fl & F_SRI5) F *= par[17] * pow(T, par[18]);
                if (JAC) {
                    const double two_iln10 = 0.86858896380650365530;
                    const double eb = exp(par[23] * iT), ec = exp(T / par[25]);
                    const double den = par[26] * eb + ec;
                    Xd += X * ((par[22] * iT * iT * eb - par[24] * ec) / den
                               - X * two_iln10 * lP * dpr * log(den) * iT);
                    if (fl & F_SRI5_DT) Xd += par[27] * iT;
                    gg -= X * X * two_iln10 * lP * log(par[19] * exp(par[20] * iT) + exp(T / par[21]));
                }
            }
            const double Fi = F * i1p;
            PM_ = low ? Fi * Pr : Fi;
            e1Fi = e1 * Fi;
            if (fl & F_PMT) pmt = gg * net;
        } else {
            PM_ = thd;
            if (fl & F_PMT) pmt = net;
        }
    }
    if (RATES && s_out >= 0) {
        if (io.fwd) put(io.fwd, io, tb.nr, s_out, rx.orig, f);
        if (io.rev && isrev) put(io.rev, io, tb.nrev, s_out, rx.rev_idx, r);
        if (PM && io.pm) put(io.pm, io, tb.npd, s_out, rx.pm_idx, PM_);
    }
    if (!JAC) {
        r4[p].x = net * PM_;
        return;
    }
    const double nre = (double)((fl >> NRE_SHIFT) & 15), npr = (double)((fl >> NPR_SHIFT) & 15);
    const double rho_inv = sc[S_RHOINV];
    const double dk = rx.b + rx.Ta * iT;
    const double sdB = (v3.z + v4.z + v5.z) - (v0.z + v1.z + v2.z);
    // irreversible: r = 0 makes this f * (dk + 1 - nre)          (cj:1461-1523)
    const double elem = net * dk + f * (1.0 - nre) - r * ((1.0 - npr) - T * sdB);
    double tT;
    if (PM) {
        if (fl & F_PDEP) tT = (PM_ * Xd * net + PM_ * iT * elem) * rho_inv;
        else tT = (-PM_ * net * iT + PM_ * iT * elem) * rho_inv;
    } else {
        tT = iT * elem * rho_inv;
    }
    if (fl & F_NO_T) tT = 0.0;
    const double extra = (PM && (fl & F_EFFN1)) ? 1.0 : 0.0;
    double inner = (nre + extra) * f - (isrev ? (npr + extra) * r : 0.0);
    if (PM && (fl & F_PMT_INJ)) inner += pmt;
    const double jy = -sc[S_MW] * rho_inv * PM_ * inner;
    if (PM && (fl & F_PMT_INJ)) pmt *= e1Fi;
    double X1 = jy, X2 = -jy;
    if (PM) { X1 += par[5] * pmt; X2 -= par[4] * pmt; }
    double* rw = raw + rx.rbase;
    const double pk = PM_ * kf;
    const double d0 = pk * v1.x * v2.x, d1 = pk * v0.x * v2.x, d2 = pk * v0.x * v1.x;
    if (rx.s0 != nsp) { if (rx.s0 == last) X2 -= d0; else *rw++ = d0; }
    if (rx.s1 != nsp) { if (rx.s1 == last) X2 -= d1; else *rw++ = d1; }
    if (rx.s2 != nsp) { if (rx.s2 == last) X2 -= d2; else *rw++ = d2; }
    if (isrev) {
        const double pr = -PM_ * kr;
        const double d3 = pr * v4.x * v5.x, d4 = pr * v3.x * v5.x, d5 = pr * v3.x * v4.x;
        if (rx.s3 != nsp) { if (rx.s3 == last) X2 -= d3; else *rw++ = d3; }
        if (rx.s4 != nsp) { if (rx.s4 == last) X2 -= d4; else *rw++ = d4; }
        if (rx.s5 != nsp) { if (rx.s5 == last) X2 -= d5; else *rw++ = d5; }
    }
    if (PM) {
        if (fl & F_EFF_SLOTS) {
            const int e0 = tb.pm_eff_off[mi], e1_ = tb.pm_eff_off[mi + 1];
            for (int e = e0; e < e1_; ++e)
                if (tb.pm_eff_sp[e] != last) *rw++ = pmt * tb.pm_eff_am1[e];
        }
        if (fl & F_WANT_PMT) *rw = pmt;
    }
    r4[p] = make_double4(net * PM_, tT, X1, X2);
    rh[p] = (v3.w + v4.w + v5.w) - (v0.w + v1.w + v2.w);
}


// Phase D for one sub-entry of LEN contributions.  TROW: energy-equation row (coefficient =
// reaction enthalpy change, scale = -1/cp_avg), else Jacobian entry (coefficient = nu as a
// bf16 pattern in the upper half of the word, scale = W_k).
template <int G, int LEN, bool TROW>
__device__ __forceinline__ void gather(const unsigned* __restrict__ cw, const double* __restrict__ raw0,
                                       int st_raw, const double* __restrict__ rh0, int st_rh,
                                       const double* __restrict__ scale, int scale_st,
                                       double* __restrict__ sval0, int st_sval, int e)
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
        const int src = w[i] & 0xFFFFu;
        if (TROW) {
            const int rxn = w[i] >> 16;
#pragma unroll
            for (int g = 0; g < G; ++g) acc[g] += rh0[g * st_rh + rxn] * raw0[g * st_raw + src];
        } else {
            const double cf = (double)__uint_as_float(w[i] & 0xFFFF0000u);
#pragma unroll
            for (int g = 0; g < G; ++g) acc[g] += cf * raw0[g * st_raw + src];
        }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) sval0[g * st_sval + e] = acc[g] * scale[g * scale_st];
}

// Phase E for one state: this warp's share of the Jacobian columns, lanes = output rows.
// NK = ceil(nsp / 32) rows per lane (0: generic loop).  SF: state-fastest output layout.
template <int NK, bool SF>
__device__ __forceinline__ void store_columns(const Tables& tb, int nsp, int lane, int w0, int wn,
                                              const double* __restrict__ rowT, const double* __restrict__ rowA,
                                              const double* __restrict__ rowB, const double* __restrict__ sv,
                                              const double* __restrict__ cp, double XT,
                                              double* __restrict__ base, long long es)
{
    const int last = nsp - 1;
    if (NK > 0) {
        double rA[NK > 0 ? NK : 1], rB[NK > 0 ? NK : 1];
#pragma unroll
        for (int i = 0; i < NK; ++i) {
            const int r = lane + 32 * i;
            rA[i] = r < nsp ? rowA[r] : 0.0;
            rB[i] = r < nsp ? rowB[r] : 0.0;
        }
        if (w0 == 0) {
#pragma unroll
            for (int i = 0; i < NK; ++i) {
                const int r = lane + 32 * i;
                if (r < nsp) base[SF ? (long long)r * es : r] = rowT[r];
            }
        }
        for (int col = w0 == 0 ? wn : w0; col < nsp; col += wn) {
            const int j = col - 1;
            const double iwj = __ldg(tb.sp_iw + j), mwfj = __ldg(tb.sp_mwf + j);
            const double ex = XT * (cp[j] - cp[last]);
            const unsigned short* jm = tb.jmap + j * nsp;
            double* out = base + (SF ? (long long)col * nsp * es : (long long)(col * nsp));
#pragma unroll
            for (int i = 0; i < NK; ++i) {
                const int r = lane + 32 * i;
                if (r < nsp) {
                    double v = iwj * (rA[i] + rB[i] * mwfj + sv[__ldg(jm + r)]);
                    if (i == 0 && r == 0) v += ex;
                    out[SF ? (long long)r * es : r] = v;
                }
            }
        }
    } else {
        if (w0 == 0)
            for (int r = lane; r < nsp; r += 32) base[SF ? (long long)r * es : r] = rowT[r];
        for (int col = w0 == 0 ? wn : w0; col < nsp; col += wn) {
            const int j = col - 1;
            const double iwj = __ldg(tb.sp_iw + j), mwfj = __ldg(tb.sp_mwf + j);
            const double ex = XT * (cp[j] - cp[last]);
            const unsigned short* jm = tb.jmap + (size_t)j * nsp;
            double* out = base + (SF ? (long long)col * nsp * es : (long long)col * nsp);
            for (int r = lane; r < nsp; r += 32) {
                double v = iwj * (rowA[r] + rowB[r] * mwfj + sv[__ldg(jm + r)]);
                if (r == 0) v += ex;
                out[SF ? (long long)r * es : r] = v;
            }
        }
    }
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
    const int nsp = tb.nsp, last = tb.nsp - 1, nsp1 = L.nsp1;
    constexpr bool JAC = (MODE & M_JAC) != 0;
    constexpr bool RATES = (MODE & M_RATES) != 0;

#define SPV(g) ((double4*)(smem + L.off_spv + (g) * L.st_spv))
#define VEC(g, v) (smem + L.off_vec + (g) * L.st_vec + (v) * nsp1)
#define CPV(b, g) (smem + L.off_cp + ((b) * G + (g)) * L.st_cp)
#define SCAL(b, g) (smem + L.off_scal + ((b) * G + (g)) * NSCAL)
#define R4V(g) ((double4*)(smem + L.off_r4 + (g) * L.st_r4))
#define RHV(g) (smem + L.off_rh + (g) * L.st_rh)
#define RAWV(g) (smem + L.off_raw + (g) * L.st_raw)
#define PARTV(g) ((double4*)(smem + L.off_part + (g) * L.st_part))
#define SVALV(g) (smem + L.off_sval + (g) * L.st_sval)

    // ---- phase A for the group starting at state s0, into buffer `buf` (one warp / state)
    auto phase_a = [&](long long s0, int buf, int g) {
        const bool live = s0 + g < io.n;
        const long long s = live ? s0 + g : (long long)io.n - 1;
        const double* ys = io.y + s * io.y_ss;
        const double T = ys[0];
        const double P = io.pres[s];
        double* Yv = smem + L.off_y + g * L.st_y;
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
                Yv[k] = ys[(long long)(k + 1) * io.y_sv] * tb.sp_w[k] / rho;
        } else {
            for (int k = lane; k < last; k += 32) {
                const double Yk = ys[(long long)(k + 1) * io.y_sv];
                Yv[k] = Yk;
                sumY += Yk;
                sumYW += Yk * tb.sp_iw[k];
            }
            sumY = warp_sum(sumY);
            sumYW = warp_sum(sumYW);
            const double yN = 1.0 - sumY;
            sumYW += yN * tb.sp_iw[last];
            mw_avg = 1.0 / sumYW;
            rho = P * mw_avg / (tb.ru * T);
            if (lane == 0) Yv[last] = yN;
        }
        __syncwarp();
        const double logT = log(T), iT = 1.0 / T;
        double cpavg = 0.0, wdcp = 0.0;
        double4* spv = SPV(g);
        double* cpv = CPV(buf, g);
        for (int k = lane; k < nsp; k += 32) {
            const double Yk = Yv[k];
            const double ck = io.in_conc ? ys[(long long)(k + 1) * io.y_sv] : rho * Yk * tb.sp_iw[k];
            if (RATES && io.conc && live) put(io.conc, io, nsp, s, k, ck);
            const double* c = tb.sp_nasa + (k * 2 + (T <= tb.sp_tmid[k] ? 0 : 1)) * 16;
            const double ruw = tb.sp_ruw[k];
            const double cp = ruw * (c[0] + T * (c[1] + T * (c[2] + T * (c[3] + c[4] * T))));
            const double hh = c[6] + T * (c[7] + T * (c[8] + c[9] * T));
            const double h = ruw * (c[5] + T * (c[0] + T * hh));
            cpv[k] = cp;
            cpavg += Yk * cp;
            double dB = 0.0;
            if (JAC) {
                const double dcp = ruw * (c[1] + T * (2.0 * c[2] + T * (3.0 * c[3] + 4.0 * c[4] * T)));
                wdcp += Yk * dcp;
                dB = (c[11] + c[5] * iT) * iT + hh;
            }
            const double Bk = c[10] + c[11] * logT + T * (c[6] + T * (c[12] + T * (c[13] + c[14] * T))) - c[5] * iT;
            spv[k] = make_double4(ck, Bk, dB, h * tb.sp_w[k]);
        }
        cpavg = warp_sum(cpavg);
        if (JAC) wdcp = warp_sum(wdcp);
        if (lane == 0) {
            spv[nsp] = make_double4(1.0, 0.0, 0.0, 0.0);       // empty reaction slot
            double* sc = SCAL(buf, g);
            sc[S_T] = T; sc[S_LOGT] = logT; sc[S_IT] = iT; sc[S_RHO] = rho;
            sc[S_RHOINV] = 1.0 / rho; sc[S_MW] = mw_avg; sc[S_M] = P / (tb.ru * T);
            sc[S_CPAVG] = cpavg; sc[S_WDCP] = wdcp; sc[S_P] = P; sc[S_NWT] = -1.0 / cpavg;
            if (RATES && io.scal3 && live) {
                double* o = io.scal3 + s * 3;
                o[0] = Yv[last]; o[1] = mw_avg; o[2] = rho;
            }
        }
    };

    if (JAC && tid < G) { SVALV(tid)[tb.zero_slot] = 0.0; RAWV(tid)[tb.nraw] = 0.0; }

    const long long ngroups = ((long long)io.n + G - 1) / G;
    int buf = 0;
    if ((long long)blockIdx.x < ngroups)
        for (int g = warp; g < G; g += nwarps) phase_a((long long)blockIdx.x * G, 0, g);
    __syncthreads();

    const bool split_roles = nwarps >= 2 * G + 2;      // enough warps to overlap A(next) with E
    for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x, buf ^= 1) {
        const long long s0 = grp * G;

        // ------------------------------------------------------------ phase B
        {
            const int n_plain = tb.first_pm;
            const int items = n_plain + tb.npm * G;
            for (int it = tid; it < items; it += blockDim.x) {
                if (it < n_plain) {
                    const RxP rx = load_rx(tb.rx_rec, it);
#pragma unroll 1
                    for (int g = 0; g < G; ++g)
                        reaction<false, JAC, RATES>(tb, io, rx, it, SPV(g), SCAL(buf, g), R4V(g), RHV(g),
                                                    RAWV(g), s0 + g < io.n ? s0 + g : -1);
                } else {
                    const int q = it - n_plain;
                    const int p = n_plain + q / G, g = q % G;
                    const RxP rx = load_rx(tb.rx_rec, p);
                    reaction<true, JAC, RATES>(tb, io, rx, p, SPV(g), SCAL(buf, g), R4V(g), RHV(g),
                                               RAWV(g), s0 + g < io.n ? s0 + g : -1);
                }
            }
        }
        __syncthreads();

        // ------------------------------------------------------------ phase C1
        for (int c = tid; c < tb.nchunk; c += blockDim.x) {
            const int4* rxp = (const int4*)(tb.chk_rx + c * RCH);
            const double2* nup = (const double2*)(tb.chk_nu + c * RCH);
            const int4 pa = __ldg(rxp), pb = __ldg(rxp + 1);
            const double2 na = __ldg(nup), nb = __ldg(nup + 1), nc = __ldg(nup + 2), nd = __ldg(nup + 3);
            const int pi[RCH] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
            const double nu[RCH] = {na.x, na.y, nb.x, nb.y, nc.x, nc.y, nd.x, nd.y};
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const double4* R = R4V(g);
                double4 a = make_double4(0.0, 0.0, 0.0, 0.0);
#pragma unroll
                for (int e = 0; e < RCH; ++e) {
                    if (JAC) {
                        const double4 v = R[pi[e]];
                        a.x += nu[e] * v.x; a.y += nu[e] * v.y; a.z += nu[e] * v.z; a.w += nu[e] * v.w;
                    } else {
                        a.x += nu[e] * R[pi[e]].x;
                    }
                }
                PARTV(g)[c] = a;
            }
        }

        // ------------------------------------------------------------ phase D
        if (JAC) {
            const double* raw0 = RAWV(0);
            const double* rh0 = RHV(0);
            double* sval0 = SVALV(0);
            const double* nwt0 = SCAL(buf, 0) + S_NWT;
            for (int e = tid; e < tb.nsub; e += blockDim.x) {
                int c = 0;
#pragma unroll
                for (int i = 1; i < 8; ++i) c += e >= tb.cls_sub[i];
                const int rel = e - tb.cls_sub[c];
                const unsigned* cw = tb.con + tb.cls_con[c];
                switch (c) {
                case 0: gather<G, 8, false>(cw + rel * 8, raw0, L.st_raw, rh0, L.st_rh, tb.sub_w + e, 0, sval0, L.st_sval, e); break;
                case 1: gather<G, 4, false>(cw + rel * 4, raw0, L.st_raw, rh0, L.st_rh, tb.sub_w + e, 0, sval0, L.st_sval, e); break;
                case 2: gather<G, 2, false>(cw + rel * 2, raw0, L.st_raw, rh0, L.st_rh, tb.sub_w + e, 0, sval0, L.st_sval, e); break;
                case 3: gather<G, 1, false>(cw + rel, raw0, L.st_raw, rh0, L.st_rh, tb.sub_w + e, 0, sval0, L.st_sval, e); break;
                case 4: gather<G, 8, true>(cw + rel * 8, raw0, L.st_raw, rh0, L.st_rh, nwt0, NSCAL, sval0, L.st_sval, e); break;
                case 5: gather<G, 4, true>(cw + rel * 4, raw0, L.st_raw, rh0, L.st_rh, nwt0, NSCAL, sval0, L.st_sval, e); break;
                case 6: gather<G, 2, true>(cw + rel * 2, raw0, L.st_raw, rh0, L.st_rh, nwt0, NSCAL, sval0, L.st_sval, e); break;
                default: gather<G, 1, true>(cw + rel, raw0, L.st_raw, rh0, L.st_rh, nwt0, NSCAL, sval0, L.st_sval, e); break;
                }
            }
        }
        __syncthreads();

        // ------------------------------------------------------------ phase C2 (+ combine)
        if (warp < G) {
            const int g = warp;
            double* sc = SCAL(buf, g);
            const double mw_rho = sc[S_MW] * sc[S_RHOINV];
            const double4* part = PARTV(g);
            const double4* spv = SPV(g);
            const double* cpv = CPV(buf, g);
            double H1 = 0.0, HA = 0.0, HB = 0.0, HT = 0.0, SCP = 0.0;
            for (int k = lane; k < nsp; k += 32) {
                const int c0 = tb.sp_chk_off[k], c1 = tb.sp_chk_off[k + 1];
                double4 s = make_double4(0.0, 0.0, 0.0, 0.0);
                for (int c = c0; c < c1; ++c) {
                    const double4 v = part[c];
                    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                }
                const double wk = tb.sp_w[k], hW = spv[k].w;
                VEC(g, V_WDOT)[k] = s.x;
                H1 += hW * s.x;
                if (JAC) {
                    const double comp = s.x * mw_rho;
                    const double a = s.z + comp, b = s.w - comp;
                    if (k < last) {
                        VEC(g, V_ROWT)[k + 1] = wk * s.y;
                        VEC(g, V_ROWA)[k + 1] = wk * a;
                        VEC(g, V_ROWB)[k + 1] = wk * b;
                    }
                    HT += hW * s.y; HA += hW * a; HB += hW * b; SCP += cpv[k] * wk * s.x;
                }
            }
            H1 = warp_sum(H1);
            if (JAC) { HA = warp_sum(HA); HB = warp_sum(HB); HT = warp_sum(HT); SCP = warp_sum(SCP); }
            if (lane == 0) {
                sc[S_H1] = H1;
                if (JAC) {
                    // energy-equation row (cj:3095-3254) and jac[0] (cj:1853-1905)
                    const double rho = sc[S_RHO], cpavg = sc[S_CPAVG];
                    const double nwt = sc[S_NWT];
                    VEC(g, V_ROWA)[0] = nwt * HA;
                    VEC(g, V_ROWB)[0] = nwt * HB;
                    VEC(g, V_ROWT)[0] = -(-sc[S_WDCP] / cpavg * H1 + SCP + HT * rho) / (rho * cpavg);
                    sc[S_XT] = H1 / (rho * cpavg * cpavg);
                }
            }
        } else if (JAC) {
            const int t0 = tid - 32 * G, tn = blockDim.x - 32 * G;
            for (int t = t0; t < tb.nsplit; t += tn) {
                const int c0 = tb.cmb_off[t], c1 = tb.cmb_off[t + 1];
                double acc[G];
#pragma unroll
                for (int g = 0; g < G; ++g) acc[g] = 0.0;
                for (int c = c0; c < c1; ++c) {
                    const int ix = tb.cmb_idx[c];
#pragma unroll
                    for (int g = 0; g < G; ++g) acc[g] += SVALV(g)[ix];
                }
#pragma unroll
                for (int g = 0; g < G; ++g) SVALV(g)[tb.nsub + t] = acc[g];
            }
        }
        __syncthreads();

        // ------------------------------------------------------------ rates / dydt outputs
        if (RATES || (MODE & M_DYDT)) {
            for (int g = warp; g < G; g += nwarps) {
                const long long s = s0 + g;
                if (s >= io.n) continue;
                const double* sc = SCAL(buf, g);
                const double* wd = VEC(g, V_WDOT);
                for (int k = lane; k < nsp; k += 32) {
                    if (RATES && io.sr) put(io.sr, io, nsp, s, k, wd[k]);
                    if (io.dy && k < last) {
                        const double v = wd[k] * tb.sp_w[k] * sc[S_RHOINV];
                        if (RATES) put(io.dy, io, nsp, s, k + 1, v);
                        else io.dy[s * io.dy_ss + (long long)(k + 1) * io.dy_sv] = v;
                    }
                }
                if (lane == 0 && io.dy) {
                    const double v = -1.0 / (sc[S_RHO] * sc[S_CPAVG]) * sc[S_H1];
                    if (RATES) put(io.dy, io, nsp, s, 0, v);
                    else io.dy[s * io.dy_ss] = v;
                }
            }
        }

        // ------------------------------------------------------------ phase E  (|| A of next group)
        const long long next = grp + gridDim.x;
        const bool a_here = split_roles && warp < G;
        if (a_here) {
            if (next < ngroups) phase_a(next * G, buf ^ 1, warp);
        } else if (JAC) {
            const int w0 = split_roles ? warp - G : warp, wn = split_roles ? nwarps - G : nwarps;
            const bool sf = io.jac_layout != 0;
            for (int g = 0; g < G; ++g) {
                const long long s = s0 + g;
                if (s >= io.n) break;
                const double* sc = SCAL(buf, g);
                const double* cp = CPV(buf, g);
                const double *rT = VEC(g, V_ROWT), *rA = VEC(g, V_ROWA), *rB = VEC(g, V_ROWB);
                const double* sv = SVALV(g);
                const double XT = sc[S_XT];
                if (sf) {
                    store_columns<0, true>(tb, nsp, lane, w0, wn, rT, rA, rB, sv, cp, XT, io.jac + s, io.jac_ld);
                } else {
                    double* base = io.jac + s * (long long)(nsp * nsp);
                    if (nsp <= 32) store_columns<1, false>(tb, nsp, lane, w0, wn, rT, rA, rB, sv, cp, XT, base, 1);
                    else if (nsp <= 64) store_columns<2, false>(tb, nsp, lane, w0, wn, rT, rA, rB, sv, cp, XT, base, 1);
                    else if (nsp <= 128) store_columns<4, false>(tb, nsp, lane, w0, wn, rT, rA, rB, sv, cp, XT, base, 1);
                    else store_columns<0, false>(tb, nsp, lane, w0, wn, rT, rA, rB, sv, cp, XT, base, 1);
                }
            }
        }
        if (!split_roles) {
            __syncthreads();
            if (next < ngroups)
                for (int g = warp; g < G; g += nwarps) phase_a(next * G, buf ^ 1, g);
        }
        __syncthreads();
    }
#undef SPV
#undef VEC
#undef CPV
#undef SCAL
#undef R4V
#undef RHV
#undef RAWV
#undef PARTV
#undef SVALV
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
