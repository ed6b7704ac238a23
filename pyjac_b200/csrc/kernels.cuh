// Data-driven sm_100a kernels: species rates, dydt and the analytical Jacobian of a
// gas-phase kinetic mechanism, evaluated for a batch of states from mechanism *tables*
// (pyjac_b200/tables.py) instead of per-mechanism generated code.
//
// Replaces the emitted library of the reference: eval_conc (rate_subs.py:1626-1706),
// eval_rxn_rates (:254-876), get_rxn_pres_mod (:879-1294), eval_spec_rates (:1297-1542),
// eval_h / eval_cp (:1806-2086), dydt (:2171-2335) and eval_jacob
// (create_jacobian.py:2189-3298).  The arithmetic is regrouped (see tables.py) so that
// each Jacobian element is produced and stored exactly once:
//
//   phase A  one warp per state: mass fractions -> concentrations, NASA-7 thermo
//            (cp, h, Gibbs term B_k for Kc, dB_k/dT, dcp_k/dT) into shared memory
//   phase B  one thread per reaction (x G states): kf, kr = exp(ln kf - sum nu B - ...),
//            rates of progress, third-body / fall-off factors and their derivatives; per
//            reaction it leaves 4 scalars (R4) and its non-zero d(rate)/dC values (raw)
//   phase C  one warp per species: warp-shuffle reductions over the species' reactions
//            -> wdot_k, T-column, dense Jacobian vectors A_k, B_k
//   phase D  one thread per structurally non-zero (k, j): gather of raw values
//   phase E  one warp per Jacobian column: dense + sparse assembly, energy-equation row by
//            warp reduction, coalesced stores straight to HBM
//
// A thread block is persistent and walks over groups of G states.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pj {

enum : int {
    F_REV = 1, F_THD = 2, F_PDEP = 4, F_LOW = 8, F_TROE = 16, F_SRI = 32, F_PMT = 64,
    F_PMT_INJ = 128, F_TROE_T2 = 256, F_SRI5 = 512, F_SRI5_DT = 1024, F_NO_T = 2048,
    F_EFFN1 = 4096, F_WANT_PMT = 1 << 16, NRE_SHIFT = 20, NPR_SHIFT = 24, NPAR = 32
};

enum : int { M_JAC = 1, M_DYDT = 2, M_RATES = 4 };

struct Tables {
    int nsp, nr, nrev, npd, nraw, nnz, ncon, ncoef, first_pm, npm, nred, maxred;
    double ru, ln_pa_ru;
    const double *sp_w, *sp_iw, *sp_ruw, *sp_tmid, *sp_mwf, *sp_nasa;
    const int* sp_seen;
    const int *rx_orig, *rx_flags, *rx_rev_idx, *rx_pm_idx, *rx_raw_base, *rx_slots;
    const double* rx_arr;
    const double* pm_par;
    const int *pm_sp, *pm_eff_off, *pm_eff_sp;
    const double* pm_eff_am1;
    const int *red_off, *red_rx;
    const double* red_nu;
    const int *ent_kj, *ent_off, *con;
    const double* coef;
    const unsigned short* jmap;
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

// shared-memory carve-up (in doubles), filled on the host
struct Layout {
    int nsp1;        // padded species vector length (>= nsp + 1)
    int off_vec;     // G * NVEC * nsp1
    int off_scal;    // G * NSCAL
    int off_r4;      // G * 4 * nr
    int off_raw;     // G * (nraw + 1)
    int off_sval;    // G * (nnz + 1)
    int total;
};

enum : int { V_CONC = 0, V_B, V_DB, V_H, V_CP, V_Y, V_WDOT, V_TCOL, V_A, V_BK, NVEC };
enum : int { S_T = 0, S_LOGT, S_IT, S_RHO, S_RHOINV, S_MW, S_M, S_CPAVG, S_WDCP, S_P, NSCAL = 16 };

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

template <int G, int MODE>
__global__ void __launch_bounds__(512, 1)
k_eval(const __grid_constant__ Tables tb, const __grid_constant__ IO io,
       const __grid_constant__ Layout L)
{
    extern __shared__ double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int nsp = tb.nsp, last = tb.nsp - 1, nr = tb.nr, nsp1 = L.nsp1;
    constexpr bool JAC = (MODE & M_JAC) != 0;
    constexpr bool RATES = (MODE & M_RATES) != 0;

    double* const vec = smem + L.off_vec;
    double* const scal = smem + L.off_scal;
    double* const r4 = smem + L.off_r4;
    double* const raw = smem + L.off_raw;
    double* const sval = smem + L.off_sval;
#define VEC(g, v) (vec + ((g) * NVEC + (v)) * nsp1)

    const long long ngroups = ((long long)io.n + G - 1) / G;
    for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const long long s0 = grp * G;

        // ------------------------------------------------------------ phase A
        for (int g = warp; g < G; g += nwarps) {
            const long long s = (s0 + g < io.n) ? s0 + g : (long long)io.n - 1;
            const double* ys = io.y + s * io.y_ss;
            const double T = ys[0];
            const double P = io.pres[s];
            double* Yv = VEC(g, V_Y);
            double sumY = 0.0, sumYW = 0.0;
            double mw_avg, rho;
            if (io.in_conc) {
                // concentrations supplied (eval_rxn_rates / get_rxn_pres_mod entry points):
                // Y_k = C_k W_k / rho with rho = sum C_k W_k
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
            for (int k = lane; k < nsp; k += 32) {
                const double Yk = Yv[k];
                const double ck = io.in_conc ? ys[(long long)(k + 1) * io.y_sv] : rho * Yk * tb.sp_iw[k];
                VEC(g, V_CONC)[k] = ck;
                if (RATES && io.conc && s0 + g < io.n) put(io.conc, io, nsp, s0 + g, k, ck);
                const double* c = tb.sp_nasa + (k * 2 + (T <= tb.sp_tmid[k] ? 0 : 1)) * 16;
                const double ruw = tb.sp_ruw[k];
                const double cp = ruw * (c[0] + T * (c[1] + T * (c[2] + T * (c[3] + c[4] * T))));
                const double hh = c[6] + T * (c[7] + T * (c[8] + c[9] * T));
                const double h = ruw * (c[5] + T * (c[0] + T * hh));
                VEC(g, V_CP)[k] = cp;
                VEC(g, V_H)[k] = h;
                cpavg += Yk * cp;
                if (JAC) {
                    const double dcp = ruw * (c[1] + T * (2.0 * c[2] + T * (3.0 * c[3] + 4.0 * c[4] * T)));
                    wdcp += Yk * dcp;
                    VEC(g, V_DB)[k] = (c[11] + c[5] * iT) * iT + hh;
                }
                VEC(g, V_B)[k] = c[10] + c[11] * logT + T * (c[6] + T * (c[12] + T * (c[13] + c[14] * T))) - c[5] * iT;
            }
            cpavg = warp_sum(cpavg);
            if (JAC) wdcp = warp_sum(wdcp);
            if (lane == 0) {
                VEC(g, V_CONC)[nsp] = 1.0;
                VEC(g, V_B)[nsp] = 0.0;
                VEC(g, V_DB)[nsp] = 0.0;
                double* sc = scal + g * NSCAL;
                sc[S_T] = T; sc[S_LOGT] = logT; sc[S_IT] = iT; sc[S_RHO] = rho;
                sc[S_RHOINV] = 1.0 / rho; sc[S_MW] = mw_avg; sc[S_M] = P / (tb.ru * T);
                sc[S_CPAVG] = cpavg; sc[S_WDCP] = wdcp; sc[S_P] = P;
                if (RATES && io.scal3 && s0 + g < io.n) {
                    double* o = io.scal3 + (s0 + g) * 3;
                    o[0] = Yv[last]; o[1] = mw_avg; o[2] = rho;
                }
            }
        }
        __syncthreads();

        // ------------------------------------------------------------ phase B
        for (int p = tid; p < nr; p += blockDim.x) {
            const int fl = tb.rx_flags[p];
            const int* sl = tb.rx_slots + p * 6;
            const int s_0 = sl[0], s_1 = sl[1], s_2 = sl[2], s_3 = sl[3], s_4 = sl[4], s_5 = sl[5];
            const double lnA = tb.rx_arr[p * 4], bexp = tb.rx_arr[p * 4 + 1], Ta = tb.rx_arr[p * 4 + 2],
                         lnKc = tb.rx_arr[p * 4 + 3];
            const bool isrev = fl & F_REV;
            const double nre = (double)((fl >> NRE_SHIFT) & 15), npr = (double)((fl >> NPR_SHIFT) & 15);
            const int mi = p - tb.first_pm;
            const double* par = tb.pm_par + (mi > 0 ? mi : 0) * NPAR;
            const int rbase = JAC ? tb.rx_raw_base[p] : 0;
#pragma unroll 1
            for (int g = 0; g < G; ++g) {
                const double* sc = scal + g * NSCAL;
                const double* conc = VEC(g, V_CONC);
                const double* Bv = VEC(g, V_B);
                const double T = sc[S_T], logT = sc[S_LOGT], iT = sc[S_IT];
                const double c0 = conc[s_0], c1 = conc[s_1], c2 = conc[s_2];
                const double c3 = conc[s_3], c4 = conc[s_4], c5 = conc[s_5];
                const double lnkf = lnA + bexp * logT - Ta * iT;
                const double kf = exp(lnkf);
                const double f = kf * c0 * c1 * c2;
                double kr = 0.0, r = 0.0;
                if (isrev) {
                    const double sB = (Bv[s_3] + Bv[s_4] + Bv[s_5]) - (Bv[s_0] + Bv[s_1] + Bv[s_2]);
                    kr = exp(lnkf - sB - lnKc);
                    r = kr * c3 * c4 * c5;
                }
                const double net = f - r;

                double PM = 1.0, pmt = 0.0, Xd = 0.0, e1Fi = 0.0;
                if (fl & (F_THD | F_PDEP)) {
                    double thd = sc[S_M];
                    for (int e = tb.pm_eff_off[mi]; e < tb.pm_eff_off[mi + 1]; ++e)
                        thd += tb.pm_eff_am1[e] * conc[tb.pm_eff_sp[e]];
                    if (fl & F_PDEP) {
                        const int csp = tb.pm_sp[mi];
                        const double ct = csp >= 0 ? conc[csp] : thd;
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
                        PM = low ? Fi * Pr : Fi;
                        e1Fi = e1 * Fi;
                        if (fl & F_PMT) pmt = gg * net;
                    } else {
                        PM = thd;
                        if (fl & F_PMT) pmt = net;
                    }
                }
                if (RATES && s0 + g < io.n) {
                    const long long s = s0 + g;
                    if (io.fwd) put(io.fwd, io, nr, s, tb.rx_orig[p], f);
                    if (io.rev && isrev) put(io.rev, io, tb.nrev, s, tb.rx_rev_idx[p], r);
                    if (io.pm && (fl & (F_THD | F_PDEP))) put(io.pm, io, tb.npd, s, tb.rx_pm_idx[p], PM);
                }
                double* R = r4 + ((size_t)g * nr + p) * 4;
                R[0] = net * PM;
                if (JAC) {
                    const double rho_inv = sc[S_RHOINV];
                    double tT = 0.0;
                    if (!(fl & F_NO_T)) {
                        const double dk = bexp + Ta * iT;
                        double elem;
                        if (isrev) {
                            const double* dB = VEC(g, V_DB);
                            const double sdB = (dB[s_3] + dB[s_4] + dB[s_5]) - (dB[s_0] + dB[s_1] + dB[s_2]);
                            elem = net * dk + f * (1.0 - nre) - r * ((1.0 - npr) - T * sdB);
                        } else {
                            elem = f * (dk + (1.0 - nre));
                        }
                        if (fl & F_PDEP) tT = (PM * Xd * net + PM * iT * elem) * rho_inv;
                        else if (fl & F_THD) tT = (-PM * net * iT + PM * iT * elem) * rho_inv;
                        else tT = iT * elem * rho_inv;
                    }
                    const double extra = (fl & F_EFFN1) ? 1.0 : 0.0;
                    double inner = (nre + extra) * f - (isrev ? (npr + extra) * r : 0.0);
                    if (fl & F_PMT_INJ) inner += pmt;
                    const double jy = -sc[S_MW] * rho_inv * PM * inner;
                    if (fl & F_PMT_INJ) pmt *= e1Fi;
                    double X1 = jy, X2 = -jy;
                    if (fl & (F_THD | F_PDEP)) { X1 += par[5] * pmt; X2 -= par[4] * pmt; }
                    double* rw = raw + (size_t)g * (tb.nraw + 1) + rbase;
                    const double pk = PM * kf;
                    const double d0 = pk * c1 * c2, d1 = pk * c0 * c2, d2 = pk * c0 * c1;
                    if (s_0 != nsp) { if (s_0 == last) X2 -= d0; else *rw++ = d0; }
                    if (s_1 != nsp) { if (s_1 == last) X2 -= d1; else *rw++ = d1; }
                    if (s_2 != nsp) { if (s_2 == last) X2 -= d2; else *rw++ = d2; }
                    if (isrev) {
                        const double pr = -PM * kr;
                        const double d3 = pr * c4 * c5, d4 = pr * c3 * c5, d5 = pr * c3 * c4;
                        if (s_3 != nsp) { if (s_3 == last) X2 -= d3; else *rw++ = d3; }
                        if (s_4 != nsp) { if (s_4 == last) X2 -= d4; else *rw++ = d4; }
                        if (s_5 != nsp) { if (s_5 == last) X2 -= d5; else *rw++ = d5; }
                    }
                    if (fl & F_WANT_PMT) *rw = pmt;
                    R[1] = tT; R[2] = X1; R[3] = X2;
                }
            }
        }
        __syncthreads();

        // ------------------------------------------------------------ phase C
        for (int k = warp; k < nsp; k += nwarps) {
            const int o0 = tb.red_off[k], o1 = tb.red_off[k + 1];
            double acc[G][4];
#pragma unroll
            for (int g = 0; g < G; ++g) acc[g][0] = acc[g][1] = acc[g][2] = acc[g][3] = 0.0;
            for (int e = o0 + lane; e < o1; e += 32) {
                const int p = tb.red_rx[e];
                const double nu = tb.red_nu[e];
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const double* R = r4 + ((size_t)g * nr + p) * 4;
                    acc[g][0] += nu * R[0];
                    if (JAC) { acc[g][1] += nu * R[1]; acc[g][2] += nu * R[2]; acc[g][3] += nu * R[3]; }
                }
            }
#pragma unroll
            for (int g = 0; g < G; ++g) {
                acc[g][0] = warp_sum(acc[g][0]);
                if (JAC) { acc[g][1] = warp_sum(acc[g][1]); acc[g][2] = warp_sum(acc[g][2]); acc[g][3] = warp_sum(acc[g][3]); }
            }
            if (lane == 0) {
                const double wk = tb.sp_w[k];
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const double* sc = scal + g * NSCAL;
                    VEC(g, V_WDOT)[k] = acc[g][0];
                    if (JAC) {
                        const double comp = acc[g][0] * sc[S_MW] * sc[S_RHOINV];
                        VEC(g, V_TCOL)[k] = wk * acc[g][1];
                        VEC(g, V_A)[k] = acc[g][2] + comp;
                        VEC(g, V_BK)[k] = acc[g][3] - comp;
                    }
                }
            }
        }

        // ------------------------------------------------------------ phase D
        if (JAC) {
            for (int e = tid; e < tb.nnz; e += blockDim.x) {
                const int c0 = tb.ent_off[e], c1 = tb.ent_off[e + 1];
                double acc[G];
#pragma unroll
                for (int g = 0; g < G; ++g) acc[g] = 0.0;
                for (int c = c0; c < c1; ++c) {
                    const unsigned cc = (unsigned)tb.con[c];
                    const double cf = tb.coef[cc >> 16];
                    const int src = cc & 0xFFFFu;
#pragma unroll
                    for (int g = 0; g < G; ++g) acc[g] += cf * raw[(size_t)g * (tb.nraw + 1) + src];
                }
#pragma unroll
                for (int g = 0; g < G; ++g) sval[(size_t)g * (tb.nnz + 1) + e] = acc[g];
            }
            if (tid < G) sval[(size_t)tid * (tb.nnz + 1) + tb.nnz] = 0.0;
        }
        __syncthreads();

        // ------------------------------------------------------------ outputs
        if (RATES || (MODE & M_DYDT)) {
            for (int g = warp; g < G; g += nwarps) {
                const long long s = s0 + g;
                if (s >= io.n) continue;
                const double* sc = scal + g * NSCAL;
                const double* wd = VEC(g, V_WDOT);
                const double* h = VEC(g, V_H);
                double H1 = 0.0;
                for (int k = lane; k < nsp; k += 32) {
                    const double wk = tb.sp_w[k];
                    H1 += wd[k] * h[k] * wk;
                    if (RATES && io.sr) put(io.sr, io, nsp, s, k, wd[k]);
                    if (io.dy && k < last) {
                        const double v = wd[k] * wk * sc[S_RHOINV];
                        if (RATES) put(io.dy, io, nsp, s, k + 1, v);
                        else io.dy[s * io.dy_ss + (long long)(k + 1) * io.dy_sv] = v;
                    }
                }
                H1 = warp_sum(H1);
                if (lane == 0 && io.dy) {
                    const double v = -1.0 / (sc[S_RHO] * sc[S_CPAVG]) * H1;
                    if (RATES) put(io.dy, io, nsp, s, 0, v);
                    else io.dy[s * io.dy_ss] = v;
                }
            }
        }

        // ------------------------------------------------------------ phase E
        if (JAC) {
            for (int g = 0; g < G; ++g) {
                const long long s = s0 + g;
                if (s >= io.n) break;
                const double* sc = scal + g * NSCAL;
                const double* wd = VEC(g, V_WDOT);
                const double* h = VEC(g, V_H);
                const double* cp = VEC(g, V_CP);
                const double* Ak = VEC(g, V_A);
                const double* Bk = VEC(g, V_BK);
                const double* tc = VEC(g, V_TCOL);
                const double* sv = sval + (size_t)g * (tb.nnz + 1);
                const double rho = sc[S_RHO], cpavg = sc[S_CPAVG];
                double H1 = 0.0;
                for (int k = lane; k < nsp; k += 32) H1 += wd[k] * h[k] * tb.sp_w[k];
                H1 = warp_sum(H1);
                const double wt = 1.0 / cpavg, jt = 1.0 / (rho * cpavg * cpavg);
                const bool sf = io.jac_layout != 0;
                double* const base = sf ? io.jac + s : io.jac + s * (long long)nsp * nsp;
                const long long es = sf ? io.jac_ld : 1;       // element stride
                for (int col = warp; col < nsp; col += nwarps) {
                    double* out = base + (long long)col * nsp * es;
                    double part = 0.0;
                    if (col == 0) {
                        const double wdcp_cp = -sc[S_WDCP] / cpavg;
                        for (int k = lane; k < nsp; k += 32) {
                            const double t = tc[k];
                            part += wd[k] * tb.sp_w[k] * (wdcp_cp * h[k] + cp[k]) + t * h[k] * rho;
                            if (k < last) out[(long long)(k + 1) * es] = t;
                        }
                        part = warp_sum(part);
                        if (lane == 0) out[0] = -part / (rho * cpavg);
                    } else {
                        const int j = col - 1;
                        const double iwj = tb.sp_iw[j], mwfj = tb.sp_mwf[j];
                        const unsigned short* jm = tb.jmap + (size_t)j * nsp;
                        for (int k = lane; k < nsp; k += 32) {
                            const double v = tb.sp_w[k] * iwj * (Ak[k] + Bk[k] * mwfj + sv[jm[k]]);
                            part += h[k] * v;
                            if (k < last) out[(long long)(k + 1) * es] = v;
                        }
                        part = warp_sum(part);
                        if (lane == 0) out[0] = -wt * part + jt * (cp[j] - cp[last]) * H1;
                    }
                }
            }
        }
        __syncthreads();
    }
#undef VEC
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
        double rate = fwd[tb.rx_orig[p]];
        if (tb.rx_rev_idx[p] >= 0) rate -= rev[tb.rx_rev_idx[p]];
        if (tb.rx_pm_idx[p] >= 0) rate *= pm[tb.rx_pm_idx[p]];
        acc += tb.red_nu[e] * rate;
    }
    sp_rates[k] = acc;
}

}  // namespace pj
