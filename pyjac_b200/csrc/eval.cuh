// eval_jacob for sm_100a: the analytical Jacobian of a batch of states, driven by mechanism
// tables (pyjac_b200/tables.py) and a static work schedule (pyjac_b200/plan.py).
//
// Replaces the reference's generated, fully unrolled one-thread-per-state eval_jacob
// (pyjac/core/create_jacobian.py:2189-3298) together with the rate routines it calls
// (rate_subs.py:254-876, 879-1294, 1297-1542, 1626-1706, 1806-2086).
//
// A persistent thread block evaluates GS states at a time.  The 32 lanes of a warp are NSUB
// = 64 / GS sub-groups of GS / 2 lanes; a lane carries two neighbouring states.  Shared
// memory holds *rows* of GS doubles (one value per state), so one 16-byte access serves both
// states of a lane, and the NSUB sub-groups of a warp work on NSUB different table items at
// the same time; table indices are decoded once per item for all GS states.  A species owns
// eight consecutive rows (C, B, dB/dT, hW, WA, WB, WT, cp), a reaction five (net, tT, X1, X2,
// dH), so that one address computation serves all values of an item.  All shared-memory
// traffic uses 32-bit shared-space addresses (ld/st.shared.v2.f64).
//
// Phases per group of GS states, separated by block barriers:
//
//   A0  warp 0: mass fractions -> Y_N, mean molecular weight, density, per-state scalars
//   A1  per species: concentration, NASA-7 cp / h / Gibbs term B_k / dB_k/dT
//   B   per reaction: kf, kr, rates of progress, third-body / fall-off factors and their
//       derivatives -> (net rate, T-column term, X1, X2, reaction enthalpy) and the non-zero
//       d(rate)/dC values ("raw" rows)
//   C   per species: sum_i nu_ki (net, tT, X1, X2) -> the dense rank-2 part of the Jacobian
//       (WA = W_k a_k, WB = W_k b_k, WT = W_k * T-column); per-warp partial dot products for
//       the energy equation
//   DE  per Jacobian element, in steps of NSUB elements of equal (padded) list length:
//       dense part + sparse gather over raw rows in registers -> one store per element.
//       Warp 0 first reduces the partial dot products; the energy-equation row (enthalpy-
//       weighted gathers) runs after the species rows, behind a named barrier that only the
//       warps owning such elements wait on.
#pragma once
#include "common.cuh"

namespace pj5 {

using namespace pj;

// development builds (-DPJ_DEV, tools/skip_mech.sh) can leave phases out when timing; a release build has no such switch
#ifdef PJ_DEV
#define PJ_SKIP(MASK) ((io.dbg_skip & (MASK)) != 0)
#else
#define PJ_SKIP(MASK) false
#endif

// the two modes that run the whole Jacobian pipeline (phases B / C in full, then DE)
__host__ __device__ constexpr bool jac_like(int mode) { return mode == M_JAC || mode == M_FACT; }

struct Plan {
    int gs, nt, nw, nsub, oSP, oRX, oRAW, oSC, oPA, total, t_sync, coop, tcoop, oCF;
    int wsg;                         // 1: the working set lives in global memory (IO::ws), not in shared memory
    const int4* rx;
    const int4* rx_out;              // {fwd index, rev index or -1, pres_mod index or -1, 0} (M_RATES)
    const int* eff_off;
    const int4* eff;
    const int *b_off, *b_npm, *b_item;
    const int* c_off;
    const int4* c_item;
    const uint2* c_str;
    const int* d_off;
    const int2* d_item;
    const uint2* d_str;
    const int* s_off;
    const uint4* s_str;
    const int* o_off;
    const uint2* o_str;
    const int* t_off;
    const int2* t_item;
    const uint2* t_str;
    const double2* colfac;
    const int* fac_map;              // M_FACT: dense element -> slot of the factored record (sparse block), -1: none
    int fac_nnz;                     // entries of the sparse block
};

// per-state scalars: phase A0 writes Q_* (two buffers of 8 rows), phase DE derives S_*
enum : int { Q_T = 0, Q_LOGT, Q_IT, Q_RHO, Q_RHOINV, Q_LNP, Q_MWR, Q_M };     // Q_LNP: only with PLOG reactions
enum : int { S_NWT = 0, S_A0, S_B0, S_XT, S_CPL, S_P = 5, NQ = 24 };   // S_P + buffer: pressure (PLOG)
enum : int { D_H1 = 0, D_HA, D_HB, D_HT, D_SCP, D_CPAVG, D_WDCP, NPART = 7 };
// A species owns SP_SLOTS rows.  Even slots (C, dB, WA, WT) are read at E + {0, 2, 4, 6} rows, odd
// slots (B, hW, WB, cp) at O + {0, 2, 4, 6} rows, where E = region + k * SPB + (k & 1) * RB is the
// species' even-slot base and O = E ^ RB: the slot pair of odd species is swapped, so that the
// rows of different species, which all sit at the same offset of a 128-byte bank line otherwise,
// spread over both halves of the banks.
enum : int { SP_SLOTS = 8, E_C = 0, E_DB = 2, E_WA = 4, E_WT = 6, O_B = 0, O_HW = 2, O_WB = 4, O_CP = 6, E_Y = E_C };
enum : int { RX_NET = 0, RX_TT, RX_X1, RX_X2, RX_DH, RX_SLOTS };
enum : unsigned { NULL_E = 0x3FFFFFu };

struct V {
    double x, y;
};
// Where the working set of a block lives.  G = false: shared memory, `a` is a 32-bit shared-space
// address.  G = true: a per-block scratch area in global memory (mechanisms whose working set
// exceeds shared memory; L1 / L2 hold it), `a` is a byte offset from the block's base `g`.
template <bool G>
struct Mem {
    const char* g;
    template <int OFF>
    __device__ __forceinline__ V ld(unsigned a) const
    {
        V v;
        if (G) asm volatile("ld.global.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "l"(g + a), "n"(OFF) : "memory");
        else asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(a), "n"(OFF));
        return v;
    }
    template <int OFF>
    __device__ __forceinline__ void st(unsigned a, V v) const
    {
        if (G) asm volatile("st.global.v2.f64 [%0+%1], {%2, %3};" ::"l"(g + a), "n"(OFF), "d"(v.x), "d"(v.y) : "memory");
        else asm volatile("st.shared.v2.f64 [%0+%1], {%2, %3};" ::"r"(a), "n"(OFF), "d"(v.x), "d"(v.y) : "memory");
    }
    // store only where p holds (no branch)
    template <int OFF>
    __device__ __forceinline__ void st_if(bool p, unsigned a, V v) const
    {
        if (G) asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %4, 0;\n\t@q st.global.v2.f64 [%0+%1], {%2, %3};\n\t}"
                            ::"l"(g + a), "n"(OFF), "d"(v.x), "d"(v.y), "r"((int)p) : "memory");
        else asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %4, 0;\n\t@q st.shared.v2.f64 [%0+%1], {%2, %3};\n\t}"
                          ::"r"(a), "n"(OFF), "d"(v.x), "d"(v.y), "r"((int)p) : "memory");
    }
};
// every routine that touches the working set has a `mem` in scope
#define LDS(OFF, ...) mem.template ld<(OFF)>(__VA_ARGS__)
#define STS(OFF, ...) mem.template st<(OFF)>(__VA_ARGS__)
#define STS_IF(OFF, ...) mem.template st_if<(OFF)>(__VA_ARGS__)
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// exp of N values at once (so that the polynomial constants are materialised once): arguments
// clamped to [-708, 708], Cody-Waite reduction to |r| <= ln2/2, degree-12 Taylor polynomial
// (truncation 1.7e-16 relative), exponent added to the high word.
template <int N>
__device__ __forceinline__ void exp_n(const double (&x)[N], double (&e)[N])
{
    double r[N], p[N];
    int k[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double xi = fmin(fmax(x[i], -708.0), 708.0);
        const double t = fma(xi, 1.4426950408889634074, 6755399441055744.0);
        k[i] = __double2loint(t);
        const double kd = t - 6755399441055744.0;
        r[i] = fma(kd, -1.90821492927058770002e-10, fma(kd, -6.93147180369123816490e-01, xi));
        p[i] = 2.08767569878680989792e-09;               // 1/12!
    }
    const double c[12] = {2.50521083854417187751e-08, 2.75573192239858906526e-07, 2.75573192239858906526e-06,
                          2.48015873015873015873e-05, 1.98412698412698412698e-04, 1.38888888888888888889e-03,
                          8.33333333333333333333e-03, 4.16666666666666666667e-02, 1.66666666666666666667e-01,
                          0.5, 1.0, 1.0};
#pragma unroll
    for (int j = 0; j < 12; ++j) {
#pragma unroll
        for (int i = 0; i < N; ++i) p[i] = fma(p[i], r[i], c[j]);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) e[i] = __hiloint2double(__double2hiint(p[i]) + (k[i] << 20), __double2loint(p[i]));
}

__device__ __forceinline__ V vfma(double a, V b, V c) { return V{fma(a, b.x, c.x), fma(a, b.y, c.y)}; }
__device__ __forceinline__ V vfma(V a, V b, V c) { return V{fma(a.x, b.x, c.x), fma(a.y, b.y, c.y)}; }
__device__ __forceinline__ V vadd(V a, V b) { return V{a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ V vsub(V a, V b) { return V{a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ V vmul(double a, V b) { return V{a * b.x, a * b.y}; }
__device__ __forceinline__ V vmul(V a, V b) { return V{a.x * b.x, a.y * b.y}; }
__device__ __forceinline__ V vexp(V a) { return V{exp_fast(a.x), exp_fast(a.y)}; }
__device__ __forceinline__ V zero_v() { return V{0.0, 0.0}; }
template <int GS>
__device__ __forceinline__ unsigned sp_even(unsigned aSP, unsigned k) { return aSP + k * (SP_SLOTS * GS * 8) + (k & 1u) * (GS * 8); }

// sum over the sub-groups of a warp (lanes with equal state pair); result in every lane
template <int GS>
__device__ __forceinline__ double sub_sum(double v)
{
#pragma unroll
    for (int o = GS / 2; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int GS>
__device__ __forceinline__ V sub_sum(V v) { return V{sub_sum<GS>(v.x), sub_sum<GS>(v.y)}; }

// Everything phase B does for one reaction and the two states of the lane.
// states of the lane that exist (tail group) and where the M_RATES outputs go
struct Out {
    long long s0;
    bool ok0, ok1;
};
__device__ __forceinline__ void put2(double* base, const IO& io, int width, const Out& o, int v, V x)
{
    if (o.ok0) put(base, io, width, o.s0, v, x.x);
    if (o.ok1) put(base, io, width, o.s0 + 1, v, x.y);
}

template <int GS, bool PM, int MODE, bool WSG>
__device__ __forceinline__ void reaction(const Mem<WSG>& mem, const Tables& tb, const Plan& pl, const IO& io, const Out& out,
                                         unsigned aSP, unsigned aRX,
                                         unsigned aRAW, unsigned aSC, int p, bool valid, bool three,
                                         const int4 q0, const int4 q1, const int4 q2, const int4 q3,
                                         const V T, const V logT, const V iT)
{
    constexpr int RB = GS * 8, RXB = RX_SLOTS * RB;
    const int nsp = tb.nsp, last = tb.nsp - 1;
    const double lnA = __hiloint2double(q0.y, q0.x), bexp = __hiloint2double(q0.w, q0.z);
    const double Ta = __hiloint2double(q1.y, q1.x), lnKc = __hiloint2double(q1.w, q1.z);
    const int fl = q2.x;
    // species slots: (offset of the even-slot base) / 16; nsp_f / last_f: the same for the empty
    // slot and the last species
    const unsigned s0 = q2.y & 0xFFFFu, s1 = (unsigned)q2.y >> 16, s2 = q2.z & 0xFFFFu;
    const unsigned s3 = (unsigned)q2.z >> 16, s4 = q2.w & 0xFFFFu, s5 = (unsigned)q2.w >> 16;
    const unsigned a0 = aSP + s0 * 16, a1 = aSP + s1 * 16, a2 = aSP + s2 * 16;
    const unsigned a3 = aSP + s3 * 16, a4 = aSP + s4 * 16, a5 = aSP + s5 * 16;
    const unsigned nsp_f = (sp_even<GS>(0u, (unsigned)nsp)) / 16, last_f = (sp_even<GS>(0u, (unsigned)last)) / 16;
    const bool isrev = fl & F_REV;

    // ---- species values and the arguments of kf, kr
    const V c0 = LDS(E_C * RB, a0), c1 = LDS(E_C * RB, a1), c3 = LDS(E_C * RB, a3), c4 = LDS(E_C * RB, a4);
    V c2{1.0, 1.0}, c5{1.0, 1.0};
    V sB = vsub(vadd(LDS(O_B * RB, a3 ^ RB), LDS(O_B * RB, a4 ^ RB)), vadd(LDS(O_B * RB, a0 ^ RB), LDS(O_B * RB, a1 ^ RB)));
    V sdB = vsub(vadd(LDS(E_DB * RB, a3), LDS(E_DB * RB, a4)), vadd(LDS(E_DB * RB, a0), LDS(E_DB * RB, a1)));
    V dH = vsub(vadd(LDS(O_HW * RB, a3 ^ RB), LDS(O_HW * RB, a4 ^ RB)), vadd(LDS(O_HW * RB, a0 ^ RB), LDS(O_HW * RB, a1 ^ RB)));
    if (three) {
        c2 = LDS(E_C * RB, a2);
        c5 = LDS(E_C * RB, a5);
        sB = vadd(sB, vsub(LDS(O_B * RB, a5 ^ RB), LDS(O_B * RB, a2 ^ RB)));
        sdB = vadd(sdB, vsub(LDS(E_DB * RB, a5), LDS(E_DB * RB, a2)));
        dH = vadd(dH, vsub(LDS(O_HW * RB, a5 ^ RB), LDS(O_HW * RB, a2 ^ RB)));
    }
    const V lnkf = vfma(bexp, logT, V{fma(-Ta, iT.x, lnA), fma(-Ta, iT.y, lnA)});
    const V lnkr{lnkf.x - sB.x - lnKc, lnkf.y - sB.y - lnKc};
    V kf, kr;

    // ---- pressure modification: PM_, and for the Jacobian gg, Xd, e1Fi.  All exponentials of a
    // reaction are evaluated in two batches (exp_n) so that the lane's two states and the
    // independent terms overlap.
    V PM_{1.0, 1.0}, gg{1.0, 1.0}, Xd{0.0, 0.0}, e1Fi{0.0, 0.0};
    const int mi = PM ? p - tb.first_pm : 0;
    const double* par = tb.pm_par + mi * NPAR;
    bool rates_done = false;
    if (PM) {
        V thd = LDS(Q_M * RB, aSC);
        // collider records {alpha - 1, species row offset, raw row}, four at a time (padded)
        const int e0 = __ldg(pl.eff_off + mi), e1_ = __ldg(pl.eff_off + mi + 1);
        for (int e = e0; e < e1_; e += 4) {
            const int4 r0 = __ldg(pl.eff + e), r1 = __ldg(pl.eff + e + 1), r2 = __ldg(pl.eff + e + 2), r3 = __ldg(pl.eff + e + 3);
            thd = vfma(__hiloint2double(r0.y, r0.x), LDS(E_C * RB, aSP + r0.z), thd);
            thd = vfma(__hiloint2double(r1.y, r1.x), LDS(E_C * RB, aSP + r1.z), thd);
            thd = vfma(__hiloint2double(r2.y, r2.x), LDS(E_C * RB, aSP + r2.z), thd);
            thd = vfma(__hiloint2double(r3.y, r3.x), LDS(E_C * RB, aSP + r3.z), thd);
        }
        if (fl & F_PDEP) {
            const int csp = __ldg(tb.pm_sp + mi);
            const V ctv = csp >= 0 ? LDS(E_C * RB, sp_even<GS>(aSP, (unsigned)csp)) : thd;
            const bool low = fl & F_LOW;
            const double p0 = par[0], p1 = par[1], p2 = par[2], p3 = par[3];
            const double ct[2] = {ctv.x, ctv.y}, Tt[2] = {T.x, T.y}, lT[2] = {logT.x, logT.y}, rT[2] = {iT.x, iT.y};
            double Pr[2], i1p[2], F[2], dpr[2], xd[2], g_[2], e1f[2];
            if (fl & F_TROE) {
                // batch 1: the Pr exponential and the three Troe exponentials (par[28], par[29] =
                // 1 / par[7], 1 / par[9]); a missing T2 term is masked out
                const double m2 = (fl & F_TROE_T2) ? 1.0 : 0.0;
                const double x1[8] = {p0 + p1 * lT[0] - p2 * rT[0], p0 + p1 * lT[1] - p2 * rT[1],
                                      Tt[0] * par[28], Tt[1] * par[28], Tt[0] * par[29], Tt[1] * par[29],
                                      m2 * par[10] * rT[0], m2 * par[10] * rT[1]};
                double y1[8];
                exp_n<8>(x1, y1);
                const double iln10 = 0.43429448190325182765;
                double fa[2];
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const double e1 = y1[g], e3 = y1[2 + g], e1t = y1[4 + g], e2 = m2 * y1[6 + g];
                    Pr[g] = ct[g] * e1;
                    const double dpr4 = p3 + p2 * rT[g] - 1.0;
                    dpr[g] = p1 + p2 * rT[g] - 1.0;
                    i1p[g] = __drcp_rn(1.0 + Pr[g]);
                    if (low) { xd[g] = dpr4 * rT[g] * i1p[g]; g_[g] = i1p[g]; }
                    else { xd[g] = -Pr[g] * dpr4 * rT[g] * i1p[g]; g_[g] = -Pr[g] * i1p[g]; }
                    e1f[g] = e1;
                    const double Fc = par[6] * e3 + par[8] * e1t + e2;
                    const double dF = par[11] * e3 - par[12] * e1t + par[13] * rT[g] * rT[g] * e2;
                    const double lnFc = log(fmax(Fc, 1.0e-300));
                    const double lF = lnFc * iln10, lP = log(fmax(Pr[g], 1.0e-300)) * iln10;
                    const double A = lP - 0.67 * lF - 0.4;
                    const double Bq = 0.806 - 1.1762 * lF - 0.14 * lP;
                    const double rB = __drcp_rn(Bq), rFc = __drcp_rn(Fc);
                    const double t = A * rB;
                    const double rq = __drcp_rn(fma(t, t, 1.0));
                    const double lnF_AB = 2.0 * lnFc * t * (rB * rq) * (rB * rq);
                    fa[g] = lnFc * rq;
                    xd[g] += (rFc * rq - lnF_AB * (-0.67 * iln10 * Bq + 1.1762 * iln10 * A) * rFc) * dF
                             - lnF_AB * (Bq * iln10 + 0.14 * iln10 * A) * dpr[g] * rT[g];
                    g_[g] -= lnF_AB * (Bq * iln10 + A * 0.14 * iln10);
                }
                // batch 2: the broadening factor F and the two rate constants
                const double x2[6] = {fa[0], fa[1], lnkf.x, lnkf.y, lnkr.x, lnkr.y};
                double y2[6];
                exp_n<6>(x2, y2);
                F[0] = y2[0]; F[1] = y2[1];
                kf = V{y2[2], y2[3]};
                kr = V{y2[4], y2[5]};
                rates_done = true;
            } else {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const double e1 = exp_fast(p0 + p1 * lT[g] - p2 * rT[g]);
                    Pr[g] = ct[g] * e1;
                    const double dpr4 = p3 + p2 * rT[g] - 1.0;
                    dpr[g] = p1 + p2 * rT[g] - 1.0;
                    i1p[g] = 1.0 / (1.0 + Pr[g]);
                    if (low) { xd[g] = dpr4 * rT[g] * i1p[g]; g_[g] = i1p[g]; }
                    else { xd[g] = -Pr[g] * dpr4 * rT[g] * i1p[g]; g_[g] = -Pr[g] * i1p[g]; }
                    F[g] = 1.0;
                    e1f[g] = e1;
                }
                if (fl & F_SRI) {
#pragma unroll 1
                    for (int g = 0; g < 2; ++g) {
                        const double lP = log10_clamped(Pr[g]);
                        const double X = 1.0 / (1.0 + lP * lP);
                        F[g] = pow(par[14] * exp(-par[15] * rT[g]) + exp(-Tt[g] / par[16]), X);
                        if (fl & F_SRI5) F[g] *= par[17] * pow(Tt[g], par[18]);
                        const double two_iln10 = 0.86858896380650365530;
                        const double eb = exp(par[23] * rT[g]), ec = exp(Tt[g] / par[25]);
                        const double den = par[26] * eb + ec;
                        xd[g] += X * ((par[22] * rT[g] * rT[g] * eb - par[24] * ec) / den
                                      - X * two_iln10 * lP * dpr[g] * log(den) * rT[g]);
                        if (fl & F_SRI5_DT) xd[g] += par[27] * rT[g];
                        g_[g] -= X * X * two_iln10 * lP * log(par[19] * exp(par[20] * rT[g]) + exp(Tt[g] / par[21]));
                    }
                }
            }
            double pm_[2];
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const double Fi = F[g] * i1p[g];
                pm_[g] = low ? Fi * Pr[g] : Fi;
                e1f[g] *= Fi;
            }
            PM_ = V{pm_[0], pm_[1]}; gg = V{g_[0], g_[1]}; Xd = V{xd[0], xd[1]}; e1Fi = V{e1f[0], e1f[1]};
        } else {
            PM_ = thd;
        }
    }

    // ---- rate constants and rates of progress
    if (!rates_done) {
        const double x2[4] = {lnkf.x, lnkf.y, lnkr.x, lnkr.y};
        double y2[4];
        exp_n<4>(x2, y2);
        kf = V{y2[0], y2[1]};
        kr = V{y2[2], y2[3]};
    }
    if (!isrev) kr = V{0.0, 0.0};
    if (fl & F_NEGA) { kf = V{-kf.x, -kf.y}; kr = V{-kr.x, -kr.y}; }   // A < 0 (rs:108-141)
    V f = vmul(kf, vmul(c0, c1)), r = vmul(kr, vmul(c3, c4));
    if (three) { f = vmul(f, c2); r = vmul(r, c5); }
    const V net = vsub(f, r);
    if (MODE == M_RATES && valid) {
        const int4 ro = __ldg(pl.rx_out + p);
        if (io.fwd) put2(io.fwd, io, tb.nr, out, ro.x, f);
        if (io.rev && isrev) put2(io.rev, io, tb.nrev, out, ro.y, r);
        if (PM && io.pm) put2(io.pm, io, tb.npd, out, ro.z, PM_);
    }
    if (!jac_like(MODE)) {
        // only the net rate is needed for the species rates
        STS_IF(RX_NET * RB, valid, aRX + p * RXB, PM ? vmul(net, PM_) : net);
        return;
    }
    V pmt{0.0, 0.0};
    if (PM && (fl & F_PMT)) pmt = vmul(gg, net);

    // ---- Jacobian scalars
    const double nre = (double)((fl >> NRE_SHIFT) & 15), npr = (double)((fl >> NPR_SHIFT) & 15);
    const V rho_inv = LDS(Q_RHOINV * RB, aSC), mwr_ = LDS(Q_MWR * RB, aSC), nmwr{-mwr_.x, -mwr_.y};
    const double extra = (PM && (fl & F_EFFN1)) ? 1.0 : 0.0;
    const double n1 = nre + extra, n2 = npr + extra, omre = 1.0 - nre, ompr = 1.0 - npr;
    V tT, X1, X2;
    {
        const double fv[2] = {f.x, f.y}, rv[2] = {r.x, r.y}, nv[2] = {net.x, net.y}, Tv[2] = {T.x, T.y};
        const double iv[2] = {iT.x, iT.y}, sd[2] = {sdB.x, sdB.y}, pmv[2] = {PM_.x, PM_.y};
        const double xdv[2] = {Xd.x, Xd.y}, ri[2] = {rho_inv.x, rho_inv.y}, nm[2] = {nmwr.x, nmwr.y};
        const double ef[2] = {e1Fi.x, e1Fi.y};
        double pt[2] = {pmt.x, pmt.y}, t_[2], x1[2], x2[2];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const double dk = bexp + Ta * iv[g];
            // irreversible: r = 0 makes this f * (dk + 1 - nre)          (cj:1461-1523)
            const double elem = nv[g] * dk + fv[g] * omre - rv[g] * (ompr - Tv[g] * sd[g]);
            double t;
            if (PM) {
                if (fl & F_PDEP) t = (pmv[g] * xdv[g] * nv[g] + pmv[g] * iv[g] * elem) * ri[g];
                else t = (-pmv[g] * nv[g] * iv[g] + pmv[g] * iv[g] * elem) * ri[g];
            } else {
                t = iv[g] * elem * ri[g];
            }
            t_[g] = (fl & F_NO_T) ? 0.0 : t;
            double inner = n1 * fv[g] - n2 * rv[g];
            if (PM && (fl & F_PMT_INJ)) inner += pt[g];
            const double jy = nm[g] * pmv[g] * inner;
            if (PM && (fl & F_PMT_INJ)) pt[g] *= ef[g];
            x1[g] = jy;
            x2[g] = -jy;
            if (PM) { x1[g] += par[5] * pt[g]; x2[g] -= par[4] * pt[g]; }
        }
        tT = V{t_[0], t_[1]}; X1 = V{x1[0], x1[1]}; X2 = V{x2[0], x2[1]}; pmt = V{pt[0], pt[1]};
    }
    const V pk = PM ? vmul(PM_, kf) : kf;
    const V prv = PM ? V{-PM_.x * kr.x, -PM_.y * kr.y} : V{-kr.x, -kr.y};

    // ---- d(rate)/dC values: to their raw rows, or folded into X2 for the last species
#define PJ_EMIT(SLOT, DST, EXPR)                                         \
    if ((SLOT) != nsp_f) {                                               \
        const V d_ = (EXPR);                                             \
        if ((SLOT) == last_f) X2 = vsub(X2, d_);                         \
        else if (valid) STS(0, aRAW + (DST) * RB, d_);                   \
    }
    if (three) {
        PJ_EMIT(s0, q3.x & 0xFFFFu, vmul(pk, vmul(c1, c2)))
        PJ_EMIT(s1, (unsigned)q3.x >> 16, vmul(pk, vmul(c0, c2)))
        PJ_EMIT(s2, q3.y & 0xFFFFu, vmul(pk, vmul(c0, c1)))
        if (isrev) {
            PJ_EMIT(s3, (unsigned)q3.y >> 16, vmul(prv, vmul(c4, c5)))
            PJ_EMIT(s4, q3.z & 0xFFFFu, vmul(prv, vmul(c3, c5)))
            PJ_EMIT(s5, (unsigned)q3.z >> 16, vmul(prv, vmul(c3, c4)))
        }
    } else {
        PJ_EMIT(s0, q3.x & 0xFFFFu, vmul(pk, c1))
        PJ_EMIT(s1, (unsigned)q3.x >> 16, vmul(pk, c0))
        if (isrev) {
            PJ_EMIT(s3, (unsigned)q3.y >> 16, vmul(prv, c4))
            PJ_EMIT(s4, q3.z & 0xFFFFu, vmul(prv, c3))
        }
    }
#undef PJ_EMIT
    if (PM && valid) {
        if (fl & F_EFF_SLOTS) {
            const int e0 = __ldg(pl.eff_off + mi), e1_ = __ldg(pl.eff_off + mi + 1);
            for (int e = e0; e < e1_; e += 4) {
                const int4 r0 = __ldg(pl.eff + e), r1 = __ldg(pl.eff + e + 1), r2 = __ldg(pl.eff + e + 2), r3 = __ldg(pl.eff + e + 3);
                const int none = tb.nraw + 1;          // padding records and colliders without a raw row
                STS_IF(0, r0.w != none, aRAW + r0.w * RB, vmul(__hiloint2double(r0.y, r0.x), pmt));
                STS_IF(0, r1.w != none, aRAW + r1.w * RB, vmul(__hiloint2double(r1.y, r1.x), pmt));
                STS_IF(0, r2.w != none, aRAW + r2.w * RB, vmul(__hiloint2double(r2.y, r2.x), pmt));
                STS_IF(0, r3.w != none, aRAW + r3.w * RB, vmul(__hiloint2double(r3.y, r3.x), pmt));
            }
        }
        if (fl & F_WANT_PMT) STS(0, aRAW + ((unsigned)q3.w >> 16) * RB, pmt);
    }
    if (valid) {
        const unsigned ar = aRX + p * RXB;
        STS(RX_NET * RB, ar, PM ? vmul(net, PM_) : net);
        STS(RX_TT * RB, ar, tT);
        STS(RX_X1 * RB, ar, X1);
        STS(RX_X2 * RB, ar, X2);
        STS(RX_DH * RB, ar, dH);
    }
}

// Phase B for one reaction without pressure modification (the common case) and the two states
// of the lane: no branches except `three` (warp-uniform) and the rare last-species fold.
template <int GS, int MODE, bool SPECIAL, bool WSG>
__device__ __forceinline__ void reaction_plain(const Mem<WSG>& mem, const Tables& tb, const Plan& pl, const IO& io, const Out& out,
                                               unsigned aSP, unsigned aRX,
                                               unsigned aRAW, unsigned aSC, unsigned aPL, int p, bool valid, bool three,
                                               const int4 q0, const int4 q1, const int4 q2, const int4 q3,
                                               const V T, const V logT, const V iT)
{
    constexpr int RB = GS * 8, RXB = RX_SLOTS * RB;
    const double lnA = __hiloint2double(q0.y, q0.x), bexp = __hiloint2double(q0.w, q0.z);
    const double Ta = __hiloint2double(q1.y, q1.x), lnKc = __hiloint2double(q1.w, q1.z);
    const int fl = q2.x;
    const unsigned s0 = q2.y & 0xFFFFu, s1 = (unsigned)q2.y >> 16, s3 = (unsigned)q2.z >> 16, s4 = q2.w & 0xFFFFu;
    const unsigned a0 = aSP + s0 * 16, a1 = aSP + s1 * 16, a3 = aSP + s3 * 16, a4 = aSP + s4 * 16;
    V c0 = LDS(E_C * RB, a0), c1 = LDS(E_C * RB, a1), c3 = LDS(E_C * RB, a3), c4 = LDS(E_C * RB, a4);
    V sB = vsub(vadd(LDS(O_B * RB, a3 ^ RB), LDS(O_B * RB, a4 ^ RB)), vadd(LDS(O_B * RB, a0 ^ RB), LDS(O_B * RB, a1 ^ RB)));
    V sdB = vsub(vadd(LDS(E_DB * RB, a3), LDS(E_DB * RB, a4)), vadd(LDS(E_DB * RB, a0), LDS(E_DB * RB, a1)));
    V dH = vsub(vadd(LDS(O_HW * RB, a3 ^ RB), LDS(O_HW * RB, a4 ^ RB)), vadd(LDS(O_HW * RB, a0 ^ RB), LDS(O_HW * RB, a1 ^ RB)));
    V c2{1.0, 1.0}, c5{1.0, 1.0};
    const unsigned s2 = q2.z & 0xFFFFu, s5 = (unsigned)q2.w >> 16;
    if (three) {
        const unsigned a2 = aSP + s2 * 16, a5 = aSP + s5 * 16;
        c2 = LDS(E_C * RB, a2);
        c5 = LDS(E_C * RB, a5);
        sB = vadd(sB, vsub(LDS(O_B * RB, a5 ^ RB), LDS(O_B * RB, a2 ^ RB)));
        sdB = vadd(sdB, vsub(LDS(E_DB * RB, a5), LDS(E_DB * RB, a2)));
        dH = vadd(dH, vsub(LDS(O_HW * RB, a5 ^ RB), LDS(O_HW * RB, a2 ^ RB)));
    }
    V lnkf = vfma(bexp, logT, V{fma(-Ta, iT.x, lnA), fma(-Ta, iT.y, lnA)});
    V dk{0.0, 0.0};                                           // d ln kf / d ln T
    if (SPECIAL) dk = V{fma(Ta, iT.x, bexp), fma(Ta, iT.y, bexp)};
    if (SPECIAL && (fl & F_PLOG)) {
        // rate constant of a PLOG reaction (rs:598-632) and its temperature derivative
        // (cj:1687-1850): the Arrhenius set of the first / last pressure outside the table,
        // linear interpolation of ln kf in ln P between two pressures
        const int o0 = __ldg(tb.plog_off + p), o1 = __ldg(tb.plog_off + p + 1);
        const V Pv = LDS(S_P * RB, aPL), lP = LDS(Q_LNP * RB, aSC);
        const double Ps[2] = {Pv.x, Pv.y}, lPs[2] = {lP.x, lP.y}, lTs[2] = {logT.x, logT.y}, rTs[2] = {iT.x, iT.y};
        double lk[2], dkk[2];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            int e = o0;
            while (e < o1 && Ps[g] > __ldg(tb.plog_par + 8 * e)) ++e;
            const bool mid = e > o0 && e < o1;
            const double* q = tb.plog_par + 8 * (e > o0 ? e - 1 : o0);
            const double A1 = __ldg(q + 1), b1 = __ldg(q + 2), E1 = __ldg(q + 3);
            double k = fma(b1, lTs[g], fma(-E1, rTs[g], A1)), d = fma(E1, rTs[g], b1);
            if (mid) {
                const double k2 = fma(__ldg(q + 10), lTs[g], fma(-__ldg(q + 11), rTs[g], __ldg(q + 9)));
                const double w = (lPs[g] - __ldg(q + 4)) * __ldg(q + 5);
                k = fma(k2 - k, w, k);
                d = fma(fma(__ldg(q + 7), rTs[g], __ldg(q + 6)), w, d);
            }
            lk[g] = k; dkk[g] = d;
        }
        lnkf = V{lk[0], lk[1]};
        dk = V{dkk[0], dkk[1]};
    }
    V rat{1.0, 1.0};
    if (SPECIAL && (fl & F_CHEB)) {
        // Chebyshev rate constant (rs:149-251) and its temperature derivative (cj:1532-1684).  The
        // generated eval_jacob evaluates the rate constant of its species part with reduced
        // variables printed to 16 digits, eval_rxn_rates with 8 digits: lnkf follows the former,
        // `rat` = kf(rates) / kf(Jacobian) rescales the rates of progress
        const double* cq = tb.cheb_par + __ldg(tb.cheb_off + p);
        const int n_t = (int)__ldg(cq), n_p = (int)__ldg(cq + 1);
        const double* c8 = cq + 12;
        const double* c16 = c8 + n_t * n_p;
        const V lP = LDS(Q_LNP * RB, aSC);
        const double lPs[2] = {lP.x, lP.y}, rTs[2] = {iT.x, iT.y};
        const double ln10 = 2.30258509299404568402, iln10 = 0.43429448190325182765;
        double lr[2], lj[2], dkk[2];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const double l10p = lPs[g] * iln10;
            const double tr8 = (2.0 * rTs[g] - __ldg(cq + 2)) / __ldg(cq + 3), pr8 = (2.0 * l10p - __ldg(cq + 4)) / __ldg(cq + 5);
            const double tr16 = (2.0 * rTs[g] - __ldg(cq + 6)) / __ldg(cq + 7), pr16 = (2.0 * l10p - __ldg(cq + 8)) / __ldg(cq + 9);
            double t8a = 1.0, t8b = tr8, t16a = 1.0, t16b = tr16, ua = 0.0, ub = 1.0;
            double s8 = 0.0, s16 = 0.0, su = 0.0;
            for (int i = 0; i < n_t; ++i) {
                const double* r8 = c8 + i * n_p;
                const double* r16 = c16 + i * n_p;
                // row i: sum_j c[i][j] T_j(Pred) for both reduced pressures
                const double c0 = __ldg(r8), c1 = __ldg(r8 + 1);
                double pa8 = 1.0, pb8 = pr8, pa16 = 1.0, pb16 = pr16;
                double dp8 = fma(pr8, c1, c0), dp16 = fma(pr16, c1, c0), dpu = fma(pr16, __ldg(r16 + 1), __ldg(r16));
                for (int j = 2; j < n_p; ++j) {
                    const double n8 = fma(2.0 * pr8, pb8, -pa8), n16 = fma(2.0 * pr16, pb16, -pa16);
                    pa8 = pb8; pb8 = n8; pa16 = pb16; pb16 = n16;
                    const double cc = __ldg(r8 + j);
                    dp8 = fma(cc, n8, dp8); dp16 = fma(cc, n16, dp16); dpu = fma(__ldg(r16 + j), n16, dpu);
                }
                // T_i(Tred) for both reduced temperatures, U_{i-1}(Tred) for the derivative
                double T8 = 1.0, T16 = 1.0, U = 0.0;
                if (i == 1) { T8 = tr8; T16 = tr16; U = 1.0; }
                else if (i > 1) {
                    T8 = fma(2.0 * tr8, t8b, -t8a); t8a = t8b; t8b = T8;
                    T16 = fma(2.0 * tr16, t16b, -t16a); t16a = t16b; t16b = T16;
                    U = fma(2.0 * tr16, ub, -ua); ua = ub; ub = U;
                }
                s8 = fma(dp8, T8, s8); s16 = fma(dp16, T16, s16); su = fma(dpu, U, su);
            }
            lr[g] = ln10 * s8; lj[g] = ln10 * s16;
            dkk[g] = su * __ldg(cq + 10) * rTs[g];
        }
        lnkf = V{lj[0], lj[1]};
        dk = V{dkk[0], dkk[1]};
        rat = V{exp_fast(lr[0] - lj[0]), exp_fast(lr[1] - lj[1])};
    }
    const double ex[4] = {lnkf.x, lnkf.y, lnkf.x - sB.x - lnKc, lnkf.y - sB.y - lnKc};
    double ev[4];
    exp_n<4>(ex, ev);
    const bool isrev = fl & F_REV;
    // A < 0 (rs:108-141): the table holds log|A|, the sign goes onto both rate constants
    const double sg = (SPECIAL && (fl & F_NEGA)) ? -1.0 : 1.0;
    const V kf{sg * ev[0], sg * ev[1]};
    const V kr{isrev ? sg * ev[2] : 0.0, isrev ? sg * ev[3] : 0.0};
    // d(rate)/dC per occupied slot; f = d0 * c0, r = -d3 * c3
    V o0 = c1, o1 = c0, o3 = c4, o4 = c3, o2 = vmul(c0, c1), o5 = vmul(c3, c4);
    if (three) { o0 = vmul(c1, c2); o1 = vmul(c0, c2); o3 = vmul(c4, c5); o4 = vmul(c3, c5); }
    const V d0 = vmul(kf, o0), d1 = vmul(kf, o1), d3 = vmul(kr, o3), d4 = vmul(kr, o4);
    V f = vmul(d0, c0), r = vmul(d3, c3);
    if (SPECIAL && (fl & F_CHEB)) { f = vmul(f, rat); r = vmul(r, rat); }
    const V net = vsub(f, r);
    if (MODE == M_RATES && valid) {
        const int4 ro = __ldg(pl.rx_out + p);
        if (io.fwd) put2(io.fwd, io, tb.nr, out, ro.x, f);
        if (io.rev && isrev) put2(io.rev, io, tb.nrev, out, ro.y, r);
    }
    if (!jac_like(MODE)) {
        STS_IF(RX_NET * RB, valid, aRX + p * RXB, net);
        return;
    }
    const double nre = (double)((fl >> NRE_SHIFT) & 15), npr = (double)((fl >> NPR_SHIFT) & 15);
    const double omre = 1.0 - nre, ompr = 1.0 - npr;
    const V rho_inv = LDS(Q_RHOINV * RB, aSC), mwr_ = LDS(Q_MWR * RB, aSC), nmwr{-mwr_.x, -mwr_.y};
    if (!SPECIAL) dk = V{fma(Ta, iT.x, bexp), fma(Ta, iT.y, bexp)};
    // irreversible: r = 0 makes this f * (dk + 1 - nre)          (cj:1461-1523)
    V elem = vfma(net, dk, vmul(omre, f));
    elem = V{elem.x - r.x * (ompr - T.x * sdB.x), elem.y - r.y * (ompr - T.y * sdB.y)};
    const double tmask = (fl & F_NO_T) ? 0.0 : 1.0;
    const V tT = vmul(tmask, vmul(vmul(iT, rho_inv), elem));
    const V X1 = vmul(nmwr, V{nre * f.x - npr * r.x, nre * f.y - npr * r.y});
    V X2{-X1.x, -X1.y};
    const V n3{-d3.x, -d3.y}, n4{-d4.x, -d4.y};
    V d2 = zero_v(), n5 = zero_v();
    if (three) { d2 = vmul(kf, o2); n5 = vmul(kr, o5); n5 = V{-n5.x, -n5.y}; }
    if (fl & F_HAS_LAST) {
        // a slot holding the last species has no column: its derivative joins the W_j / W_N term
        const unsigned last = sp_even<GS>(0u, (unsigned)(tb.nsp - 1)) / 16;
        if (s0 == last) X2 = vsub(X2, d0);
        if (s1 == last) X2 = vsub(X2, d1);
        if (s2 == last) X2 = vsub(X2, d2);
        if (s3 == last) X2 = vsub(X2, n3);
        if (s4 == last) X2 = vsub(X2, n4);
        if (s5 == last) X2 = vsub(X2, n5);
    }
    // raw rows; slots without one carry the scratch row index and store nothing
    const unsigned none = tb.nraw + 1;
    auto emit = [&](unsigned dst, V d) { STS_IF(0, valid && dst != none, aRAW + dst * RB, d); };
    emit(q3.x & 0xFFFFu, d0);
    emit((unsigned)q3.x >> 16, d1);
    emit((unsigned)q3.y >> 16, n3);
    emit(q3.z & 0xFFFFu, n4);
    if (three) {
        emit(q3.y & 0xFFFFu, d2);
        emit((unsigned)q3.z >> 16, n5);
    }
    const unsigned ar = aRX + p * RXB;
    STS_IF(RX_NET * RB, valid, ar, net);
    STS_IF(RX_TT * RB, valid, ar, tT);
    STS_IF(RX_X1 * RB, valid, ar, X1);
    STS_IF(RX_X2 * RB, valid, ar, X2);
    STS_IF(RX_DH * RB, valid, ar, dH);
}

template <int GS, int MAXT, int MODE, bool WSG = false>
__global__ void __launch_bounds__(MAXT, 1)
k_eval(const __grid_constant__ Tables tb, const __grid_constant__ Plan pl, const __grid_constant__ IO io)
{
    extern __shared__ __align__(16) double smem[];
    constexpr int NSUB = 64 / GS;          // table items a warp works on at the same time
    constexpr int NPR = GS / 2;            // state pairs = lanes per sub-group
    constexpr int RB = GS * 8, SPB = SP_SLOTS * RB, RXB = RX_SLOTS * RB;
    constexpr int SCB = 8 * RB;            // one buffer of the phase A0 scalars
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nw = pl.nw;
    const int sub = lane / NPR, pr = lane % NPR;
    const int nsp = tb.nsp, last = tb.nsp - 1;
    // working set: shared memory, or (WSG) this block's scratch area in global memory
    const unsigned ws0 = WSG ? 0u : (unsigned)__cvta_generic_to_shared(smem);
    const Mem<WSG> mem{WSG ? io.ws + (size_t)blockIdx.x * ((size_t)pl.total * 8) : nullptr};
    // 32-bit addresses (WSG: offsets) of this lane's state pair in each region
    const unsigned sb = ws0 + pr * 16;
    if ((ws0 + pl.oSP * 8) & (2 * RB - 1)) __trap();   // E ^ RB needs this
    const unsigned aSP = sb + pl.oSP * 8, aRX = sb + pl.oRX * 8, aRAW = sb + pl.oRAW * 8;
    const unsigned aSC0 = sb + pl.oSC * 8, aPA = sb + pl.oPA * 8;
    const unsigned aSD = aSC0 + 2 * SCB;   // scalars derived in phase DE
    const unsigned aCF = ws0 + pl.oCF * 8;   // column factors, [col][2]
    const V zero{0.0, 0.0};

    // the column factors (848 bytes for 53 species) live in shared memory: they are needed once
    // per Jacobian element and do not survive in the small L1 next to the streamed tables
    for (int i = tid; i < nsp; i += blockDim.x) {
        const double2 c = __ldg(pl.colfac + i);
        STS(0, aCF + i * 16, V{c.x, c.y});
    }

    // rows that never change: the empty reaction slot, two all-zero reactions and raw rows (the
    // padding of the gather lists; one on either half of a bank line)
    if (warp == 0 && sub == 0) {
        const unsigned a = sp_even<GS>(aSP, (unsigned)nsp), o = a ^ RB;
        STS(E_C * RB, a, V{1.0, 1.0});
        STS(O_B * RB, o, zero); STS(E_DB * RB, a, zero); STS(O_HW * RB, o, zero);
        STS(E_WA * RB, a, zero); STS(O_WB * RB, o, zero); STS(E_WT * RB, a, zero); STS(O_CP * RB, o, zero);
        for (int z = 0; z < 2; ++z) {
            const unsigned ar = aRX + (tb.nr + z) * RXB;
            STS(RX_NET * RB, ar, zero); STS(RX_TT * RB, ar, zero); STS(RX_X1 * RB, ar, zero);
            STS(RX_X2 * RB, ar, zero); STS(RX_DH * RB, ar, zero);
        }
        STS(0, aRAW + tb.nraw * RB, zero);
        STS(0, aRAW + (tb.nraw + 1) * RB, zero);
    }

    const bool sf = io.jac_layout != 0;
    const bool vec_ok = sf && ((io.jac_ld & 1) == 0) && ((reinterpret_cast<unsigned long long>(io.jac) & 15) == 0);
    // values per state of the output: the dense Jacobian, or the factored record
    const long long nn = MODE == M_FACT ? (long long)nsp + 3 * last + pl.fac_nnz : (long long)nsp * nsp;
    const long long ngroups = ((long long)io.n + GS - 1) / GS;
    // element e of state s: SoA jac[e * ld + s], AoS jac[s * nn + e]
    const unsigned ld8 = sf ? (unsigned)(io.jac_ld * 8) : 8u;
    const long long second = sf ? 8 : nn * 8;

    // phase A0 of group g into scalar buffer b: mass fractions (to the C slot of the species
    // rows, where A1 turns them into concentrations), Y_N, mean molecular weight, density
    auto phase_a0 = [&](long long g, int b) {
        const long long s0 = g * GS + 2 * pr;
        const long long i0 = s0 < io.n ? s0 : (long long)io.n - 1, i1 = s0 + 1 < io.n ? s0 + 1 : (long long)io.n - 1;
        const double* y0 = io.y + i0 * io.y_ss;
        const double* y1 = io.y + i1 * io.y_ss;
        // in_conc (eval_rxn_rates / get_rxn_pres_mod entry points): the row holds T and all NSP
        // concentrations; rho = sum C_k W_k, mw_avg = rho / sum C_k
        const bool inc = MODE == M_RATES && io.in_conc;
        V sumY = zero, sumYW = zero;
        // four species per pass, their loads (HBM latency: this is the only per-state input) issued
        // before the first is used; same summation order as one species at a time
        const int kend = inc ? nsp : last;
        const double* wtab = inc ? tb.sp_w : tb.sp_iw;
        for (int k = sub; k < kend; k += 4 * NSUB) {
            V Yk[4];
            double wk[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int kk = k + u * NSUB;
                const bool in = kk < kend;
                const long long o = (long long)((in ? kk : k) + 1) * io.y_sv;
                Yk[u] = V{y0[o], y1[o]};
                wk[u] = __ldg(wtab + (in ? kk : k));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int kk = k + u * NSUB;
                if (kk < kend) {
                    STS(E_Y * RB, sp_even<GS>(aSP, (unsigned)kk), Yk[u]);
                    sumY = vadd(sumY, Yk[u]);
                    sumYW = vfma(wk[u], Yk[u], sumYW);
                }
            }
        }
        sumY = sub_sum<GS>(sumY);
        sumYW = sub_sum<GS>(sumYW);
        if (sub == 0) {
            const double T[2] = {y0[0], y1[0]};
            double P[2] = {io.pres[i0], io.pres[i1]};
            const double yN[2] = {1.0 - sumY.x, 1.0 - sumY.y};
            const double sw[2] = {sumYW.x, sumYW.y}, sy[2] = {sumY.x, sumY.y};
            double o[8][2];
#pragma unroll
            for (int g2 = 0; g2 < 2; ++g2) {
                const double mw = inc ? sw[g2] / sy[g2] : 1.0 / (sw[g2] + yN[g2] * __ldg(tb.sp_iw + last));
                double rho = inc ? sw[g2] : P[g2] * mw / (tb.ru * T[g2]);
                if (!jac_like(MODE) && io.conv) {
                    // constant volume: the caller's variable is the density, the pressure follows (rs:1708-1800)
                    rho = P[g2];
                    P[g2] = rho * tb.ru * T[g2] / mw;
                }
                const double rho_inv = 1.0 / rho;
                if (MODE == M_RATES && io.scal3 && (g2 ? s0 + 1 : s0) < io.n) {
                    double* q = io.scal3 + (g2 ? s0 + 1 : s0) * 3;
                    q[0] = yN[g2]; q[1] = mw; q[2] = rho;
                }
                o[Q_T][g2] = T[g2]; o[Q_LOGT][g2] = log(T[g2]); o[Q_IT][g2] = 1.0 / T[g2];
                o[Q_RHO][g2] = rho; o[Q_RHOINV][g2] = rho_inv;
                o[Q_LNP][g2] = (tb.nplog | tb.ncheb) ? log(P[g2]) : 0.0; o[Q_MWR][g2] = mw * rho_inv;
                o[Q_M][g2] = P[g2] / (tb.ru * T[g2]);
            }
            const unsigned a = aSC0 + b * SCB;
            if (!inc) STS(E_Y * RB, sp_even<GS>(aSP, (unsigned)last), V{yN[0], yN[1]});
            STS(Q_T * RB, a, V{o[Q_T][0], o[Q_T][1]});
            STS(Q_LOGT * RB, a, V{o[Q_LOGT][0], o[Q_LOGT][1]});
            STS(Q_IT * RB, a, V{o[Q_IT][0], o[Q_IT][1]});
            STS(Q_RHO * RB, a, V{o[Q_RHO][0], o[Q_RHO][1]});
            STS(Q_RHOINV * RB, a, V{o[Q_RHOINV][0], o[Q_RHOINV][1]});
            STS(Q_LNP * RB, a, V{o[Q_LNP][0], o[Q_LNP][1]});
            STS(Q_MWR * RB, a, V{o[Q_MWR][0], o[Q_MWR][1]});
            STS(Q_M * RB, a, V{o[Q_M][0], o[Q_M][1]});
            if (tb.nplog) STS(S_P * RB, aSD + b * RB, V{P[0], P[1]});
        }
    };

    if (warp == 0 && (long long)blockIdx.x < ngroups) phase_a0(blockIdx.x, 0);
    __syncthreads();

#ifdef PJ_PHASE_CLOCKS
    long long clk[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#endif
#ifdef PJ_PHASE_CLOCKS          // development build only (tools/phase_clocks.py)
#define PJ_TICK(i) { const long long tn_ = clock64(); clk[i] += tn_ - tprev; tprev = tn_; }
#else
#define PJ_TICK(i)
#endif
    int buf = 0;
    for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x, buf ^= 1) {
        const long long s0 = grp * GS + 2 * pr;           // first state of this lane
        const bool ok0 = s0 < io.n, ok1 = s0 + 1 < io.n;
        const Out out{s0, ok0, ok1};
        char* const out0 = reinterpret_cast<char*>(sf ? io.jac + s0 : io.jac + s0 * nn);
        const unsigned aSC = aSC0 + buf * SCB;
        // fast: both states of the lane in range and 16-byte stores possible (all but tail groups)
        const bool fast = vec_ok && ok1;
        const bool nostore = PJ_SKIP(128);      // development builds only
        auto store = [&](unsigned e, V v, bool on) {
            char* o = out0 + (unsigned long long)e * ld8;
            if (nostore) on = false;
            if (fast) {
                // with the working set in global memory the Jacobian is stored with the streaming hint, so
                // that it does not displace the working sets from L2 (measured: +12 % / +18 % on the
                // USC-II- / n-heptane-sized mechanisms, -1 % with the working set in shared memory)
                if (WSG)
                    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\t@q st.global.cs.v2.f64 [%0], {%1, %2};\n\t}"
                                 ::"l"(o), "d"(v.x), "d"(v.y), "r"((int)on) : "memory");
                else
                    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\t@q st.global.v2.f64 [%0], {%1, %2};\n\t}"
                                 ::"l"(o), "d"(v.x), "d"(v.y), "r"((int)on) : "memory");
            } else {
                if (on && ok0) *reinterpret_cast<double*>(o) = v.x;
                if (on && ok1) *reinterpret_cast<double*>(o + second) = v.y;
            }
        };

        // ------------------------------------------------------------ phase A1: species thermo
        if (!PJ_SKIP(1)) {
            const V T = LDS(Q_T * RB, aSC), logT = LDS(Q_LOGT * RB, aSC), iT = LDS(Q_IT * RB, aSC);
            const V rho = LDS(Q_RHO * RB, aSC);
            const double Tv[2] = {T.x, T.y}, lT[2] = {logT.x, logT.y}, rT[2] = {iT.x, iT.y}, rh[2] = {rho.x, rho.y};
            double cpavg[2] = {0.0, 0.0}, wdcp[2] = {0.0, 0.0};
            for (int k = warp * NSUB + sub; k < nsp; k += nw * NSUB) {
                const unsigned a = sp_even<GS>(aSP, (unsigned)k), o = a ^ RB;
                const V Yv = LDS(E_Y * RB, a);
                const double iw = __ldg(tb.sp_iw + k), ruw = __ldg(tb.sp_ruw + k), wk = __ldg(tb.sp_w + k);
                const bool inc = MODE == M_RATES && io.in_conc;       // the slot holds C_k, not Y_k
                const double Yk[2] = {inc ? Yv.x * wk / rh[0] : Yv.x, inc ? Yv.y * wk / rh[1] : Yv.y};
                const double tmid = __ldg(tb.sp_tmid + k);
                double ck[2], cp[2], Bk[2], dB[2], hW[2];
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const double* c = tb.sp_nasa + (k * 2 + (Tv[g] <= tmid ? 0 : 1)) * 16;
                    const double t = Tv[g];
                    ck[g] = inc ? (g ? Yv.y : Yv.x) : rh[g] * Yk[g] * iw;
                    // constant volume (dydt only): cv and u in place of cp and h (rs:1876-2019), c[11] = a0 - 1
                    const double a0 = (!jac_like(MODE) && io.conv) ? c[11] : c[0];
                    cp[g] = ruw * (a0 + t * (c[1] + t * (c[2] + t * (c[3] + c[4] * t))));
                    const double hh = c[6] + t * (c[7] + t * (c[8] + c[9] * t));
                    hW[g] = ruw * (c[5] + t * (a0 + t * hh)) * wk;
                    const double dcp = ruw * (c[1] + t * (2.0 * c[2] + t * (3.0 * c[3] + 4.0 * c[4] * t)));
                    cpavg[g] += Yk[g] * cp[g];
                    wdcp[g] += Yk[g] * dcp;
                    dB[g] = (c[11] + c[5] * rT[g]) * rT[g] + hh;
                    Bk[g] = c[10] + c[11] * lT[g] + t * (c[6] + t * (c[12] + t * (c[13] + c[14] * t))) - c[5] * rT[g];
                }
                STS(E_C * RB, a, V{ck[0], ck[1]});
                if (MODE == M_RATES && io.conc) put2(io.conc, io, nsp, out, k, V{ck[0], ck[1]});
                STS(O_B * RB, o, V{Bk[0], Bk[1]});
                STS(E_DB * RB, a, V{dB[0], dB[1]});
                STS(O_HW * RB, o, V{hW[0], hW[1]});
                STS(O_CP * RB, o, V{cp[0], cp[1]});
            }
            const V ca = sub_sum<GS>(V{cpavg[0], cpavg[1]}), wd = sub_sum<GS>(V{wdcp[0], wdcp[1]});
            if (sub == 0) {
                STS(D_CPAVG * RB, aPA + warp * NPART * RB, ca);
                STS(D_WDCP * RB, aPA + warp * NPART * RB, wd);
            }
        }
        // the first reaction record of phase B is requested before the barrier (tables come from L2)
        const int b_r0 = __ldg(pl.b_off + warp), b_r1 = __ldg(pl.b_off + warp + 1);
        const int b_rpm = b_r0 + __ldg(pl.b_npm + warp);
        const int b_item0 = __ldg(pl.b_item + b_r0 * NSUB + sub);
        int4 b_q0, b_q1, b_q2, b_q3;
        {
            const int4* rp = pl.rx + (b_item0 >= 0 ? b_item0 : (b_r0 < b_rpm ? tb.first_pm : 0)) * 4;
            b_q0 = __ldg(rp); b_q1 = __ldg(rp + 1); b_q2 = __ldg(rp + 2); b_q3 = __ldg(rp + 3);
        }
        __syncthreads();
        PJ_TICK(0)

        // ------------------------------------------------------------ phase B: reactions
        if (!PJ_SKIP(2)) {
            const V T = LDS(Q_T * RB, aSC), logT = LDS(Q_LOGT * RB, aSC), iT = LDS(Q_IT * RB, aSC);
            const unsigned nsp_f = sp_even<GS>(0u, (unsigned)nsp) / 16;
            int item = b_item0;
            int4 q0 = b_q0, q1 = b_q1, q2 = b_q2, q3 = b_q3;       // first record: requested before the barrier
            for (int r = b_r0; r < b_r1; ++r) {
                const bool valid = item >= 0;
                const int p = valid ? item : (r < b_rpm ? tb.first_pm : 0);
                if (r > b_r0) {
                    const int4* rp = pl.rx + p * 4;
                    q0 = __ldg(rp); q1 = __ldg(rp + 1); q2 = __ldg(rp + 2); q3 = __ldg(rp + 3);
                }
                const int nxt = __ldg(pl.b_item + (r + 1) * NSUB + sub);    // the table ends with a null round
                if (r < b_rpm) {
                    reaction<GS, true, MODE, WSG>(mem, tb, pl, io, out, aSP, aRX, aRAW, aSC, p, valid, true, q0, q1, q2, q3, T, logT, iT);
                } else {
                    const bool has3 = ((q2.z & 0xFFFFu) != nsp_f) || (((unsigned)q2.w >> 16) != nsp_f);
                    const bool three = __any_sync(0xffffffffu, has3);
                    // rounds holding a PLOG / Chebyshev / negative-A reaction take the build of the routine that knows them
                    if (__any_sync(0xffffffffu, (q2.x & (F_PLOG | F_CHEB | F_NEGA)) != 0))
                        reaction_plain<GS, MODE, true, WSG>(mem, tb, pl, io, out, aSP, aRX, aRAW, aSC, aSD + buf * RB, p, valid, three, q0, q1, q2, q3, T, logT, iT);
                    else
                        reaction_plain<GS, MODE, false, WSG>(mem, tb, pl, io, out, aSP, aRX, aRAW, aSC, aSD + buf * RB, p, valid, three, q0, q1, q2, q3, T, logT, iT);
                }
                item = nxt;
            }
        }
        PJ_TICK(1)
        __syncthreads();
        PJ_TICK(2)

        // ------------------------------------------------------------ phase C: species sums
        if (!PJ_SKIP(4)) {
            const int i0 = __ldg(pl.c_off + warp), i1 = __ldg(pl.c_off + warp + 1);
            const V mwr = LDS(Q_MWR * RB, aSC);
            V pH1 = zero, pHA = zero, pHB = zero, pHT = zero, pSCP = zero;
            // round header {first unit, #(+1) units, #(-1) units}; unit 0 = {species row offset or
            // NONE, 1 if this sub-group stores}; pl.coop sub-groups share one species
            int4 h = __ldg(pl.c_item + i0);
            for (int it = i0; it < i1; ++it) {
                const uint2* cp = pl.c_str + (long long)h.x * NSUB + sub;
                const int np_ = h.y, n = h.y + h.z;
                h = __ldg(pl.c_item + it + 1);
                const uint2 hd = __ldg(cp);
                V aN = zero, aT = zero, a1 = zero, a2 = zero;
                auto unit = [&](uint2 c, double sg) {
                    const unsigned x = aRX + c.x, y = aRX + c.y;
                    aN = vfma(sg, vadd(LDS(RX_NET * RB, x), LDS(RX_NET * RB, y)), aN);
                    if (!jac_like(MODE)) return;                 // only the net rates are summed
                    aT = vfma(sg, vadd(LDS(RX_TT * RB, x), LDS(RX_TT * RB, y)), aT);
                    a1 = vfma(sg, vadd(LDS(RX_X1 * RB, x), LDS(RX_X1 * RB, y)), a1);
                    a2 = vfma(sg, vadd(LDS(RX_X2 * RB, x), LDS(RX_X2 * RB, y)), a2);
                };
                int i = 0;
                // two units per iteration, the next two requested before these are summed (the
                // stream is padded: reading past the round is harmless)
                uint2 c = __ldg(cp + NSUB), d = __ldg(cp + 2 * NSUB);
                for (; i + 2 <= n; i += 2) {
                    const uint2 c2 = __ldg(cp + (i + 3) * NSUB), d2 = __ldg(cp + (i + 4) * NSUB);
                    unit(c, i < np_ ? 1.0 : -1.0);
                    unit(d, i + 1 < np_ ? 1.0 : -1.0);
                    c = c2; d = d2;
                }
                if (i < n) unit(c, i < np_ ? 1.0 : -1.0);
                for (int o = NPR; o < NPR * pl.coop; o <<= 1) {
                    aN = V{aN.x + __shfl_xor_sync(0xffffffffu, aN.x, o), aN.y + __shfl_xor_sync(0xffffffffu, aN.y, o)};
                    if (!jac_like(MODE)) continue;
                    aT = V{aT.x + __shfl_xor_sync(0xffffffffu, aT.x, o), aT.y + __shfl_xor_sync(0xffffffffu, aT.y, o)};
                    a1 = V{a1.x + __shfl_xor_sync(0xffffffffu, a1.x, o), a1.y + __shfl_xor_sync(0xffffffffu, a1.y, o)};
                    a2 = V{a2.x + __shfl_xor_sync(0xffffffffu, a2.x, o), a2.y + __shfl_xor_sync(0xffffffffu, a2.y, o)};
                }
                if (hd.y && !jac_like(MODE)) {
                    // dydt / rates: omega_k, dY_k/dt = omega_k W_k / rho, share of sum_k h_k W_k omega_k
                    const int k = hd.x / SPB;
                    const double wk = __ldg(tb.sp_w + k);
                    pH1 = vfma(LDS(O_HW * RB, (aSP + hd.x) ^ RB), aN, pH1);
                    const V ri = LDS(Q_RHOINV * RB, aSC);
                    const V dyk = vmul(wk, vmul(aN, ri));
                    if (MODE == M_RATES) {
                        if (io.sr) put2(io.sr, io, nsp, out, k, aN);
                        if (io.dy && k < last) put2(io.dy, io, nsp, out, k + 1, dyk);
                    } else if (k < last) {
                        if (ok0) io.dy[s0 * io.dy_ss + (long long)(k + 1) * io.dy_sv] = dyk.x;
                        if (ok1) io.dy[(s0 + 1) * io.dy_ss + (long long)(k + 1) * io.dy_sv] = dyk.y;
                    }
                }
                if (hd.y && jac_like(MODE)) {
                    const unsigned a = aSP + hd.x, o = a ^ RB;      // hd.x: even-slot base of species k
                    const double wk = __ldg(tb.sp_w + hd.x / SPB);
                    const V comp = vmul(aN, mwr);
                    a1 = vadd(a1, comp);
                    a2 = vsub(a2, comp);
                    const V hW = LDS(O_HW * RB, o), cp_ = LDS(O_CP * RB, o);
                    pH1 = vfma(hW, aN, pH1);
                    pHA = vfma(hW, a1, pHA);
                    pHB = vfma(hW, a2, pHB);
                    pHT = vfma(hW, aT, pHT);
                    pSCP = vfma(vmul(wk, cp_), aN, pSCP);
                    STS(E_WA * RB, a, vmul(wk, a1));
                    STS(O_WB * RB, o, vmul(wk, a2));
                    STS(E_WT * RB, a, vmul(wk, aT));
                }
            }
            // the warp's share of the energy-equation dot products
            pH1 = sub_sum<GS>(pH1); pHA = sub_sum<GS>(pHA); pHB = sub_sum<GS>(pHB);
            pHT = sub_sum<GS>(pHT); pSCP = sub_sum<GS>(pSCP);
            if (sub == 0) {
                const unsigned a = aPA + warp * NPART * RB;
                STS(D_H1 * RB, a, pH1); STS(D_HA * RB, a, pHA); STS(D_HB * RB, a, pHB);
                STS(D_HT * RB, a, pHT); STS(D_SCP * RB, a, pSCP);
            }
        }
        __syncthreads();
        PJ_TICK(3)

        // ------------------------------------------------------------ dydt / rates: energy equation
        if (!jac_like(MODE)) {
            if (warp == 0) {
                V H1 = zero, cpavg = zero;
                for (int w = sub; w < nw; w += NSUB) {
                    H1 = vadd(H1, LDS(D_H1 * RB, aPA + w * NPART * RB));
                    cpavg = vadd(cpavg, LDS(D_CPAVG * RB, aPA + w * NPART * RB));
                }
                H1 = sub_sum<GS>(H1);
                cpavg = sub_sum<GS>(cpavg);
                if (sub == 0 && io.dy) {
                    const V rho = LDS(Q_RHO * RB, aSC);
                    const V d0{-1.0 / (rho.x * cpavg.x) * H1.x, -1.0 / (rho.y * cpavg.y) * H1.y};
                    if (MODE == M_RATES) put2(io.dy, io, nsp, out, 0, d0);
                    else {
                        if (ok0) io.dy[s0 * io.dy_ss] = d0.x;
                        if (ok1) io.dy[(s0 + 1) * io.dy_ss] = d0.y;
                    }
                }
                if (grp + gridDim.x < ngroups) phase_a0(grp + gridDim.x, buf ^ 1);
            }
            __syncthreads();
            continue;
        }

        // ------------------------------------------------------------ phase DE
        if (warp == 0 && !PJ_SKIP(64)) {
            // energy-equation scalars from the per-warp partial sums; the result of quantity q
            // replaces warp 0's own partial
            for (int q = sub; q < NPART; q += NSUB) {
                V a = zero;
                for (int w = 0; w < nw; ++w) a = vadd(a, LDS(0, aPA + (w * NPART + q) * RB));
                STS(0, aPA + q * RB, a);
            }
            __syncwarp();
            if (sub == 0) {
                const V H1 = LDS(D_H1 * RB, aPA), HA = LDS(D_HA * RB, aPA), HB = LDS(D_HB * RB, aPA);
                const V HT = LDS(D_HT * RB, aPA), SCP = LDS(D_SCP * RB, aPA);
                const V cpavg = LDS(D_CPAVG * RB, aPA), wdcp = LDS(D_WDCP * RB, aPA);
                const V rho = LDS(Q_RHO * RB, aSC), cpl = LDS(O_CP * RB, sp_even<GS>(aSP, (unsigned)last) ^ RB);
                const V nwt{-1.0 / cpavg.x, -1.0 / cpavg.y};
                STS(S_NWT * RB, aSD, nwt);
                STS(S_A0 * RB, aSD, vmul(nwt, HA));
                STS(S_B0 * RB, aSD, vmul(nwt, HB));
                STS(S_XT * RB, aSD, V{H1.x / (rho.x * cpavg.x * cpavg.x), H1.y / (rho.y * cpavg.y * cpavg.y)});
                STS(S_CPL * RB, aSD, cpl);
                // jac[0] (cj:1853-1905)
                store(0u, V{-(-wdcp.x / cpavg.x * H1.x + SCP.x + HT.x * rho.x) / (rho.x * cpavg.x),
                            -(-wdcp.y / cpavg.y * H1.y + SCP.y + HT.y * rho.y) / (rho.y * cpavg.y)}, true);
            }
            __syncwarp();          // the other sub-groups of warp 0 read these scalars in class T
            if (pl.t_sync > 32) {
                __threadfence_block();
                asm volatile("bar.arrive 1, %0;" ::"r"(pl.t_sync) : "memory");
            }
            // the next group's phase A0 (its inputs come from HBM: latency hidden behind DE)
            if (grp + gridDim.x < ngroups) phase_a0(grp + gridDim.x, buf ^ 1);
        }
        if (!PJ_SKIP(8)) {
            // class S: elements with a sparse part, two steps per iteration, next pair in flight
            const int st0 = __ldg(pl.s_off + warp), st1 = __ldg(pl.s_off + warp + 1);
            const uint4* sp = pl.s_str + (long long)st0 * 2 * NSUB + sub;
            const uint2* ov = pl.o_str + (long long)__ldg(pl.o_off + warp) * NSUB + sub;
            uint2 ob0 = __ldg(ov), ob1 = __ldg(ov + NSUB), ob2 = __ldg(ov + 2 * NSUB), ob3 = __ldg(ov + 3 * NSUB);
            // one step: the bundle (A, B) is consumed while `nA`, `nB` of a later step are fetched
            auto step = [&](const uint4& A, const uint4& B) {
                const unsigned L = A.x >> 22;
                unsigned e = A.x & NULL_E;
                const unsigned x = aSP + (A.y & 0xFFFFFu);
                V cf{0.0, 0.0};
                if (MODE == M_FACT) { if (e != NULL_E) e = (unsigned)__ldg(pl.fac_map + e); }
                else cf = LDS(0, aCF + (A.y >> 20) * 16);
                // signed entries: byte offset of a raw row | 1 for weight -1; two entries per unit
                auto sgn = [](unsigned c) { return __hiloint2double((int)(0x3FF00000u | (c << 31)), 0); };
                V p = vmul(sgn(B.x), LDS(0, aRAW + (B.x & ~1u))), m = vmul(sgn(B.y), LDS(0, aRAW + (B.y & ~1u)));
                V v{0.0, 0.0};                                               // M_FACT: the sparse part alone
                if (MODE != M_FACT) v = vfma(cf.y, LDS(0, x ^ RB), vmul(cf.x, LDS(0, x)));     // W_k a_k at x, W_k b_k at x ^ RB
                if (L > 1) {
                    p = vfma(sgn(B.z), LDS(0, aRAW + (B.z & ~1u)), p);
                    m = vfma(sgn(B.w), LDS(0, aRAW + (B.w & ~1u)), m);
                    // units 3..L come from the overflow stream in batches of four (padded)
#pragma unroll 1
                    for (unsigned i = 2; i < L; i += 4) {
                        // the batch was requested when the previous one was taken (the stream is
                        // consumed in order and padded at its end)
                        const uint2 c0 = ob0, c1 = ob1, c2 = ob2, c3 = ob3;
                        ov += 4 * NSUB;
                        ob0 = __ldg(ov); ob1 = __ldg(ov + NSUB); ob2 = __ldg(ov + 2 * NSUB); ob3 = __ldg(ov + 3 * NSUB);
                        p = vfma(sgn(c0.x), LDS(0, aRAW + (c0.x & ~1u)), p);
                        m = vfma(sgn(c0.y), LDS(0, aRAW + (c0.y & ~1u)), m);
                        p = vfma(sgn(c1.x), LDS(0, aRAW + (c1.x & ~1u)), p);
                        m = vfma(sgn(c1.y), LDS(0, aRAW + (c1.y & ~1u)), m);
                        p = vfma(sgn(c2.x), LDS(0, aRAW + (c2.x & ~1u)), p);
                        m = vfma(sgn(c2.y), LDS(0, aRAW + (c2.y & ~1u)), m);
                        p = vfma(sgn(c3.x), LDS(0, aRAW + (c3.x & ~1u)), p);
                        m = vfma(sgn(c3.y), LDS(0, aRAW + (c3.y & ~1u)), m);
                    }
                }
                v = vfma(__hiloint2double((int)A.w, (int)A.z), vadd(p, m), v);
                store(e, v, e != NULL_E);
            };
            // four bundles rotate through registers; every bundle is fetched two steps ahead
            uint4 A0 = __ldg(sp), B0 = __ldg(sp + NSUB), A1 = __ldg(sp + 2 * NSUB), B1 = __ldg(sp + 3 * NSUB);
            for (int st = st0; st < st1; st += 4) {
                const uint4 A2 = __ldg(sp + 4 * NSUB), B2 = __ldg(sp + 5 * NSUB);
                const uint4 A3 = __ldg(sp + 6 * NSUB), B3 = __ldg(sp + 7 * NSUB);
                step(A0, B0);
                step(A1, B1);
                if (st + 2 >= st1) break;
                sp += 8 * NSUB;
                A0 = __ldg(sp); B0 = __ldg(sp + NSUB); A1 = __ldg(sp + 2 * NSUB); B1 = __ldg(sp + 3 * NSUB);
                step(A2, B2);
                step(A3, B3);
            }
        }
        PJ_TICK(4)
        if (!PJ_SKIP(16)) {
            // class D: dense-only elements by row.  A sub-group keeps W_k a_k, W_k b_k of its row
            // in registers and walks the row's dense-only columns, four in flight.
            const int i0 = __ldg(pl.d_off + warp), i1 = __ldg(pl.d_off + warp + 1);
            for (int it = i0; it < i1; ++it) {
                const int2 di = __ldg(pl.d_item + it);                 // first unit, #columns (multiple of 4)
                const uint2* dp = pl.d_str + (long long)di.x * NSUB + sub;
                const uint2 hd = __ldg(dp);                            // {even-slot base or NONE, element of column 0}
                const bool on = hd.x != 0xFFFFFFFFu;
                const unsigned x = aSP + (on ? hd.x : 0u);
                const V wa = LDS(E_WA * RB, x), wb = LDS(O_WB * RB, x ^ RB);
                if (MODE == M_FACT) {
                    // the row's three factors, once (the item that owns the temperature column)
                    const bool own = on && hd.y != NULL_E;
                    const unsigned k_ = own ? hd.y - 1u : 0u;                  // hd.y: element (k + 1, 0)
                    store((unsigned)nsp + k_, LDS(E_WT * RB, x), own);
                    store((unsigned)(nsp + last) + k_, wa, own);
                    store((unsigned)(nsp + 2 * last) + k_, wb, own);
                    continue;
                }
                store(hd.y, LDS(E_WT * RB, x), on && hd.y != NULL_E);  // temperature column: W_k * T-term
                // units are fetched LA batches of four ahead (the tables come from L2: with all of
                // shared memory in use there is no L1 to speak of)
                constexpr int LA = 1;
                uint2 nx[4 * LA];
#pragma unroll
                for (int j = 0; j < 4 * LA; ++j) nx[j] = __ldg(dp + (1 + j) * NSUB);
                for (int c = 0; c < di.y; c += 4) {
                    uint2 r[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) r[j] = nx[j];
#pragma unroll
                    for (int j = 0; j < 4 * (LA - 1); ++j) nx[j] = nx[j + 4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) nx[4 * (LA - 1) + j] = __ldg(dp + (1 + 4 * LA + c + j) * NSUB);
                    V cf[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) cf[j] = LDS(0, aCF + r[j].y * 16);
#pragma unroll
                    for (int j = 0; j < 4; ++j) store(r[j].x, vfma(cf[j].y, wb, vmul(cf[j].x, wa)), r[j].x != NULL_E);
                }
            }
        }
        PJ_TICK(5)
        if (!PJ_SKIP(32)) {
            // class T, the energy-equation row (cj:3095-3254): per column an enthalpy-weighted
            // gather; pl.tcoop sub-groups share a column, units come four at a time
            const int i0 = __ldg(pl.t_off + warp), i1 = __ldg(pl.t_off + warp + 1);
            if (i0 < i1) {
                int2 ti = __ldg(pl.t_item + i0);
                if (warp != 0) asm volatile("bar.sync 1, %0;" ::"r"(pl.t_sync) : "memory");
                const V nwt = LDS(S_NWT * RB, aSD), A0 = LDS(S_A0 * RB, aSD), B0 = LDS(S_B0 * RB, aSD);
                const V XT = LDS(S_XT * RB, aSD), cpl = LDS(S_CPL * RB, aSD);
                for (int it = i0; it < i1; ++it) {
                    const uint2* up = pl.t_str + (long long)ti.x * NSUB + sub;
                    const int n = ti.y;
                    ti = __ldg(pl.t_item + it + 1);
                    const uint2 hd = __ldg(up);                        // {col | store << 16, offset of cp_j}
                    V acc0 = zero, acc1 = zero;
                    uint2 nx[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) nx[j] = __ldg(up + (1 + j) * NSUB);
                    for (int c = 0; c < n; c += 4) {
                        uint2 r[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) { r[j] = nx[j]; nx[j] = __ldg(up + (5 + c + j) * NSUB); }
                        acc0 = vfma(LDS(RX_DH * RB, aRX + r[0].y), LDS(0, aRAW + r[0].x), acc0);
                        acc1 = vfma(LDS(RX_DH * RB, aRX + r[1].y), LDS(0, aRAW + r[1].x), acc1);
                        acc0 = vfma(LDS(RX_DH * RB, aRX + r[2].y), LDS(0, aRAW + r[2].x), acc0);
                        acc1 = vfma(LDS(RX_DH * RB, aRX + r[3].y), LDS(0, aRAW + r[3].x), acc1);
                    }
                    V E0 = vadd(acc0, acc1);
                    for (int o = NPR; o < NPR * pl.tcoop; o <<= 1)
                        E0 = V{E0.x + __shfl_xor_sync(0xffffffffu, E0.x, o), E0.y + __shfl_xor_sync(0xffffffffu, E0.y, o)};
                    const unsigned col = hd.x & 0xFFFFu;
                    const V cf = LDS(0, aCF + col * 16);
                    const V cpj = LDS(0, aSP + hd.y);
                    V v = vmul(cf.x, vfma(nwt, E0, A0));
                    v = vfma(cf.y, B0, v);
                    v = vfma(XT, vsub(cpj, cpl), v);
                    store(MODE == M_FACT ? col : col * (unsigned)nsp, v, (hd.x >> 16) != 0u);
                }
            }
        }
        PJ_TICK(6)
        __syncthreads();
        PJ_TICK(7)
    }
#undef PJ_TICK
#ifdef PJ_PHASE_CLOCKS
    if (io.dbg_clk && blockIdx.x == 0 && lane == 0)
        for (int i = 0; i < 8; ++i) io.dbg_clk[warp * 8 + i] = clk[i];
#endif
}

}  // namespace pj5
