// eval_jacob for sm_100a, second generation: the kernel of eval.cuh (same algebra, same mapping of
// state pairs to lanes and of table items to sub-groups) with every table word delivered through a
// per-warp *record stream* (pyjac_b200/plan6.py).
//
// Replaces the reference's generated eval_jacob (pyjac/core/create_jacobian.py:2189-3298) and the
// rate routines it calls (rate_subs.py:254-876, 879-1294, 1297-1542, 1626-1706, 1806-2086).
//
// What differs from pj5::k_eval<.., M_JAC>:
//   * A warp's table items (reaction records, species-sum lists, Jacobian element records) form one
//     stream in global memory, consumed strictly in order and identical for every group of states.
//     Lane 0 copies it ahead of use into a ring of NSLOT chunks in shared memory with bulk-asynchronous
//     copies (cp.async.bulk -> SASS UBLKCP) that complete on an mbarrier; a record is NSUB x 16 bytes
//     and one conflict-free ld.shared.v4 hands every sub-group its word.  No table word of phases
//     B / C / DE goes through the load/store pipe from global memory any more.
//   * Jacobian elements are produced by rows: a sub-group keeps W_k a_k, W_k b_k and W_k of its row in
//     registers for a whole segment, an element record holds up to six signed 16-bit raw-row indices
//     and the step's warp-uniform entry count, so dense-only elements issue no gather at all and the
//     padding of a step is bounded by its longest list.
//   * Working set per state: species rows C B dB hW WT cp (W_k a_k / W_k b_k overwrite dB / B after
//     phase B), reaction rows net tT X1 dH, X2 = -X1 + a correction row for the few reactions where
//     that does not hold.  The slot pairs of odd species / reactions are swapped (E ^ RB) so that
//     equal slots of different items fall on both halves of a 128-byte bank line.
#pragma once
#include <type_traits>

#include "eval.cuh"

namespace pj6 {

using namespace pj;
using pj5::Mem;
using pj5::V;
using pj5::exp_n;
using pj5::sub_sum;
using pj5::vadd;
using pj5::vfma;
using pj5::vmul;
using pj5::vsub;
using pj5::zero_v;
using pj5::Out;

// p6_cfg (plan6.py), then the device tables
struct Plan6 {
    int gs, nt, nw, nsub, oSP, oRX, oXC, oRAW, oET, oSC, oPA, oCF, ring, mbar, bytes, t_sync, coop, tcoop, p_c0, ncorr,
        chb, nslot, chr, pad;
    const uint4* str;               // record streams, chunks of chb bytes
    const int* hdr;                 // per warp {first chunk, #chunks, #pm rounds, #plain rounds, #species rounds,
                                    //           #energy-row rounds, #segments, #energy-row records}
    const int* eff_off;
    const int4* eff;
    const double2* colfac;
};

enum : int { Q_T = 0, Q_LOGT, Q_IT, Q_RHO, Q_RHOINV, Q_LNP, Q_MWR, Q_M };
enum : int { S_NWT = 0, S_A0, S_B0, S_XT, S_CPL, S_P = 5 };
enum : int { D_H1 = 0, D_HA, D_HB, D_HT, D_SCP, D_CPAVG, D_WDCP, NPART = 7 };
// species rows: even slots at E + {0, 2, 4} rows, odd slots at (E ^ RB) + {0, 2, 4} rows
enum : int { SP_SLOTS = 6, E_C = 0, E_DB = 2, E_WA = 2, E_WT = 4, O_B = 0, O_WB = 0, O_HW = 2, O_CP = 4, E_Y = E_C };
// reaction rows: even slots (net, X1) at E + {0, 2}, odd slots (tT, dH) at (E ^ RB) + {0, 2}
enum : int { RX_SLOTS = 4, E_NET = 0, E_X1 = 2, O_TT = 0, O_DH = 2 };
#ifndef PJ_NSLOT
#define PJ_NSLOT 3
#endif
enum : int { CHB = 512, NSLOT = PJ_NSLOT };
enum : unsigned { F_NULL = 1u << 28, F_CORR = 1u << 29, D_CIN = 1u << 28, D_COUT = 1u << 29, D_VALID = 1u << 30, D_HDR = 1u << 31,
                  CLS_CARRY = 7u, NONE32 = 0xFFFFFFFFu };

#define LDS(OFF, ...) mem.template ld<(OFF)>(__VA_ARGS__)
#define STS(OFF, ...) mem.template st<(OFF)>(__VA_ARGS__)
#define STS_IF(OFF, ...) mem.template st_if<(OFF)>(__VA_ARGS__)

template <int GS>
__device__ __forceinline__ unsigned sp_even(unsigned aSP, unsigned k) { return aSP + k * (SP_SLOTS * GS * 8) + (k & 1u) * (GS * 8); }

// ---- mbarrier / bulk copy (PTX; SASS: SYNCS.*, UBLKCP)
__device__ __forceinline__ void mbar_init(unsigned a, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned a, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned a, unsigned parity)
{
    // try_wait suspends the thread up to a hardware time limit; a copy that never lands (a broken
    // stream table) traps instead of hanging the device
    unsigned done, spins = 0;
    do {
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, P1;\n\t}" : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 22)) __trap();
    } while (!done);
}
__device__ __forceinline__ uint4 lds128(unsigned a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}

// The warp's record stream.  All lanes call get() together; lane 0 requests the chunks.
template <int NSUB>
struct Stream {
    static constexpr unsigned RECB = NSUB * 16, CHR = CHB / RECB;
    unsigned ring0;      // shared address of the warp's ring
    unsigned ring;       // ... + this sub-group's word
    unsigned mbar;       // shared address of the warp's NSLOT mbarriers
    const char* src;     // the warp's stream
    unsigned nch;        // chunks per group
    unsigned left;       // chunks not requested yet (over all groups of this block)
    unsigned nxt;        // chunk of the stream the next request fetches
    unsigned slot, par, rec;
    unsigned pend;       // slot + 1 whose re-arming is due, 0 = none
    unsigned guard;      // shared address of a word that always holds zero
    unsigned chunks;     // chunks consumed so far (development checks)
    bool lane0;

    __device__ __forceinline__ void request(unsigned s)
    {
        if (lane0) {
            mbar_expect_tx(mbar + s * 8, CHB);
            bulk_g2s(ring0 + s * CHB, src + (size_t)nxt * CHB, CHB, mbar + s * 8);
        }
        nxt = nxt + 1 == nch ? 0 : nxt + 1;
        --left;
    }
    __device__ __forceinline__ void start(unsigned ring_, unsigned word, unsigned mbar_, unsigned guard_, const char* src_, unsigned nch_, unsigned total, bool lane0_)
    {
        guard = guard_; ring0 = ring_; ring = ring_ + word; mbar = mbar_; src = src_; nch = nch_; left = total; nxt = 0; slot = 0; par = 0; rec = 0; pend = 0; chunks = 0; lane0 = lane0_;
        if (lane0) {
            for (unsigned s = 0; s < NSLOT; ++s) mbar_init(mbar + s * 8, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
        for (unsigned s = 0; s < NSLOT && left; ++s) request(s);
    }
    // The chunk in `slot` has been handed out.  Its slot is re-armed (the next bulk copy into it is
    // issued) not here but by flush(), which the next wait_chunk() or align() calls: by then the
    // records of the chunk have been *used*, so every ld.shared of the slot has returned and the copy
    // (async proxy) cannot overtake a read (generic proxy) still in flight.  Measured on a B200:
    // re-arming right after the last ld.shared corrupts about one record per hundred groups; a
    // fence.proxy.async in between costs a MEMBAR per chunk.
    __device__ __forceinline__ void release()
    {
        rec = 0;
        ++chunks;
        par ^= 1u << slot;
        pend = slot + 1;
        slot = slot + 1 == NSLOT ? 0 : slot + 1;
    }
    __device__ __forceinline__ void flush()
    {
        if (pend) {
            __syncwarp();
            // Shared-memory loads of a warp complete in order: once this load of a word that is always
            // zero has returned (the branch needs its value), every earlier ld.shared of the slot has
            // returned too, whether or not its value has been used yet.
            unsigned g;
            asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(g) : "r"(guard) : "memory");
            if (g != 0u) __trap();
            if (left) request(pend - 1);
            pend = 0;
        }
    }
    __device__ __forceinline__ void wait_chunk()
    {
        flush();
        mbar_wait(mbar + slot * 8, (par >> slot) & 1u);
    }
    // address of record i of the current chunk (this sub-group's word)
    __device__ __forceinline__ unsigned rec_addr(unsigned i) const { return ring + slot * CHB + i * RECB; }
    __device__ __forceinline__ uint4 get()
    {
        if (rec == 0) wait_chunk();
        const uint4 v = lds128(rec_addr(rec));
        if (++rec == CHR) release();
        return v;
    }
    // skip to the next chunk boundary (the phases that take whole chunks start on one; the stream of
    // a group ends with its last chunk)
    __device__ __forceinline__ void align()
    {
        if (rec != 0) release();
        flush();
    }
};

#ifdef PJ_DEV
// development: a record that fails a sanity check is written to io.dbg_clk = {count, 16 words per event}
__device__ __noinline__ void dbg_event(long long* buf, int phase, long long grp, unsigned chunks, unsigned rec, unsigned slot,
                                       uint4 v, unsigned extra)
{
    if (!buf) return;
    const unsigned long long i = atomicAdd((unsigned long long*)buf, 1ull);
    if (i >= 64) return;
    long long* o = buf + 1 + i * 16;
    o[0] = blockIdx.x; o[1] = threadIdx.x; o[2] = phase; o[3] = grp; o[4] = chunks; o[5] = rec; o[6] = slot;
    o[7] = v.x; o[8] = v.y; o[9] = v.z; o[10] = v.w; o[11] = extra;
    __threadfence_system();
}
#define PJ_CHECK(COND, PHASE, V, EXTRA) if (!(COND)) dbg_event(io.dbg_clk, PHASE, grp, rd.chunks, rd.rec, rd.slot, V, EXTRA)
#else
#define PJ_CHECK(COND, PHASE, V, EXTRA)
#endif

__device__ __forceinline__ double dbl(unsigned lo, unsigned hi) { return __hiloint2double((int)hi, (int)lo); }
// +1.0 / -1.0 from bit 15 of a 16-bit entry
__device__ __forceinline__ double sgn15(unsigned x) { return __hiloint2double((int)(0x3FF00000u | ((x & 0x8000u) << 16)), 0); }

// Phase B, one pressure-modified reaction and the two states of the lane (pj5::reaction, Jacobian only)
template <int GS>
__device__ __forceinline__ void reaction(const Mem<false>& mem, const Tables& tb, const Plan6& pl,
                                         unsigned aSP, unsigned aRX, unsigned aXC, unsigned aRAW, unsigned aSC, int p, bool valid,
                                         const uint4 q0, const uint4 q1, const uint4 q2, const uint4 q3,
                                         const V T, const V logT, const V iT, const V rho_inv, const V nmwr)
{
    constexpr int RB = GS * 8;
    const int nsp = tb.nsp, last = tb.nsp - 1;
    const double lnA = dbl(q0.x, q0.y), bexp = dbl(q0.z, q0.w);
    const double Ta = dbl(q1.x, q1.y), lnKc = dbl(q1.z, q1.w);
    const int fl = (int)q2.x;
    const unsigned s0 = q2.y & 0xFFFFu, s1 = q2.y >> 16, s2 = q2.z & 0xFFFFu;
    const unsigned s3 = q2.z >> 16, s4 = q2.w & 0xFFFFu, s5 = q2.w >> 16;
    const unsigned a0 = aSP + s0 * 16, a1 = aSP + s1 * 16, a2 = aSP + s2 * 16;
    const unsigned a3 = aSP + s3 * 16, a4 = aSP + s4 * 16, a5 = aSP + s5 * 16;
    const unsigned nsp_f = (sp_even<GS>(0u, (unsigned)nsp)) / 16, last_f = (sp_even<GS>(0u, (unsigned)last)) / 16;
    const bool isrev = fl & F_REV;

    const V c0 = LDS(E_C * RB, a0), c1 = LDS(E_C * RB, a1), c2 = LDS(E_C * RB, a2);
    const V c3 = LDS(E_C * RB, a3), c4 = LDS(E_C * RB, a4), c5 = LDS(E_C * RB, a5);
    V sB = vsub(vadd(LDS(O_B * RB, a3 ^ RB), LDS(O_B * RB, a4 ^ RB)), vadd(LDS(O_B * RB, a0 ^ RB), LDS(O_B * RB, a1 ^ RB)));
    V sdB = vsub(vadd(LDS(E_DB * RB, a3), LDS(E_DB * RB, a4)), vadd(LDS(E_DB * RB, a0), LDS(E_DB * RB, a1)));
    V dH = vsub(vadd(LDS(O_HW * RB, a3 ^ RB), LDS(O_HW * RB, a4 ^ RB)), vadd(LDS(O_HW * RB, a0 ^ RB), LDS(O_HW * RB, a1 ^ RB)));
    sB = vadd(sB, vsub(LDS(O_B * RB, a5 ^ RB), LDS(O_B * RB, a2 ^ RB)));
    sdB = vadd(sdB, vsub(LDS(E_DB * RB, a5), LDS(E_DB * RB, a2)));
    dH = vadd(dH, vsub(LDS(O_HW * RB, a5 ^ RB), LDS(O_HW * RB, a2 ^ RB)));
    const V lnkf = vfma(bexp, logT, V{fma(-Ta, iT.x, lnA), fma(-Ta, iT.y, lnA)});
    const V lnkr{lnkf.x - sB.x - lnKc, lnkf.y - sB.y - lnKc};
    V kf, kr;

    V PM_{1.0, 1.0}, gg{1.0, 1.0}, Xd{0.0, 0.0}, e1Fi{0.0, 0.0};
    const int mi = p - tb.first_pm;
    const double* par = tb.pm_par + mi * NPAR;
    bool rates_done = false;
    {
        V thd = LDS(Q_M * RB, aSC);
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

    if (!rates_done) {
        const double x2[4] = {lnkf.x, lnkf.y, lnkr.x, lnkr.y};
        double y2[4];
        exp_n<4>(x2, y2);
        kf = V{y2[0], y2[1]};
        kr = V{y2[2], y2[3]};
    }
    if (!isrev) kr = V{0.0, 0.0};
    if (fl & F_NEGA) { kf = V{-kf.x, -kf.y}; kr = V{-kr.x, -kr.y}; }   // A < 0 (rs:108-141)
    const V f = vmul(vmul(kf, vmul(c0, c1)), c2), r = vmul(vmul(kr, vmul(c3, c4)), c5);
    const V net = vsub(f, r);
    V pmt{0.0, 0.0};
    if (fl & F_PMT) pmt = vmul(gg, net);

    const double nre = (double)((fl >> NRE_SHIFT) & 15), npr = (double)((fl >> NPR_SHIFT) & 15);
    const double extra = (fl & F_EFFN1) ? 1.0 : 0.0;
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
            const double elem = nv[g] * dk + fv[g] * omre - rv[g] * (ompr - Tv[g] * sd[g]);     // cj:1461-1523
            double t;
            if (fl & F_PDEP) t = (pmv[g] * xdv[g] * nv[g] + pmv[g] * iv[g] * elem) * ri[g];
            else t = (-pmv[g] * nv[g] * iv[g] + pmv[g] * iv[g] * elem) * ri[g];
            t_[g] = (fl & F_NO_T) ? 0.0 : t;
            double inner = n1 * fv[g] - n2 * rv[g];
            if (fl & F_PMT_INJ) inner += pt[g];
            const double jy = nm[g] * pmv[g] * inner;
            if (fl & F_PMT_INJ) pt[g] *= ef[g];
            x1[g] = jy + par[5] * pt[g];
            x2[g] = -jy - par[4] * pt[g];
        }
        tT = V{t_[0], t_[1]}; X1 = V{x1[0], x1[1]}; X2 = V{x2[0], x2[1]}; pmt = V{pt[0], pt[1]};
    }
    const V pk = vmul(PM_, kf);
    const V prv{-PM_.x * kr.x, -PM_.y * kr.y};

    // d(rate)/dC values: to their raw rows, or folded into X2 for the last species
#define PJ_EMIT(SLOT, DST, EXPR)                                         \
    if ((SLOT) != nsp_f) {                                               \
        const V d_ = (EXPR);                                             \
        if ((SLOT) == last_f) X2 = vsub(X2, d_);                         \
        else if (valid) STS(0, aRAW + (DST) * RB, d_);                   \
    }
    PJ_EMIT(s0, q3.x & 0xFFFFu, vmul(pk, vmul(c1, c2)))
    PJ_EMIT(s1, q3.x >> 16, vmul(pk, vmul(c0, c2)))
    PJ_EMIT(s2, q3.y & 0xFFFFu, vmul(pk, vmul(c0, c1)))
    if (isrev) {
        PJ_EMIT(s3, q3.y >> 16, vmul(prv, vmul(c4, c5)))
        PJ_EMIT(s4, q3.z & 0xFFFFu, vmul(prv, vmul(c3, c5)))
        PJ_EMIT(s5, q3.z >> 16, vmul(prv, vmul(c3, c4)))
    }
#undef PJ_EMIT
    if (valid) {
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
        if (fl & F_WANT_PMT) STS(0, aRAW + (q3.w >> 16) * RB, pmt);
        const unsigned ar = aRX + (unsigned)p * (RX_SLOTS * RB) + ((unsigned)p & 1u) * RB;
        STS(E_NET * RB, ar, vmul(net, PM_));
        STS(E_X1 * RB, ar, X1);
        STS(O_TT * RB, ar ^ RB, tT);
        STS(O_DH * RB, ar ^ RB, dH);
        STS(0, aXC + (unsigned)(p - pl.p_c0) * RB, vadd(X1, X2));     // every pressure-modified reaction has a correction row
    }
}

// Phase B, one reaction without pressure modification (pj5::reaction_plain, Jacobian only)
template <int GS, bool SPECIAL>
__device__ __forceinline__ void reaction_plain(const Mem<false>& mem, const Tables& tb, const Plan6& pl,
                                               unsigned aSP, unsigned aRX, unsigned aXC, unsigned aRAW, unsigned aSC, unsigned aPL,
                                               int p, bool valid, bool three,
                                               const uint4 q0, const uint4 q1, const uint4 q2, const uint4 q3,
                                               const V T, const V logT, const V iT, const V rho_inv, const V nmwr)
{
    constexpr int RB = GS * 8;
    const double lnA = dbl(q0.x, q0.y), bexp = dbl(q0.z, q0.w);
    const double Ta = dbl(q1.x, q1.y), lnKc = dbl(q1.z, q1.w);
    const int fl = (int)q2.x;
    const unsigned s0 = q2.y & 0xFFFFu, s1 = q2.y >> 16, s3 = q2.z >> 16, s4 = q2.w & 0xFFFFu;
    const unsigned a0 = aSP + s0 * 16, a1 = aSP + s1 * 16, a3 = aSP + s3 * 16, a4 = aSP + s4 * 16;
    V c0 = LDS(E_C * RB, a0), c1 = LDS(E_C * RB, a1), c3 = LDS(E_C * RB, a3), c4 = LDS(E_C * RB, a4);
    V sB = vsub(vadd(LDS(O_B * RB, a3 ^ RB), LDS(O_B * RB, a4 ^ RB)), vadd(LDS(O_B * RB, a0 ^ RB), LDS(O_B * RB, a1 ^ RB)));
    V sdB = vsub(vadd(LDS(E_DB * RB, a3), LDS(E_DB * RB, a4)), vadd(LDS(E_DB * RB, a0), LDS(E_DB * RB, a1)));
    V dH = vsub(vadd(LDS(O_HW * RB, a3 ^ RB), LDS(O_HW * RB, a4 ^ RB)), vadd(LDS(O_HW * RB, a0 ^ RB), LDS(O_HW * RB, a1 ^ RB)));
    V c2{1.0, 1.0}, c5{1.0, 1.0};
    const unsigned s2 = q2.z & 0xFFFFu, s5 = q2.w >> 16;
    if (three) {
        const unsigned a2 = aSP + s2 * 16, a5 = aSP + s5 * 16;
        c2 = LDS(E_C * RB, a2);
        c5 = LDS(E_C * RB, a5);
        sB = vadd(sB, vsub(LDS(O_B * RB, a5 ^ RB), LDS(O_B * RB, a2 ^ RB)));
        sdB = vadd(sdB, vsub(LDS(E_DB * RB, a5), LDS(E_DB * RB, a2)));
        dH = vadd(dH, vsub(LDS(O_HW * RB, a5 ^ RB), LDS(O_HW * RB, a2 ^ RB)));
    }
    V lnkf = vfma(bexp, logT, V{fma(-Ta, iT.x, lnA), fma(-Ta, iT.y, lnA)});
    V dk{fma(Ta, iT.x, bexp), fma(Ta, iT.y, bexp)};               // d ln kf / d ln T
    if (SPECIAL && (fl & F_PLOG)) {
        // rate constant of a PLOG reaction (rs:598-632) and its temperature derivative (cj:1687-1850)
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
        // Chebyshev rate constant (rs:149-251) and its temperature derivative (cj:1532-1684); see
        // pj5::reaction_plain for the two sets of reduced variables
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
                const double cc0 = __ldg(r8), cc1 = __ldg(r8 + 1);
                double pa8 = 1.0, pb8 = pr8, pa16 = 1.0, pb16 = pr16;
                double dp8 = fma(pr8, cc1, cc0), dp16 = fma(pr16, cc1, cc0), dpu = fma(pr16, __ldg(r16 + 1), __ldg(r16));
                for (int j = 2; j < n_p; ++j) {
                    const double n8 = fma(2.0 * pr8, pb8, -pa8), n16 = fma(2.0 * pr16, pb16, -pa16);
                    pa8 = pb8; pb8 = n8; pa16 = pb16; pb16 = n16;
                    const double cc = __ldg(r8 + j);
                    dp8 = fma(cc, n8, dp8); dp16 = fma(cc, n16, dp16); dpu = fma(__ldg(r16 + j), n16, dpu);
                }
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
    const double nre = (double)((fl >> NRE_SHIFT) & 15), npr = (double)((fl >> NPR_SHIFT) & 15);
    const double omre = 1.0 - nre, ompr = 1.0 - npr;
    // irreversible: r = 0 makes this f * (dk + 1 - nre)          (cj:1461-1523)
    V elem = vfma(net, dk, vmul(omre, f));
    elem = V{elem.x - r.x * (ompr - T.x * sdB.x), elem.y - r.y * (ompr - T.y * sdB.y)};
    const double tmask = (fl & F_NO_T) ? 0.0 : 1.0;
    const V tT = vmul(tmask, vmul(vmul(iT, rho_inv), elem));
    const V X1 = vmul(nmwr, V{nre * f.x - npr * r.x, nre * f.y - npr * r.y});
    const V n3{-d3.x, -d3.y}, n4{-d4.x, -d4.y};
    V d2 = zero_v(), n5 = zero_v();
    if (three) { d2 = vmul(kf, o2); n5 = vmul(kr, o5); n5 = V{-n5.x, -n5.y}; }
    if (fl & F_CORR) {
        // X1 + X2: a slot holding the last species has no column, its derivative joins the W_j / W_N term
        V xc = zero_v();
        const unsigned last = sp_even<GS>(0u, (unsigned)(tb.nsp - 1)) / 16;
        if (s0 == last) xc = vsub(xc, d0);
        if (s1 == last) xc = vsub(xc, d1);
        if (s2 == last) xc = vsub(xc, d2);
        if (s3 == last) xc = vsub(xc, n3);
        if (s4 == last) xc = vsub(xc, n4);
        if (s5 == last) xc = vsub(xc, n5);
        STS_IF(0, valid, aXC + (unsigned)(p - pl.p_c0) * RB, xc);
    }
    // raw rows; slots without one carry the scratch row index and store nothing
    const unsigned none = tb.nraw + 1;
    auto emit = [&](unsigned dst, V d) { STS_IF(0, valid && dst != none, aRAW + dst * RB, d); };
    emit(q3.x & 0xFFFFu, d0);
    emit(q3.x >> 16, d1);
    emit(q3.y >> 16, n3);
    emit(q3.z & 0xFFFFu, n4);
    if (three) {
        emit(q3.y & 0xFFFFu, d2);
        emit(q3.z >> 16, n5);
    }
    const unsigned ar = aRX + (unsigned)p * (RX_SLOTS * RB) + ((unsigned)p & 1u) * RB;
    STS_IF(E_NET * RB, valid, ar, net);
    STS_IF(E_X1 * RB, valid, ar, X1);
    STS_IF(O_TT * RB, valid, ar ^ RB, tT);
    STS_IF(O_DH * RB, valid, ar ^ RB, dH);
}

template <int GS, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
k_jac6(const __grid_constant__ Tables tb, const __grid_constant__ Plan6 pl, const __grid_constant__ IO io)
{
    extern __shared__ __align__(1024) double smem6[];
    constexpr int NSUB = 64 / GS;
    constexpr int NPR = GS / 2;
    constexpr int RB = GS * 8, SPB = SP_SLOTS * RB, RXB = RX_SLOTS * RB;
    constexpr int SCB = 8 * RB;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nw = pl.nw;
    const int sub = lane / NPR, pr = lane % NPR;
    const int nsp = tb.nsp, last = tb.nsp - 1;
    const unsigned ws0 = (unsigned)__cvta_generic_to_shared(smem6);
    const Mem<false> mem{nullptr};
    const unsigned sb = ws0 + pr * 16;
    if (((ws0 + pl.oSP * 8) | (ws0 + pl.oRX * 8)) & (2 * RB - 1)) __trap();     // E ^ RB needs this
    const unsigned aSP = sb + pl.oSP * 8, aRX = sb + pl.oRX * 8, aXC = sb + pl.oXC * 8, aRAW = sb + pl.oRAW * 8;
    const unsigned aET = sb + pl.oET * 8;
    const unsigned aSC0 = sb + pl.oSC * 8, aPA = sb + pl.oPA * 8;
    const unsigned aSD = aSC0 + 2 * SCB;
    const unsigned aCF = ws0 + pl.oCF * 8;
    const V zero{0.0, 0.0};

    const bool sf = io.jac_layout != 0;
    const bool vec_ok = sf && ((io.jac_ld & 1) == 0) && ((reinterpret_cast<unsigned long long>(io.jac) & 15) == 0);
    const long long nn = (long long)nsp * nsp;
    const long long ngroups = ((long long)io.n + GS - 1) / GS;
    const unsigned ld8 = sf ? (unsigned)(io.jac_ld * 8) : 8u;
    const long long second = sf ? 8 : nn * 8;

    // the warp's record stream: what it does per group, identical for every group
    const int* hd = pl.hdr + warp * 8;
    const int h_ch0 = __ldg(hd), h_nch = __ldg(hd + 1), n_pm = __ldg(hd + 2), n_pl = __ldg(hd + 3);
    const int n_c = __ldg(hd + 4), n_t = __ldg(hd + 5), n_seg = __ldg(hd + 6), n_e = __ldg(hd + 7);
    const long long my_groups = (long long)blockIdx.x < ngroups ? (ngroups - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    Stream<NSUB> rd;
    rd.start(ws0 + pl.ring + warp * (NSLOT * CHB), sub * 16, ws0 + pl.mbar + warp * (NSLOT * 8), ws0 + (pl.oRAW + tb.nraw * GS) * 8,
             reinterpret_cast<const char*>(pl.str) + (size_t)h_ch0 * CHB, (unsigned)h_nch,
             (unsigned)(my_groups * h_nch), lane == 0);

    for (int i = tid; i < nsp; i += blockDim.x) {
        const double2 c = __ldg(pl.colfac + i);
        STS(0, aCF + i * 16, V{c.x, c.y});
    }
    // rows that never change: the empty reaction slot, the all-zero reactions, correction and raw rows
    if (warp == 0 && sub == 0) {
        const unsigned a = sp_even<GS>(aSP, (unsigned)nsp), o = a ^ RB;
        STS(E_C * RB, a, V{1.0, 1.0});
        STS(O_B * RB, o, zero); STS(E_DB * RB, a, zero); STS(O_HW * RB, o, zero);
        STS(E_WT * RB, a, zero); STS(O_CP * RB, o, zero);
        for (int z = 0; z < 2; ++z) {
            const unsigned ar = aRX + (unsigned)(tb.nr + z) * RXB + ((unsigned)(tb.nr + z) & 1u) * RB;
            STS(E_NET * RB, ar, zero); STS(E_X1 * RB, ar, zero); STS(O_TT * RB, ar ^ RB, zero); STS(O_DH * RB, ar ^ RB, zero);
            STS(0, aRAW + (tb.nraw + z) * RB, zero);
            STS(0, aXC + (pl.ncorr + z) * RB, zero);
        }
    }

    auto phase_a0 = [&](long long g, int b) {
        const long long s0 = g * GS + 2 * pr;
        const long long i0 = s0 < io.n ? s0 : (long long)io.n - 1, i1 = s0 + 1 < io.n ? s0 + 1 : (long long)io.n - 1;
        const double* y0 = io.y + i0 * io.y_ss;
        const double* y1 = io.y + i1 * io.y_ss;
        V sumY = zero, sumYW = zero;
        for (int k = sub; k < last; k += 4 * NSUB) {
            V Yk[4];
            double wk[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int kk = k + u * NSUB;
                const bool in = kk < last;
                const long long o = (long long)((in ? kk : k) + 1) * io.y_sv;
                Yk[u] = V{y0[o], y1[o]};
                wk[u] = __ldg(tb.sp_iw + (in ? kk : k));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int kk = k + u * NSUB;
                if (kk < last) {
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
            const double P[2] = {io.pres[i0], io.pres[i1]};
            const double yN[2] = {1.0 - sumY.x, 1.0 - sumY.y};
            const double sw[2] = {sumYW.x, sumYW.y};
            double o[8][2];
#pragma unroll
            for (int g2 = 0; g2 < 2; ++g2) {
                const double mw = 1.0 / (sw[g2] + yN[g2] * __ldg(tb.sp_iw + last));
                const double rho = P[g2] * mw / (tb.ru * T[g2]);
                const double rho_inv = 1.0 / rho;
                o[Q_T][g2] = T[g2]; o[Q_LOGT][g2] = log(T[g2]); o[Q_IT][g2] = 1.0 / T[g2];
                o[Q_RHO][g2] = rho; o[Q_RHOINV][g2] = rho_inv;
                o[Q_LNP][g2] = (tb.nplog | tb.ncheb) ? log(P[g2]) : 0.0; o[Q_MWR][g2] = mw * rho_inv;
                o[Q_M][g2] = P[g2] / (tb.ru * T[g2]);
            }
            const unsigned a = aSC0 + b * SCB;
            STS(E_Y * RB, sp_even<GS>(aSP, (unsigned)last), V{yN[0], yN[1]});
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

    int buf = 0;
    for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x, buf ^= 1) {
        const long long s0 = grp * GS + 2 * pr;
        const bool ok0 = s0 < io.n, ok1 = s0 + 1 < io.n;
        char* const out0 = reinterpret_cast<char*>(sf ? io.jac + s0 : io.jac + s0 * nn);
        const unsigned aSC = aSC0 + buf * SCB;
        const bool fast = vec_ok && ok1;
        auto store = [&](unsigned e, V v, bool on) {
            char* o = out0 + (unsigned long long)e * ld8;
            if (fast) {
                asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\t@q st.global.v2.f64 [%0], {%1, %2};\n\t}"
                             ::"l"(o), "d"(v.x), "d"(v.y), "r"((int)on) : "memory");
            } else {
                if (on && ok0) *reinterpret_cast<double*>(o) = v.x;
                if (on && ok1) *reinterpret_cast<double*>(o + second) = v.y;
            }
        };
        const V T = LDS(Q_T * RB, aSC), logT = LDS(Q_LOGT * RB, aSC), iT = LDS(Q_IT * RB, aSC);

        // ------------------------------------------------------------ phase A1: species thermo
        {
            const V rho = LDS(Q_RHO * RB, aSC);
            const double Tv[2] = {T.x, T.y}, lT[2] = {logT.x, logT.y}, rT[2] = {iT.x, iT.y}, rh[2] = {rho.x, rho.y};
            double cpavg[2] = {0.0, 0.0}, wdcp[2] = {0.0, 0.0};
            for (int k = warp * NSUB + sub; k < nsp; k += nw * NSUB) {
                const unsigned a = sp_even<GS>(aSP, (unsigned)k), o = a ^ RB;
                const V Yv = LDS(E_Y * RB, a);
                const double iw = __ldg(tb.sp_iw + k), ruw = __ldg(tb.sp_ruw + k), wk = __ldg(tb.sp_w + k);
                const double Yk[2] = {Yv.x, Yv.y};
                const double tmid = __ldg(tb.sp_tmid + k);
                double ck[2], cp[2], Bk[2], dB[2], hW[2];
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const double* c = tb.sp_nasa + (k * 2 + (Tv[g] <= tmid ? 0 : 1)) * 16;
                    const double t = Tv[g];
                    ck[g] = rh[g] * Yk[g] * iw;
                    cp[g] = ruw * (c[0] + t * (c[1] + t * (c[2] + t * (c[3] + c[4] * t))));
                    const double hh = c[6] + t * (c[7] + t * (c[8] + c[9] * t));
                    hW[g] = ruw * (c[5] + t * (c[0] + t * hh)) * wk;
                    const double dcp = ruw * (c[1] + t * (2.0 * c[2] + t * (3.0 * c[3] + 4.0 * c[4] * t)));
                    cpavg[g] += Yk[g] * cp[g];
                    wdcp[g] += Yk[g] * dcp;
                    dB[g] = (c[11] + c[5] * rT[g]) * rT[g] + hh;
                    Bk[g] = c[10] + c[11] * lT[g] + t * (c[6] + t * (c[12] + t * (c[13] + c[14] * t))) - c[5] * rT[g];
                }
                STS(E_C * RB, a, V{ck[0], ck[1]});
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
        __syncthreads();

        // ------------------------------------------------------------ phase B: reactions
        {
            const V rho_inv = LDS(Q_RHOINV * RB, aSC), mwr_ = LDS(Q_MWR * RB, aSC), nmwr{-mwr_.x, -mwr_.y};
            const unsigned nsp_f = sp_even<GS>(0u, (unsigned)nsp) / 16;
            for (int r = 0; r < n_pm + n_pl; ++r) {
                const uint4 q0 = rd.get(), q1 = rd.get(), q2 = rd.get(), q3 = rd.get();
                const bool valid = !(q2.x & F_NULL);
                const int p = (int)(q3.w & 0xFFFFu);
                PJ_CHECK(p < tb.nr && (q2.y & 0xFFFFu) <= nsp_f && (q2.w >> 16) <= nsp_f, 1, q2, q3.w);
                if (r < n_pm) {
                    reaction<GS>(mem, tb, pl, aSP, aRX, aXC, aRAW, aSC, p, valid, q0, q1, q2, q3, T, logT, iT, rho_inv, nmwr);
                } else {
                    const bool has3 = ((q2.z & 0xFFFFu) != nsp_f) || ((q2.w >> 16) != nsp_f);
                    const bool three = __any_sync(0xffffffffu, has3);
                    if (__any_sync(0xffffffffu, (q2.x & (F_PLOG | F_CHEB | F_NEGA)) != 0))
                        reaction_plain<GS, true>(mem, tb, pl, aSP, aRX, aXC, aRAW, aSC, aSD + buf * RB, p, valid, three, q0, q1, q2, q3, T, logT, iT, rho_inv, nmwr);
                    else
                        reaction_plain<GS, false>(mem, tb, pl, aSP, aRX, aXC, aRAW, aSC, aSD + buf * RB, p, valid, three, q0, q1, q2, q3, T, logT, iT, rho_inv, nmwr);
                }
            }
        }
        __syncthreads();

        // ------------------------------------------------------------ phase C: species sums, energy-row gathers
        {
            const V mwr = LDS(Q_MWR * RB, aSC);
            V pH1 = zero, pHA = zero, pHB = zero, pHT = zero, pSCP = zero;
            for (int rnd = 0; rnd < n_c; ++rnd) {
                const uint4 h = rd.get();
                const int nm = (h.y >> 8) & 0xFF, nx = (h.y >> 16) & 0xFF;
                PJ_CHECK((h.x == NONE32 || h.x <= (unsigned)nsp * SPB + RB) && nm < 40 && nx < 40 && (h.y & 0xFF) <= 1u, 2, h, 0);
                V aN = zero, aT = zero, a1 = zero, ac = zero;
                for (int u = 0; u < nm; ++u) {
                    const uint4 e = rd.get();
                    const unsigned w4[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const unsigned x = (i & 1) ? (w4[i >> 1] >> 16) : (w4[i >> 1] & 0xFFFFu);
                        const unsigned a = aRX + (x & 0x7FFFu) * RB;
                        const double sg = sgn15(x);
                        aN = vfma(sg, LDS(E_NET * RB, a), aN);
                        aT = vfma(sg, LDS(O_TT * RB, a ^ RB), aT);
                        a1 = vfma(sg, LDS(E_X1 * RB, a), a1);
                    }
                }
                for (int u = 0; u < nx; ++u) {
                    const uint4 e = rd.get();
                    const unsigned w4[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const unsigned x = (i & 1) ? (w4[i >> 1] >> 16) : (w4[i >> 1] & 0xFFFFu);
                        ac = vfma(sgn15(x), LDS(0, aXC + (x & 0x7FFFu) * RB), ac);
                    }
                }
                for (int o = NPR; o < NPR * pl.coop; o <<= 1) {
                    aN = V{aN.x + __shfl_xor_sync(0xffffffffu, aN.x, o), aN.y + __shfl_xor_sync(0xffffffffu, aN.y, o)};
                    aT = V{aT.x + __shfl_xor_sync(0xffffffffu, aT.x, o), aT.y + __shfl_xor_sync(0xffffffffu, aT.y, o)};
                    a1 = V{a1.x + __shfl_xor_sync(0xffffffffu, a1.x, o), a1.y + __shfl_xor_sync(0xffffffffu, a1.y, o)};
                    ac = V{ac.x + __shfl_xor_sync(0xffffffffu, ac.x, o), ac.y + __shfl_xor_sync(0xffffffffu, ac.y, o)};
                }
                if (h.y & 1u) {
                    const unsigned a = aSP + h.x, o = a ^ RB;       // h.x: even-slot base of species k
                    const double wk = dbl(h.z, h.w);
                    a1 = vadd(a1, vmul(aN, mwr));
                    const V a2 = vsub(ac, a1);                      // sum nu X2 - comp = -(sum nu X1 + comp) + sum nu (X1 + X2)
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
            for (int rnd = 0; rnd < n_t; ++rnd) {
                const uint4 h = rd.get();
                const int n = (int)h.y;
                PJ_CHECK((h.x & 0xFFFFu) < (unsigned)nsp && n < 100 && h.z == 0u && h.w == 0u, 3, h, 0);
                V acc0 = zero, acc1 = zero;
                for (int u = 0; u < n; ++u) {
                    const uint4 e = rd.get();
                    acc0 = vfma(LDS(O_DH * RB, (aRX + (e.x >> 16) * RB) ^ RB), LDS(0, aRAW + (e.x & 0xFFFFu) * RB), acc0);
                    acc1 = vfma(LDS(O_DH * RB, (aRX + (e.y >> 16) * RB) ^ RB), LDS(0, aRAW + (e.y & 0xFFFFu) * RB), acc1);
                    acc0 = vfma(LDS(O_DH * RB, (aRX + (e.z >> 16) * RB) ^ RB), LDS(0, aRAW + (e.z & 0xFFFFu) * RB), acc0);
                    acc1 = vfma(LDS(O_DH * RB, (aRX + (e.w >> 16) * RB) ^ RB), LDS(0, aRAW + (e.w & 0xFFFFu) * RB), acc1);
                }
                V E0 = vadd(acc0, acc1);
                for (int o = NPR; o < NPR * pl.tcoop; o <<= 1)
                    E0 = V{E0.x + __shfl_xor_sync(0xffffffffu, E0.x, o), E0.y + __shfl_xor_sync(0xffffffffu, E0.y, o)};
                if (h.x >> 16) STS(0, aET + ((h.x & 0xFFFFu) - 1u) * RB, E0);
            }
            pH1 = sub_sum<GS>(pH1); pHA = sub_sum<GS>(pHA); pHB = sub_sum<GS>(pHB);
            pHT = sub_sum<GS>(pHT); pSCP = sub_sum<GS>(pSCP);
            if (sub == 0) {
                const unsigned a = aPA + warp * NPART * RB;
                STS(D_H1 * RB, a, pH1); STS(D_HA * RB, a, pHA); STS(D_HB * RB, a, pHB);
                STS(D_HT * RB, a, pHT); STS(D_SCP * RB, a, pSCP);
            }
        }
        __syncthreads();

        // ------------------------------------------------------------ phase DE
        if (warp == 0) {
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
            __syncwarp();
            if (pl.t_sync > 32) {
                __threadfence_block();
                asm volatile("bar.arrive 1, %0;" ::"r"(pl.t_sync) : "memory");
            }
            // the next group's phase A0 (its inputs come from HBM: latency hidden behind DE)
            if (grp + gridDim.x < ngroups) phase_a0(grp + gridDim.x, buf ^ 1);
        }
        // Elements by rows.  A segment's header hands every sub-group the species rows of one Jacobian
        // row (W_k a_k, W_k b_k, W_k stay in registers), a second record the number of records of
        // each kind; then the records kind by kind (plan6.py): K7 / K6 / K4 one element with up to six
        // signed raw rows, K2 three elements with two, K1 five with one, K0 sixteen dense-only ones.
        // The element index follows from the column: e = col * NSP + k + 1; col = 0: no element.
        auto de_phase = [&](auto fast_tag) {
            constexpr bool FAST = decltype(fast_tag)::value;
            V wa = zero, wb = zero, cp_ = zero, cm_ = zero;      // row constants; accumulators carried between records
            double wk = 0.0;
            unsigned kp1 = 0;
            const unsigned k3ff = 0x3FF00000u;
            // +1.0 / -1.0 from bit 15 of the low (shift = 16) or high (shift = 0) half of a record word
            auto sgn = [&](unsigned w, int shift) {
                unsigned hi;
                asm("lop3.b32 %0, %1, 0x80000000, %2, 0xEA;" : "=r"(hi) : "r"(w << shift), "r"(k3ff));
                return __hiloint2double((int)hi, 0);
            };
            auto lo = [&](unsigned w) { return LDS(0, aRAW + (w & 0x7FFFu) * RB); };
            auto hi = [&](unsigned w) { return LDS(0, aRAW + ((w >> 16) & 0x7FFFu) * RB); };
            auto put = [&](unsigned e, V v, unsigned on) {
                char* o;
                asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(o) : "r"(e), "r"(ld8), "l"(out0));
                if (FAST) {
                    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q st.global.v2.f64 [%0], {%1, %2};\n\t}"
                                 ::"l"(o), "d"(v.x), "d"(v.y), "r"(on) : "memory");
                } else if (on) {
                    if (ok0) *reinterpret_cast<double*>(o) = v.x;
                    if (ok1) *reinterpret_cast<double*>(o + second) = v.y;
                }
            };
            // element of column col = (1 / W_j) tt + (1 / W_N) W_k b_k,  tt = W_k a_k + W_k S_kj
            auto finish = [&](unsigned col, V tt) {
                const V cf = LDS(0, aCF + col * 16);
                put(col * (unsigned)nsp + kp1, vfma(cf.y, wb, vmul(cf.x, tt)), col);
            };
            for (int sg_ = 0; sg_ < n_seg; ++sg_) {
                const uint4 h = rd.get(), cw = rd.get();
                {
                    const unsigned x = aSP + h.y;
                    wa = LDS(E_WA * RB, x);
                    wb = LDS(O_WB * RB, x ^ RB);
                    wk = dbl(h.z, h.w);
                    kp1 = ((h.x >> 8) & 0xFFu) + 1u;
                    put(kp1, LDS(E_WT * RB, x), h.x & 1u);       // temperature column: W_k * T-term
                }
                const int n7 = cw.x & 0xFF, n6 = (cw.x >> 8) & 0xFF, n4 = (cw.x >> 16) & 0xFF, n2 = cw.x >> 24;
                const int n1 = cw.y & 0xFF, n0 = (cw.y >> 8) & 0xFF;
                for (int i = 0; i < n7; ++i) {                   // lists longer than six span records
                    const uint4 r = rd.get();
                    const V e0 = lo(r.y), e1 = hi(r.y), e2 = lo(r.z), e3 = hi(r.z), e4 = lo(r.w), e5 = hi(r.w);
                    V p = zero, m = zero;
                    if (r.x & D_CIN) { p = cp_; m = cm_; }
                    p = vfma(sgn(r.w, 16), e4, vfma(sgn(r.z, 16), e2, vfma(sgn(r.y, 16), e0, p)));
                    m = vfma(sgn(r.w, 0), e5, vfma(sgn(r.z, 0), e3, vfma(sgn(r.y, 0), e1, m)));
                    if (r.x & D_COUT) { cp_ = p; cm_ = m; }
                    else finish(r.x & 0xFFu, vfma(wk, vadd(p, m), wa));
                }
                for (int i = 0; i < n6; ++i) {
                    const uint4 r = rd.get();
                    const V e0 = lo(r.y), e1 = hi(r.y), e2 = lo(r.z), e3 = hi(r.z), e4 = lo(r.w), e5 = hi(r.w);
                    const V p = vfma(sgn(r.w, 16), e4, vfma(sgn(r.z, 16), e2, vmul(sgn(r.y, 16), e0)));
                    const V m = vfma(sgn(r.w, 0), e5, vfma(sgn(r.z, 0), e3, vmul(sgn(r.y, 0), e1)));
                    finish(r.x & 0xFFu, vfma(wk, vadd(p, m), wa));
                }
                for (int i = 0; i < n4; ++i) {
                    const uint4 r = rd.get();
                    const V e0 = lo(r.y), e1 = hi(r.y), e2 = lo(r.z), e3 = hi(r.z);
                    const V p = vfma(sgn(r.z, 16), e2, vmul(sgn(r.y, 16), e0)), m = vfma(sgn(r.z, 0), e3, vmul(sgn(r.y, 0), e1));
                    finish(r.x & 0xFFu, vfma(wk, vadd(p, m), wa));
                }
                for (int i = 0; i < n2; ++i) {
                    const uint4 r = rd.get();
                    const V a0 = lo(r.y), a1 = hi(r.y), b0 = lo(r.z), b1 = hi(r.z), c0 = lo(r.w), c1 = hi(r.w);
                    finish(r.x & 0xFFu, vfma(wk, vfma(sgn(r.y, 0), a1, vmul(sgn(r.y, 16), a0)), wa));
                    finish((r.x >> 8) & 0xFFu, vfma(wk, vfma(sgn(r.z, 0), b1, vmul(sgn(r.z, 16), b0)), wa));
                    finish((r.x >> 16) & 0xFFu, vfma(wk, vfma(sgn(r.w, 0), c1, vmul(sgn(r.w, 16), c0)), wa));
                }
                for (int i = 0; i < n1; ++i) {
                    const uint4 r = rd.get();
                    const V a0 = hi(r.y), a1 = lo(r.z), a2 = hi(r.z), a3 = lo(r.w), a4 = hi(r.w);
                    finish(r.x & 0xFFu, vfma(wk, vmul(sgn(r.y, 0), a0), wa));
                    finish((r.x >> 8) & 0xFFu, vfma(wk, vmul(sgn(r.z, 16), a1), wa));
                    finish((r.x >> 16) & 0xFFu, vfma(wk, vmul(sgn(r.z, 0), a2), wa));
                    finish(r.x >> 24, vfma(wk, vmul(sgn(r.w, 16), a3), wa));
                    finish(r.y & 0xFFu, vfma(wk, vmul(sgn(r.w, 0), a4), wa));
                }
                for (int i = 0; i < n0; ++i) {
                    const uint4 r = rd.get();
                    const unsigned w4[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                    for (int j = 0; j < 16; ++j) finish((w4[j >> 2] >> (8 * (j & 3))) & 0xFFu, wa);
                }
            }
        };
        if (__all_sync(0xffffffffu, fast)) de_phase(std::true_type{});
        else de_phase(std::false_type{});
        if (n_e) {
            // the energy-equation row (cj:3095-3254) from the gathers of phase C and warp 0's scalars
            if (warp != 0) asm volatile("bar.sync 1, %0;" ::"r"(pl.t_sync) : "memory");
            const V nwt = LDS(S_NWT * RB, aSD), A0 = LDS(S_A0 * RB, aSD), B0 = LDS(S_B0 * RB, aSD);
            const V XT = LDS(S_XT * RB, aSD), cpl = LDS(S_CPL * RB, aSD);
            for (int i = 0; i < n_e; ++i) {
                const uint4 r = rd.get();
                const unsigned col = r.x;
                PJ_CHECK(col < (unsigned)nsp && r.y == 0u, 6, r, (unsigned)i);
                const bool on = col != 0u;
                const unsigned j = on ? col - 1u : 0u;
                const V cf = LDS(0, aCF + col * 16);
                const V cpj = LDS(O_CP * RB, sp_even<GS>(aSP, j) ^ RB);
                const V E0 = LDS(0, aET + j * RB);
                V v = vmul(cf.x, vfma(nwt, E0, A0));
                v = vfma(cf.y, B0, v);
                v = vfma(XT, vsub(cpj, cpl), v);
                store(col * (unsigned)nsp, v, on);
            }
        }
        rd.align();
        __syncthreads();
    }
}

#undef LDS
#undef STS
#undef STS_IF

}  // namespace pj6
