"""Mechanism -> device tables for the data-driven sm_100a kernels.

This replaces the reference's *source generators* (pyjac/core/rate_subs.py and
pyjac/core/create_jacobian.py): instead of unrolling the mechanism into C/CUDA text, the
same information is exported once as flat arrays (packed with :mod:`pyjac_b200.blob`) that
one fixed kernel family (pyjac_b200/csrc/eval.cuh) interprets.

The kernel does not follow the generated code's statement order; it uses the algebraic
structure of the emitted Jacobian (SURVEY.md section 7 / 8a'):

    jac[k+1, j+1] = W_k/W_j * ( A_k + B_k * W_j/W_N + S_kj )
    A_k = sum_i nu_ki X1_i + wdot_k mw_avg/rho         B_k = sum_i nu_ki X2_i - wdot_k mw_avg/rho
    S_kj = sum over reactions i, sum over "raw" per-reaction derivative values r:
           nu * raw[r]        (only for j among reactants/products/listed colliders of i)

so every Jacobian element is produced once, from a dense rank-2 part and a sparse gather.
Where the generator prints a constant with a lossy format string and the difference can
exceed ~1e-16 relative, the same quantisation is applied here (citations: rs =
pyjac/core/rate_subs.py, cj = pyjac/core/create_jacobian.py).

Table reference (all reaction-indexed arrays are in *kernel order*, see ``rx_orig``):

  dims      int32[16]  NSP NR NREV NPD NRAW NPLOG NCHEB - FIRST_PM NPM NRED MAXRED - - - -
  cst       f64[4]     RU ({:.8e}), ln(PA/RU)
  sp_*      per species (internal, moved-last order): w, iw (=1/W {:.16e}), ruw (=RU/W),
            tmid, mwf (=W_j/W_N), seen;  sp_nasa[k][branch][16] polynomial coefficients
  rx_*      orig (original reaction index), flags, rev_idx, pm_idx (positions in the
            reference's rev_rates / pres_mod arrays), raw_base, slots[6] (3 reactant + 3
            product species, NSP = empty slot), arr[4] = lnA, b, Ta, sum(nu)*ln(PA/RU),
            dst[8] (raw row written per slot, first collider row, pres_mod_temp row);
            the kernels read these through the packed records p5_rx / p5_rxout / p5_eff
  plog_*    off[NR + 1] (kernel order) into par[NPLOG][8] = threshold ({:.4e} Pa), ln A, b, Ta,
            ln P, 1 / (ln P' - ln P), b' - b, Ta' - Ta
  cheb_*    off[NR + 1] (kernel order, in doubles) into par: per Chebyshev reaction a 12-double
            header (n_T, n_P, reduced-variable constants) and two n_T x n_P coefficient blocks
  pm_*      per pressure-modified reaction (kernel index - FIRST_PM): collider list
            (eff_off/eff_sp/eff_am1 = alpha-1), sp (specific collider or -1), par[32]
  red_*     per species CSR of (reaction, nu) (eval_spec_rates entry point, plan input)
  p5_*      the records and the static work schedule of k_eval: see :mod:`pyjac_b200.plan`
"""
from __future__ import annotations

import math
from typing import Dict, List

import numpy as np

from . import plan, plan6
from .chem import PA, RU
from .mechanism import Mechanism

# flag bits shared with csrc/common.cuh
F_REV, F_THD, F_PDEP, F_LOW, F_TROE, F_SRI = 1, 2, 4, 8, 16, 32
F_PMT, F_PMT_INJ, F_TROE_T2, F_SRI5, F_SRI5_DT, F_NO_T = 64, 128, 256, 512, 1024, 2048
F_EFFN1 = 4096         # third-body (non fall-off) reaction with a collider list: n' += 1
F_HAS_LAST = 1 << 13   # (Jacobian kernel record only) an occupied slot holds the last species
F_NEGA = 1 << 14       # A < 0 (rs:108-141): the record holds log|A|, kf and kr take the sign
F_WANT_PMT = 1 << 16   # the kernel stores pres_mod_temp as a raw value
F_EFF_SLOTS = 1 << 17  # ... and pres_mod_temp * (alpha_j - 1) for each listed collider j
F_PLOG = 1 << 18       # rate constant interpolated in log P between Arrhenius sets (plog_*)
F_CHEB = 1 << 19       # Chebyshev rate constant in (1/T, log10 P) (cheb_*)
NCHEB_HDR = 12         # doubles ahead of a Chebyshev reaction's coefficient blocks
NRE_SHIFT, NPR_SHIFT = 20, 24    # occupied reactant / product slots

MAXS = 3               # concentration slots per side of a reaction
UNROLL = 40            # CParams.Jacob_Unroll (cj:2651): scope of the stale pres_mod_temp quirk
NPAR = 32
SCHEMA_VERSION = 2     # of the table set; the library refuses blobs written for another one (csrc/pjtable.h)


class UnsupportedMechanism(NotImplementedError):
    pass


def q(fmt: str, x: float) -> float:
    """x after a round trip through one of the generator's format strings."""
    return float(fmt.format(x))


def _slots(sp: List[int], nu: list, empty: int, what: str) -> List[int]:
    out: List[int] = []
    for s, n in zip(sp, nu):
        if not float(n).is_integer() or n < 0:
            raise UnsupportedMechanism('non-integer stoichiometric coefficient in ' + what)
        out += [s] * int(n)
    if len(out) > MAXS:
        raise UnsupportedMechanism('%s has more than %d molecules on one side' % (what, MAXS))
    return out + [empty] * (MAXS - len(out))


def _nasa_row(a) -> List[float]:
    """Coefficient forms the generator prints (rs:540-561, 1834-1857, 2049-2066; cj:829-858)."""
    return [a[0], a[1], a[2], a[3], a[4],
            a[5], a[1] / 2.0, a[2] / 3.0, a[3] / 4.0, a[4] / 5.0,
            a[6] - a[0], a[0] - 1.0, a[2] / 6.0, a[3] / 12.0, a[4] / 20.0, 0.0]


def build(mech: Mechanism, gs: int = 0, threads: int = 0, ws_global=None, streams=None, conv: bool = False) -> Dict[str, np.ndarray]:
    """gs / threads: states per block and block size of the Jacobian kernel's plan (0 = automatic).
    ws_global: True puts the per-block working set in global memory instead of shared memory, False
    forbids that; None = automatic (global memory when fewer than plan.SMEM_MIN_GS states fit in
    shared memory: USC-II- and n-heptane-sized mechanisms).
    streams: True adds the record streams of k_jac6 (plan6.py: table items staged through a shared-memory
    ring by bulk-asynchronous copies) and makes eval_jacob run on them when the working set lives in
    shared memory; None / False = eval_jacob on the schedule tables of k_eval like dydt and the rate
    routines (measured faster on a B200: profiles/README.md).
    conv: the reference-named dydt of a library loaded from these tables is the constant-volume one (what
    `#define CONV` in header.h selects in the reference)."""
    specs, reacs = mech.specs, mech.reacs
    nsp, nr = len(specs), len(reacs)
    last = nsp - 1
    if nsp < 2:
        raise UnsupportedMechanism('need at least two species')
    if nsp >= 0xFFFF:
        raise UnsupportedMechanism('too many species')
    rev_reacs, pdep_reacs = mech.rev_reacs, mech.pdep_reacs
    f64 = lambda x: np.asarray(x, dtype=np.float64)
    i32 = lambda x: np.asarray(x, dtype=np.int32)
    T: Dict[str, np.ndarray] = {}

    for i, rx in enumerate(reacs):
        if rx.cheb and (rx.pdep or rx.thd_body or rx.plog or rx.cheb_n_temp < 2 or rx.cheb_n_pres < 2):
            raise UnsupportedMechanism('Chebyshev reaction %d: third body / PLOG / fewer than 2 x 2 '
                                       'coefficients' % i)
        if rx.plog:
            pp = rx.plog_par
            if rx.pdep or rx.thd_body:
                raise UnsupportedMechanism('PLOG reaction %d with a third body' % i)
            if len(pp) < 2 or any(not e[1] > 0 for e in pp) or \
                    any(q('{:.4e}', pp[e + 1][0]) <= q('{:.4e}', pp[e][0]) for e in range(len(pp) - 1)):
                # the reference does not handle these either (cj:1744-1749, 1768)
                raise UnsupportedMechanism('PLOG reaction %d: pressures must ascend, A > 0' % i)
        if rx.plog or rx.cheb:
            pass        # the rate constant comes from the PLOG / Chebyshev sets; A is not used
        elif rx.A < 0 and not rx.pdep:
            pass        # rs:108-141 (a duplicate with a negative rate); flagged F_NEGA below
        elif not rx.A > 0:
            # A == 0: the reference raises as well (rs:143-144); A < 0 in a fall-off reaction
            # makes its generated code take the logarithm of a negative reduced pressure
            raise UnsupportedMechanism('non-positive pre-exponential (reaction %d)' % i)
        if rx.pdep and not (rx.low or rx.high):
            raise UnsupportedMechanism('fall-off reaction %d without LOW or HIGH' % i)

    # ---------------- species
    mwN = specs[last].mw
    T['sp_w'] = f64([sp.mw for sp in specs])
    T['sp_iw'] = f64([q('{:.16e}', 1.0 / sp.mw) for sp in specs])            # rs:1678,1698
    T['sp_ruw'] = f64([q('{:.16e}', RU / sp.mw) for sp in specs])            # rs:1834,2049
    T['sp_tmid'] = f64([sp.Trange[1] for sp in specs])
    T['sp_mwf'] = f64([q('{:.16e}', sp.mw / mwN) for sp in specs])           # cj:467
    T['sp_nasa'] = f64([[_nasa_row(sp.lo), _nasa_row(sp.hi)] for sp in specs]).ravel()
    seen = [False] * nsp
    for rx in reacs:
        for k in set(rx.reac + rx.prod):
            if rx.net_nu(k) != 0:
                seen[k] = True
    T['sp_seen'] = i32(seen)                                                   # rs:1425-1527

    # ---------------- kernel order: plain, third-body, fall-off; reversible first
    # Reactions without pressure modification that hold the last species come last among their
    # kind: from `p_c0` on a reaction may have X1 + X2 != 0 (plan6: correction rows)
    def sort_key(i):
        rx = reacs[i]
        cls = 2 if rx.pdep else (1 if rx.thd_body else 0)
        sub = (1 if rx.troe else (2 if rx.sri else 0)) if rx.pdep else 0
        has_last = 1 if (cls == 0 and last in (rx.reac + rx.prod)) else 0
        return (cls, sub, has_last, 0 if rx.rev else 1, i)
    order = sorted(range(nr), key=sort_key)
    pos_of = {orig: p for p, orig in enumerate(order)}
    npm = len(pdep_reacs)
    first_pm = nr - npm

    # the stale pres_mod_temp of cj:154,226 (a specific collider that is species 0 is tested
    # by truthiness): the value left behind by the closest earlier reaction that assigned it
    def has_pmt(rx):
        return bool((rx.pdep or rx.thd_body) and (rx.thd_body_eff or rx.pdep_sp))
    stale_src = {}
    for i, rx in enumerate(reacs):
        if rx.pdep and rx.pdep_sp == 0 and not has_pmt(rx):
            lo = (i // UNROLL) * UNROLL if nr > UNROLL else 0
            src = [s for s in range(lo, i) if has_pmt(reacs[s])]
            stale_src[i] = src[-1] if src else None

    flags, rev_idx, pm_idx, raw_base = [], [], [], []
    plog_off, plog_par = [0], []
    cheb_off, cheb_par = [], []
    slots = np.full((nr, 2 * MAXS), nsp, dtype=np.int32)
    arr = np.zeros((nr, 4))
    pm_par = np.zeros((max(npm, 1), NPAR))
    pm_sp = np.full(max(npm, 1), -1, dtype=np.int32)
    eff_off, eff_sp, eff_am1 = [0], [], []
    ln_pa_ru = math.log(PA / RU)

    def eff_case(rx):
        return bool(((rx.pdep and rx.pdep_sp is None) or rx.thd_body) and rx.thd_body_eff)

    # raw slot bookkeeping: for kernel reaction p the kernel stores, in this order,
    #   one value per occupied reactant slot whose species is not the last one,
    #   one value per occupied product slot (reversible only) likewise,
    #   (collider list reactions) pres_mod_temp * (alpha_j - 1) per listed j != last, alpha_j != 1,
    #   pres_mod_temp itself if ``want_pmt`` (specific collider, or source of a stale read).
    want_pmt = [False] * nr
    for i, rx in enumerate(reacs):
        if has_pmt(rx) and not eff_case(rx) and rx.pdep_sp is not None and rx.pdep_sp != last:
            want_pmt[i] = True
    for i, src in stale_src.items():
        if src is not None:
            want_pmt[src] = True

    raw_of_slot: List[List[int]] = [None] * nr      # per original reaction: raw index per slot / -1
    raw_of_eff: List[Dict[int, int]] = [dict() for _ in range(nr)]
    raw_of_pmt = [-1] * nr
    raw_rxn: List[int] = []                         # kernel reaction index of every raw slot
    nraw = 0
    for p, i in enumerate(order):
        rx = reacs[i]
        what = 'reaction %d' % i
        rs = _slots(rx.reac, rx.reac_nu, nsp, what)
        ps = _slots(rx.prod, rx.prod_nu, nsp, what)
        slots[p, :MAXS] = rs
        slots[p, MAXS:] = ps
        fl = 0
        if rx.rev:
            fl |= F_REV
        if rx.thd_body:
            fl |= F_THD
        if rx.pdep:
            fl |= F_PDEP
            if rx.low:
                fl |= F_LOW
        if rx.troe:
            fl |= F_TROE
        if rx.sri:
            fl |= F_SRI
        if has_pmt(rx):
            fl |= F_PMT
        if rx.pdep and (rx.pdep_sp or rx.thd_body_eff):
            fl |= F_PMT_INJ
        if rx.thd_body_eff and not rx.pdep:
            fl |= F_EFFN1                                                      # cj:201-206
        b_on, E_on = abs(rx.b) > 1.0e-90, abs(rx.E) > 1.0e-90
        if not rx.rev and not b_on and not E_on and sum(rx.reac_nu) == 1.0 and not rx.plog and not rx.cheb:
            fl |= F_NO_T                                                       # cj:1507-1523
        if rx.plog:
            # per pressure: threshold as printed ({:.4e}: rs:601-629), ln A, b, Ta, ln P, and towards
            # the next pressure 1 / (ln P' - ln P), b' - b, Ta' - Ta  (rs:598-632, cj:1738-1770)
            fl |= F_PLOG
            pp = rx.plog_par
            for e, (p1, A1, b1, E1) in enumerate(pp):
                row = [q('{:.4e}', p1), q('{:.16e}', math.log(A1)), b1, q('{:.16e}', E1),
                       q('{:.16e}', math.log(p1)), 0.0, 0.0, 0.0]
                if e + 1 < len(pp):
                    p2, _, b2, E2 = pp[e + 1]
                    row[5:8] = [1.0 / q('{:.16e}', math.log(p2) - math.log(p1)),
                                q('{:.16e}', b2 - b1), q('{:.16e}', E2 - E1)]
                plog_par.append(row)
        plog_off.append(len(plog_par))
        cheb_off.append(len(cheb_par))
        if rx.cheb:
            # header: n_T, n_P; (tsum, tsub, psum, psub) as eval_rxn_rates prints them ({:.8e},
            # rs:175-192) and as eval_jacob does ({:.16e}, cj:1641-1660); -2 ln10 / tsub (cj:1665).
            # Then the n_T x n_P coefficients of the rate ({:.8e}, rs:197-217) and those of the
            # temperature derivative, row i scaled by i ({:.16e}, cj:1555-1575)
            fl |= F_CHEB
            tsum = 1.0 / rx.cheb_tlim[0] + 1.0 / rx.cheb_tlim[1]
            tsub = 1.0 / rx.cheb_tlim[1] - 1.0 / rx.cheb_tlim[0]
            psum = math.log10(rx.cheb_plim[0]) + math.log10(rx.cheb_plim[1])
            psub = math.log10(rx.cheb_plim[1]) - math.log10(rx.cheb_plim[0])
            cheb_par += [float(rx.cheb_n_temp), float(rx.cheb_n_pres),
                         q('{:.8e}', tsum), q('{:.8e}', tsub), q('{:.8e}', psum), q('{:.8e}', psub),
                         q('{:.16e}', tsum), q('{:.16e}', tsub), q('{:.16e}', psum), q('{:.16e}', psub),
                         q('{:.16e}', -2.0 * math.log(10) / tsub), 0.0]
            cheb_par += [q('{:.8e}', v) for row in rx.cheb_par for v in row]
            cheb_par += [q('{:.16e}', r_ * v) for r_, row in enumerate(rx.cheb_par) for v in row]
        rev_idx.append(rev_reacs.index(i) if rx.rev else -1)
        pm_idx.append(pdep_reacs.index(i) if (rx.thd_body or rx.pdep) else -1)
        if want_pmt[i]:
            fl |= F_WANT_PMT
        if eff_case(rx):
            fl |= F_EFF_SLOTS
        fl |= sum(1 for s in rs if s != nsp) << NRE_SHIFT
        fl |= sum(1 for s in ps if s != nsp) << NPR_SHIFT

        # Arrhenius (rs:27-146, A > 0: always the exponential form for parsed floats)
        sum_nu = sum(rx.prod_nu) - sum(rx.reac_nu)
        if rx.plog or rx.cheb:
            lnA = 0.0
        elif rx.A < 0:
            fl |= F_NEGA
            lnA = math.log(q('{:.16e}', -rx.A))
        else:
            lnA = q('{:.16e}', math.log(rx.A))
        arr[p] = [lnA, rx.b, q('{:.16e}', rx.E),
                  float(sum_nu) * ln_pa_ru if rx.rev else 0.0]

        raw_base.append(nraw)
        ros = []
        for s in rs:
            if s != nsp and s != last:
                ros.append(nraw)
                nraw += 1
            else:
                ros.append(-1)
        for s in ps:
            if rx.rev and s != nsp and s != last:
                ros.append(nraw)
                nraw += 1
            else:
                ros.append(-1)
        raw_of_slot[i] = ros
        if eff_case(rx):
            for s, a in rx.thd_body_eff:
                if a != 1.0 and s != last:
                    raw_of_eff[i][s] = nraw
                    nraw += 1
        if want_pmt[i]:
            raw_of_pmt[i] = nraw
            nraw += 1
        raw_rxn += [p] * (nraw - raw_base[-1])

        if rx.thd_body or rx.pdep:
            m = p - first_pm
            assert m >= 0
            par = pm_par[m]
            for s, a in rx.thd_body_eff:
                if a != 1.0:
                    eff_sp.append(s)
                    eff_am1.append(a - 1.0)                                    # rs:1128-1130
            eff_off.append(len(eff_sp))
            pm_sp[m] = rx.pdep_sp if rx.pdep_sp is not None else -1
            if eff_case(rx):
                par[4] = next((a for s, a in rx.thd_body_eff if s == last), 1.0)
                par[5] = 1.0
            elif rx.pdep_sp == last and has_pmt(rx):
                par[4] = 1.0
            if rx.pdep:
                k0 = rx.low if rx.low else [rx.A, rx.b, rx.E]
                kinf = [rx.A, rx.b, rx.E] if rx.low else rx.high
                beta, Ea = k0[1] - kinf[1], k0[2] - kinf[2]                   # cj:641-655
                par[0] = q('{:.16e}', math.log(k0[0] / kinf[0]))
                par[1] = q('{:.16e}', beta)
                par[2] = q('{:.16e}', Ea)
                par[3] = q('{:.4e}', beta)                                    # cj:1167
            if rx.troe:
                a, T3, T1 = rx.troe_par[:3]
                T2 = rx.troe_par[3] if len(rx.troe_par) == 4 else 0.0
                if len(rx.troe_par) == 4 and T2 != 0.0:
                    fl |= F_TROE_T2
                par[6:14] = [q('{:.16e}', 1.0 - a), q('{:.16e}', -T3), q('{:.16e}', a),
                             q('{:.16e}', -T1), q('{:.16e}', -T2),
                             q('{:.16e}', -(1.0 - a) / T3), q('{:.16e}', a / T1),
                             q('{:.16e}', T2)]                                 # cj:1083-1090,1262-1282
                # reciprocals of par[7], par[9] for the Jacobian kernel (T / (-T3) as a product)
                par[28] = 1.0 / par[7] if par[7] != 0.0 else 0.0
                par[29] = 1.0 / par[9] if par[9] != 0.0 else 0.0
            elif rx.sri:
                s_ = rx.sri_par
                five = len(s_) == 5
                if five and s_[3] != 1.0 and s_[4] != 0.0:
                    fl |= F_SRI5
                if five and s_[4] != 0.0:
                    fl |= F_SRI5_DT
                d, e = (s_[3], s_[4]) if five else (1.0, 0.0)
                par[14:19] = [q('{:.6}', s_[0]), q('{:.6}', s_[1]), q('{:.6}', s_[2]),
                              q('{:.8e}', d), q('{:.6}', e)]                   # rs:1239-1255
                par[19:22] = [q('{:.4}', s_[0]), q('{:.4}', -s_[1]), q('{:.4}', -s_[2])]  # cj:173-180
                par[22:28] = [q('{:.16}', s_[0] * s_[1]), q('{:.16}', -s_[1]),
                              q('{:.16e}', 1.0 / s_[2]), q('{:.16}', -s_[2]),
                              q('{:.16}', s_[0]), q('{:.16}', e)]              # cj:1215-1235
        flags.append(fl)
    if not npm:
        eff_off = [0, 0]
    if nraw >= 0xFFFF or nr >= 0xFFFF:
        raise UnsupportedMechanism('mechanism too large for 16-bit sparse indices')

    # one 64-byte record per reaction: 4 doubles (lnA, b, Ta, sum(nu) ln(PA/RU)) then 8 ints
    # (flags, raw_base, slots packed two per int, rev_idx, pm_idx, original index)
    rec = np.zeros((nr, 16), dtype=np.int32)
    rec[:, :8] = arr.view(np.int32).reshape(nr, 8)
    rec[:, 8] = flags
    rec[:, 9] = raw_base
    for a in range(3):
        rec[:, 10 + a] = slots[:, 2 * a] | (slots[:, 2 * a + 1] << 16)
    rec[:, 13] = rev_idx
    rec[:, 14] = pm_idx
    rec[:, 15] = order
    T['rx_rec'] = rec.ravel()
    T['rx_orig'] = i32(order)
    T['rx_flags'] = i32(flags)
    T['rx_rev_idx'] = i32(rev_idx)
    T['rx_pm_idx'] = i32(pm_idx)
    T['rx_raw_base'] = i32(raw_base)
    T['rx_slots'] = slots.ravel()
    T['rx_arr'] = arr.ravel()
    T['cheb_off'] = i32(cheb_off + [len(cheb_par)])
    T['cheb_par'] = f64(cheb_par or [0.0])
    T['plog_off'] = i32(plog_off)
    T['plog_par'] = f64(plog_par or [[0.0] * 8]).ravel()
    T['pm_par'] = pm_par.ravel()
    T['pm_sp'] = pm_sp
    T['pm_eff_off'] = i32(eff_off)
    T['pm_eff_sp'] = i32(eff_sp if eff_sp else [0])
    T['pm_eff_am1'] = f64(eff_am1 if eff_am1 else [0.0])

    # ---------------- species-side reductions: wdot, T column, A, B share one (reaction, nu)
    # list per species
    red = [[] for _ in range(nsp)]
    for p, i in enumerate(order):
        rx = reacs[i]
        for k in sorted(set(rx.reac + rx.prod)):
            nu = rx.net_nu(k)
            if nu != 0:
                red[k].append((p, float(nu)))
    red_off = [0]
    for k in range(nsp):
        red_off.append(red_off[-1] + len(red[k]))
    T['red_off'] = i32(red_off)
    T['red_rx'] = i32([p for lst in red for p, _ in lst] or [0])
    T['red_nu'] = f64([nu for lst in red for _, nu in lst] or [0.0])

    # ---------------- sparse part: entry (k, j) <- sum coef * raw[src]
    contrib: Dict[tuple, list] = {}

    def add(k, j, src, coef):
        if coef != 0.0 and src >= 0:
            contrib.setdefault((k, j), []).append((src, float(coef)))

    for i, rx in enumerate(reacs):
        part = [(k, rx.net_nu(k)) for k in sorted(set(rx.reac + rx.prod)) if rx.net_nu(k) != 0]
        p = pos_of[i]
        sl = slots[p]
        for k, nu in part:
            for a in range(2 * MAXS):
                if raw_of_slot[i][a] >= 0:
                    add(k, int(sl[a]), raw_of_slot[i][a], nu)                  # cj:410-448
            if has_pmt(rx):
                if eff_case(rx):
                    for s, src in raw_of_eff[i].items():
                        add(k, s, src, nu)                                     # cj:379-400
                elif rx.pdep_sp is not None and rx.pdep_sp != last:
                    add(k, rx.pdep_sp, raw_of_pmt[i], nu)                      # cj:401-404
            elif i in stale_src and stale_src[i] is not None:
                add(k, 0, raw_of_pmt[stale_src[i]], nu)
    # energy-equation row: sum_k h_k W_k S_kj = sum over raw values of column j of
    # dH_i * raw  (dH_i = sum_k nu_ki h_k W_k, evaluated per state by the kernel), so its
    # contributions carry the *reaction* whose dH multiplies the raw value
    tcontrib: Dict[int, list] = {}
    for i, rx in enumerate(reacs):
        p = pos_of[i]
        sl = slots[p]
        if not any(rx.net_nu(k) != 0 for k in set(rx.reac + rx.prod)):
            continue
        for a in range(2 * MAXS):
            if raw_of_slot[i][a] >= 0:
                tcontrib.setdefault(int(sl[a]), []).append((raw_of_slot[i][a], p))
        if has_pmt(rx):
            if eff_case(rx):
                for s, src in raw_of_eff[i].items():
                    tcontrib.setdefault(s, []).append((src, p))
            elif rx.pdep_sp is not None and rx.pdep_sp != last:
                tcontrib.setdefault(rx.pdep_sp, []).append((raw_of_pmt[i], p))
        elif i in stale_src and stale_src[i] is not None:
            tcontrib.setdefault(0, []).append((raw_of_pmt[stale_src[i]], p))

    # ---------------- raw rows written by each reaction
    # rx_dst[p][0..5]: raw slot written by concentration slot a (0xFFFF: none);
    # [6]: first collider-list raw slot, [7]: raw slot of pres_mod_temp (0xFFFF: none)
    NONE = 0xFFFF
    rx_dst = np.full((nr, 8), NONE, dtype=np.uint16)
    for p, i in enumerate(order):
        for a in range(2 * MAXS):
            if raw_of_slot[i][a] >= 0:
                rx_dst[p, a] = raw_of_slot[i][a]
        if raw_of_eff[i]:
            rx_dst[p, 6] = min(raw_of_eff[i].values())
        if raw_of_pmt[i] >= 0:
            rx_dst[p, 7] = raw_of_pmt[i]
    T['rx_dst'] = rx_dst.ravel()

    # ---------------- schedule of the Jacobian kernel (plan.py)
    # states per block, block size, and where the working set lives: shared memory when at least
    # plan.SMEM_MIN_GS states fit there, else a per-block scratch area in global memory
    nt = threads or plan.DEFAULT_THREADS
    if not gs:
        gs_smem = 0 if ws_global else plan.choose_gs(nsp, nr, nraw, nt // 32)
        if ws_global is False and not gs_smem:
            raise UnsupportedMechanism('working set of one state pair exceeds shared memory')
        if ws_global is False or (ws_global is None and gs_smem >= plan.SMEM_MIN_GS):
            gs = gs_smem
        else:
            gs, nt_g = plan.wsg_shape(nsp)
            ws_global = True
            if not threads:
                nt = nt_g
    if ws_global and nt > 384:
        raise UnsupportedMechanism('plans with the working set in global memory take at most 384 threads')
    if ws_global is None and plan.layout(nsp, nr, nraw, gs, nt // 32)['total'] * 8 > plan.SMEM_LIMIT:
        ws_global = True               # a requested gs that does not fit in shared memory
    if ws_global and gs not in (2, 8, 16, 32):
        raise UnsupportedMechanism('plans with the working set in global memory hold 2, 8, 16 or 32 states per block')
    kinds, n_eff = [], []
    for p, i in enumerate(order):
        rx = reacs[i]
        kinds.append('sri' if rx.sri else 'troe' if rx.troe else 'lind' if rx.pdep else
                     'thd' if rx.thd_body else 'plog' if (rx.plog or rx.cheb) else 'plain')
        n_eff.append(sum(1 for s, a in rx.thd_body_eff if a != 1.0) if (rx.thd_body or rx.pdep) else 0)
    T.update(plan.build_plan(nsp, nr, nraw, first_pm, kinds, [bool(reacs[i].rev) for i in order],
                             [bool(slots[p, 2] != nsp or slots[p, 5] != nsp) for p in range(nr)],
                             n_eff, red, {kj: v for kj, v in contrib.items() if kj[0] != last},
                             tcontrib, T['sp_w'], T['sp_iw'], T['sp_mwf'], gs, nt, 1 if ws_global else 0))
    # the 64-byte reaction record of the Jacobian kernel: lnA, b, Ta, sum(nu) ln(PA/RU); flags;
    # six species slots and eight raw destinations packed two per int
    rec5 = np.zeros((nr, 16), dtype=np.int32)
    rec5[:, :9] = rec[:, :9]
    # species slots as (byte offset of the species' even-slot base) / 16; the slot pair of odd
    # species is swapped (plan.py, sp_even) so that rows of different species use different banks
    gs_ = int(T['p5_cfg'][0])
    spf = (slots * (plan.SP_SLOTS * gs_ * 8) + (slots & 1) * (gs_ * 8)) // 16
    if spf.max() > 0xFFFF:
        raise UnsupportedMechanism('too many species for 16-bit species-row offsets')
    for a in range(3):
        rec5[:, 9 + a] = spf[:, 2 * a] | (spf[:, 2 * a + 1] << 16)
    dst5 = np.where(rx_dst == NONE, nraw + 1, rx_dst).astype(np.int32)     # none -> scratch raw row
    for a in range(4):
        rec5[:, 12 + a] = dst5[:, 2 * a] | (dst5[:, 2 * a + 1] << 16)
    for p in range(nr):
        if any(int(sl) == last for sl in slots[p]):
            rec5[p, 8] |= F_HAS_LAST
    T['p5_rx'] = rec5.ravel()
    # positions of kernel-order reaction p in the reference's fwd / rev / pres_mod arrays
    T['p5_rxout'] = i32([[order[p], rev_idx[p], pm_idx[p], 0] for p in range(nr)]).ravel()
    # collider lists of the Jacobian kernel: per pressure-modified reaction a list padded to a
    # multiple of four records {alpha - 1 (double), byte offset of the collider's species row,
    # raw row that receives pres_mod_temp * (alpha - 1)}; padding: alpha - 1 = 0, empty species
    # slot, scratch raw row.  Four records are fetched at a time.
    spb_ = plan.SP_SLOTS * gs_ * 8
    sp_off = lambda k: k * spb_ + (k & 1) * gs_ * 8
    eff4, eff4_off = [], [0]
    for p in range(first_pm, nr):
        i = order[p]
        m = p - first_pm
        recs = []
        for e in range(eff_off[m], eff_off[m + 1]):
            sp_, am1 = eff_sp[e], eff_am1[e]
            dst = raw_of_eff[i].get(sp_, nraw + 1) if eff_case(reacs[i]) else nraw + 1
            lohi = np.array([am1], dtype=np.float64).view(np.int32)
            recs.append([int(lohi[0]), int(lohi[1]), sp_off(sp_), dst])
        while len(recs) % 4:
            recs.append([0, 0, sp_off(nsp), nraw + 1])
        for r_ in recs:
            eff4 += r_
        eff4_off.append(len(eff4) // 4)
    T['p5_eff_off'] = i32(eff4_off if npm else [0, 0])
    T['p5_eff'] = i32(eff4 + [0, 0, sp_off(nsp), nraw + 1] * 4)

    # ---------------- record streams of the Jacobian kernel k_jac6 (plan6.py): plans whose working
    # set lives in shared memory
    p_c0 = next((p for p in range(nr) if p >= first_pm or last in (reacs[order[p]].reac + reacs[order[p]].prod)), nr)
    if streams and not ws_global and gs_ in plan6.GS6 and plan6.fits(nsp, nr, nr - p_c0, nraw, gs_, nt // 32):
        corr_rx = [False] * nr
        for p in range(p_c0, nr):
            par = pm_par[p - first_pm] if p >= first_pm else None
            corr_rx[p] = bool(any(int(sl) == last for sl in slots[p]) or (par is not None and par[4] != par[5]))
        rec6 = rec5.copy()
        spf6 = (slots * (plan6.SP_SLOTS * gs_ * 8) + (slots & 1) * (gs_ * 8)) // 16
        for a in range(3):
            rec6[:, 9 + a] = spf6[:, 2 * a] | (spf6[:, 2 * a + 1] << 16)
        rec6[p_c0:, 8] |= plan6.F_CORR
        rec6[:, 15] = np.arange(nr, dtype=np.int32) | (dst5[:, 7] << 16)
        T.update(plan6.build_plan6(nsp, nr, nraw, first_pm, p_c0, kinds, [bool(reacs[i].rev) for i in order],
                                   [bool(slots[p, 2] != nsp or slots[p, 5] != nsp) for p in range(nr)],
                                   n_eff, rec6, red, corr_rx,
                                   {kj: v for kj, v in contrib.items() if kj[0] != last},
                                   tcontrib, T['sp_w'], T['sp_iw'], T['sp_mwf'], gs_, nt))
        spb6 = plan6.SP_SLOTS * gs_ * 8
        eff6 = np.asarray(T['p5_eff']).reshape(-1, 4).copy()
        ksp = eff6[:, 2] // spb_                       # species of each collider record
        eff6[:, 2] = ksp * spb6 + (ksp & 1) * gs_ * 8
        T['p6_eff'] = eff6.ravel()

    T['sp_fwd_map'] = i32(mech.fwd_spec_map)            # internal position -> index in the mechanism file (apply_mask)
    T['meta'] = i32([SCHEMA_VERSION, plan.PLAN_VERSION, plan6.PLAN_VERSION, 1 if conv else 0])
    T['cst'] = f64([q('{:.8e}', RU), ln_pa_ru, 0.0, 0.0])
    T['dims'] = i32([nsp, nr, len(rev_reacs), npm, nraw, len(plog_par), len(cheb_par), 0,
                     first_pm, npm, red_off[-1], max(len(l) for l in red), 0, 0, 0, 0])
    return T
