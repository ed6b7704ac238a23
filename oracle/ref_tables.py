"""TEST INFRASTRUCTURE -- tables for the CPU restatement oracle (oracle/pyjac_oracle.c).

Every *literal constant* the reference generator would print into its emitted C is
computed here with the same Python arithmetic, in the same order, and quantised with
the same format string, so that the C restatement -- which performs the run-time
arithmetic in the emitted code's order -- reproduces the reference bit for bit where
libm agrees.  Citations: rs = pyjac/core/rate_subs.py, cj = pyjac/core/create_jacobian.py.

Scope: elementary, third-body, fall-off (Lindemann / Troe / SRI, LOW or HIGH), PLOG and
Chebyshev (at least 3 x 2 coefficients: cj:1583-1585 indexes dot_prod[2]) reactions with integer stoichiometric coefficients and positive pre-exponentials.  PLOG is
restated for the inputs the generator emits valid C for: every two consecutive pressures
must have different activation energies (cj:1759-1770 builds an unbalanced expression
otherwise).
"""
from __future__ import annotations

import math
from typing import Dict, List

import numpy as np

from pyjac_b200.chem import PA, RU
from pyjac_b200.mechanism import Mechanism

# flag bits shared with pyjac_oracle.c
F_REV, F_THD, F_PDEP, F_LOW, F_TROE, F_SRI, F_EFF = 1, 2, 4, 8, 16, 32, 64
F_PDEPSP_TRUTHY, F_NO_T, F_TROE_T2, F_SRI5, F_SRI5_DT = 128, 256, 512, 1024, 2048
F_PMT, F_PMT_IN_JTEMP, F_HAS_DBDT, F_KCJ_PREF = 4096, 8192, 16384, 32768
F_PLOG, F_CHEB = 65536, 131072

UNROLL = 40   # CParams.Jacob_Unroll: conc_temp collapsing restarts every 40 reactions


def q(fmt: str, x: float) -> float:
    """Value of x after a round trip through the generator's format string."""
    return float(fmt.format(x))


def is_int(v) -> bool:
    return float(v).is_integer()


def arrhenius_form(A: float, b: float, E: float):
    """rs:27-146: returns [form, c0, b, E];  A > 0 (c0 = log A):  kf =
    0: c0 (= A) | 1: exp(c0 + b*logT) | 2: exp(c0 - (E/T)) | 3: exp(c0 + b*logT - (E/T));
    A < 0 (rs:108-141, c0 = A):  4: c0 * T^b for an integer-valued b | 5: c0 * exp(b*logT) |
    6: c0 * exp(-(E/T)) | 7: c0 * exp(b*logT - (E/T))."""
    if A < 0:
        if isinstance(b, int):
            raise NotImplementedError('integer-typed temperature exponent')
        if not E:
            if not b:
                return [0.0, float(str(A)), 0.0, 0.0]
            if is_int(b):
                return [4.0, float(str(A)), float(int(b)), 0.0]
            return [5.0, q('{:.16e}', A), float(str(b)), 0.0]
        if not b:
            return [6.0, q('{:.16e}', A), 0.0, q('{:.16e}', E)]
        return [7.0, q('{:.16e}', A), float(str(b)), q('{:.16e}', E)]
    if not A > 0:
        raise NotImplementedError('zero pre-exponential factor')
    if isinstance(b, int):
        raise NotImplementedError('integer-typed temperature exponent')
    logA = math.log(A)
    if not E:
        if not b:
            return [0.0, float(str(A)), 0.0, 0.0]
        return [1.0, q('{:.16e}', logA), float(str(b)), 0.0]
    if not b:
        return [2.0, q('{:.16e}', logA), 0.0, q('{:.16e}', E)]
    return [3.0, q('{:.16e}', logA), float(str(b)), q('{:.16e}', E)]


def _nasa_arrays(sp, nu):
    """rs:540-561 / cj:522-538."""
    def one(a):
        arr = [nu, a[6], a[0], a[0] - 1.0, a[1] / 2.0, a[2] / 6.0, a[3] / 12.0, a[4] / 20.0, a[5]]
        return [x * arr[0] for x in [arr[1] - arr[2]] + arr[3:]]
    return one(sp.lo), one(sp.hi)


def _acc(coeffs, tmid, lo, hi):
    if tmid not in coeffs:
        coeffs[tmid] = lo, hi
    else:
        coeffs[tmid] = ([lo[i] + coeffs[tmid][0][i] for i in range(len(lo))],
                        [hi[i] + coeffs[tmid][1][i] for i in range(len(hi))])


def kc_rates(specs, rxn):
    """Per-T_mid pre-summed coefficients as eval_rxn_rates prints them (rs:668-809)."""
    coeffs = {}
    sum_nu = 0
    for isp, psp in enumerate(rxn.prod):
        if psp in rxn.reac:
            nu = rxn.prod_nu[isp] - rxn.reac_nu[rxn.reac.index(psp)]
        else:
            nu = rxn.prod_nu[isp]
        if nu == 0:
            continue
        sum_nu += nu
        lo, hi = _nasa_arrays(specs[psp], nu * 1.0)
        _acc(coeffs, specs[psp].Trange[1], lo, hi)
    for isp, rsp in enumerate(rxn.reac):
        if rsp in rxn.prod:
            continue
        nu = rxn.reac_nu[isp]
        sum_nu -= nu
        lo, hi = _nasa_arrays(specs[rsp], nu * -1.0)
        _acc(coeffs, specs[rsp].Trange[1], lo, hi)
    return coeffs, (PA / RU) ** sum_nu


def kc_jac(specs, rxn):
    """Same constants as eval_jacob prints them (cj:511-547, 614-618)."""
    coeffs = {}
    sum_nu = 0
    for isp in set(rxn.reac + rxn.prod):
        nu = rxn.net_nu(isp)
        if nu == 0:
            continue
        sum_nu += nu
        lo, hi = _nasa_arrays(specs[isp], nu)
        _acc(coeffs, specs[isp].Trange[1], lo, hi)
    pref = (PA / RU) ** sum_nu if sum_nu != 0 else None
    return coeffs, pref


class _CSR:
    def __init__(self):
        self.off = [0]
        self.cols: Dict[str, list] = {}

    def add(self, **rows):
        n = None
        for k, v in rows.items():
            self.cols.setdefault(k, []).extend(v)
            n = len(v)
        self.off.append(self.off[-1] + (n or 0))


def build(mech: Mechanism) -> Dict[str, np.ndarray]:
    specs, reacs = mech.specs, mech.reacs
    nsp, nr = len(specs), len(reacs)
    last = nsp - 1
    rev_reacs = mech.rev_reacs
    pdep_reacs = mech.pdep_reacs
    T: Dict[str, np.ndarray] = {}
    f64 = lambda x: np.asarray(x, dtype=np.float64)
    i32 = lambda x: np.asarray(x, dtype=np.int32)

    for rx in reacs:
        if rx.cheb and (rx.pdep or rx.thd_body or rx.cheb_n_temp < 3 or rx.cheb_n_pres < 2):
            raise NotImplementedError('Chebyshev reaction with a third body / fewer than 3 x 2 coefficients')
        if rx.plog and (rx.pdep or rx.thd_body or len(rx.plog_par) < 2):
            raise NotImplementedError('PLOG reaction with a third body / fewer than two pressures')
        if not all(is_int(v) for v in rx.reac_nu + rx.prod_nu):
            raise NotImplementedError('non-integer stoichiometric coefficients')

    T['dims'] = i32([nsp, nr, len(rev_reacs), len(pdep_reacs)])
    T['ru8'] = f64([q('{:.8e}', RU)])

    # ---------------- species
    T['sp_mw'] = f64([sp.mw for sp in specs])
    T['sp_mw_inv'] = f64([q('{:.16e}', 1.0 / sp.mw) for sp in specs])          # rs:1678,1698
    T['sp_mw8'] = f64([q('{:.8e}', sp.mw) for sp in specs])                    # cj:1885,3230
    T['sp_ru_mw'] = f64([q('{:.16e}', RU / sp.mw) for sp in specs])            # rs:1834,2049
    T['sp_tmid'] = f64([sp.Trange[1] for sp in specs])
    T['sp_lo'] = f64([sp.lo for sp in specs]).ravel()
    T['sp_hi'] = f64([sp.hi for sp in specs]).ravel()
    # eval_h (rs:1834-1857): a5, a0, a1/2, a2/3, a3/4, a4/5
    hcoef = lambda a: [a[5], a[0], a[1] / 2.0, a[2] / 3.0, a[3] / 4.0, a[4] / 5.0]
    T['sp_h_lo'] = f64([hcoef(sp.lo) for sp in specs]).ravel()
    T['sp_h_hi'] = f64([hcoef(sp.hi) for sp in specs]).ravel()
    # dBdT (cj:829-858): a0-1, a5, a1/2, a2/3, a3/4, a4/5
    dcoef = lambda a: [a[0] - 1.0, a[5], a[1] / 2.0, a[2] / 3.0, a[3] / 4.0, a[4] / 5.0]
    T['sp_db_lo'] = f64([dcoef(sp.lo) for sp in specs]).ravel()
    T['sp_db_hi'] = f64([dcoef(sp.hi) for sp in specs]).ravel()
    # dcp/dT (cj:1347-1351): a1, 2a2, 3a3, 4a4
    ccoef = lambda a: [a[1], 2.0 * a[2], 3.0 * a[3], 4.0 * a[4]]
    T['sp_dcp_lo'] = f64([ccoef(sp.lo) for sp in specs]).ravel()
    T['sp_dcp_hi'] = f64([ccoef(sp.hi) for sp in specs]).ravel()
    # T_mid buckets of write_dcp_dt (cj:1314-1322): sorted T_mid, species sorted
    buckets: Dict[float, List[int]] = {}
    for isp, sp in enumerate(specs):
        buckets.setdefault(sp.Trange[1], []).append(isp)
    csr = _CSR()
    tm = []
    for t_mid in sorted(buckets):
        tm.append(t_mid)
        csr.add(sp=sorted(buckets[t_mid]))
    T['dcp_off'], T['dcp_sp'], T['dcp_tmid'] = i32(csr.off), i32(csr.cols['sp']), f64(tm)

    # species with a non-zero rate (rs:1425-1527 'seen') and dBdT flags (cj:800-817)
    seen = [False] * nsp
    dbdt_flag = [False] * nsp
    for rx in reacs:
        for k in set(rx.reac + rx.prod):
            if rx.net_nu(k) != 0:
                seen[k] = True
        if rx.rev:
            for k in rx.reac + rx.prod:
                dbdt_flag[k] = True
    T['sp_seen'] = i32(seen)
    T['sp_dbdt_flag'] = i32(dbdt_flag)
    mwN = specs[last].mw
    T['j_mwfrac'] = f64([q('{:.16e}', sp.mw / mwN) for sp in specs[:-1]])          # cj:467
    T['j_cj'] = f64([q('{:.16e}', 1. - sp.mw / mwN) for sp in specs[:-1]])        # cj:378

    # ---------------- reactions
    flags = [0] * nr
    pdep_sp = [-1] * nr
    rev_idx = [-1] * nr
    pm_idx = [-1] * nr
    reac, prod, net, dbl, eff = _CSR(), _CSR(), _CSR(), _CSR(), _CSR()
    kcr, kcj, prl = _CSR(), _CSR(), _CSR()
    kcr_pref, kcj_pref = [], []
    arr_main = []
    arr_k0 = np.zeros((nr, 4))
    arr_kinf = np.zeros((nr, 4))
    arr_ratio = np.zeros((nr, 4))
    troe_pm = np.zeros((nr, 8))
    troe_j = np.zeros((nr, 8))
    sri_pm = np.zeros((nr, 8))
    sri_j = np.zeros((nr, 12))
    dt = np.zeros((nr, 8))
    pdt = np.zeros((nr, 4))
    drdy = np.zeros((nr, 2))
    pr_mode = [0] * nr
    alpha_mode = np.zeros((nr, max(nsp - 1, 1)), dtype=np.int32)
    alpha_val = np.zeros((nr, max(nsp - 1, 1)))

    plog_off, plog_p4, plog_arr, plog_lp, plog_dlp, plog_dt, plog_mid = [0], [], [], [], [], [], []
    cheb_off, cheb_dim, cheb_red, cheb_c8, cheb_c16 = [0], [], [], [], []

    last_conc_temp = None
    do_unroll = nr > UNROLL
    for i, rx in enumerate(reacs):
        fl = 0
        if rx.plog:
            # rs:598-632 / cj:293-327 (rate constant), cj:1687-1850 (temperature derivative)
            fl |= F_PLOG
            pp = rx.plog_par
            for e, (p1, A1, b1, E1) in enumerate(pp):
                plog_p4.append(q('{:.4e}', p1))
                plog_arr.append(arrhenius_form(A1, b1, E1))
                plog_lp.append(q('{:.16e}', math.log(p1)))
                b_on, E_on = abs(b1) > 1.0e-90, abs(E1) > 1.0e-90
                if not rx.rev and not b_on and not E_on and sum(rx.reac_nu) == 1.0:
                    raise NotImplementedError('PLOG end range without a temperature derivative')
                plog_dt.append([(1 if b_on else 0) | (2 if E_on else 0), q('{:.16e}', b1), q('{:.16e}', E1)])
                if e + 1 < len(pp):
                    p2, A2, b2, E2 = pp[e + 1]
                    if A2 / A1 < 0 or E2 - E1 == 0.0 or p1 == p2:
                        raise NotImplementedError('PLOG pair the generator emits no valid C for')
                    plog_dlp.append(q('{:.16e}', math.log(p2) - math.log(p1)))
                    plog_mid.append([1.0 if b1 != 0.0 else 0.0, q('{:.16e}', b1),
                                     1.0 if E1 != 0.0 else 0.0, q('{:.16e}', E1),
                                     1.0 if b2 - b1 != 0.0 else 0.0, q('{:.16e}', b2 - b1),
                                     q('{:.16e}', E2 - E1), 1.0 if p1 != 1.0 else 0.0])
                else:
                    plog_dlp.append(0.0)
                    plog_mid.append([0.0] * 8)
        plog_off.append(len(plog_p4))
        if rx.cheb:
            # rs:149-251 ({:.8e}) and cj:1532-1684 ({:.16e}, coefficients times their row index)
            fl |= F_CHEB
            tsum = 1.0 / rx.cheb_tlim[0] + 1.0 / rx.cheb_tlim[1]
            tsub = 1.0 / rx.cheb_tlim[1] - 1.0 / rx.cheb_tlim[0]
            psum = math.log10(rx.cheb_plim[0]) + math.log10(rx.cheb_plim[1])
            psub = math.log10(rx.cheb_plim[1]) - math.log10(rx.cheb_plim[0])
            cheb_red.append([q('{:.8e}', tsum), q('{:.8e}', tsub), q('{:.8e}', psum), q('{:.8e}', psub),
                             q('{:.16e}', tsum), q('{:.16e}', tsub), q('{:.16e}', psum), q('{:.16e}', psub),
                             q('{:.16e}', -2.0 * math.log(10) / tsub), 0.0, 0.0, 0.0])
            cheb_dim.append([rx.cheb_n_temp, rx.cheb_n_pres])
            for r_ in range(rx.cheb_n_temp):
                for c_ in range(rx.cheb_n_pres):
                    cheb_c8.append(q('{:.8e}', rx.cheb_par[r_][c_]))
                    cheb_c16.append(q('{:.16e}', r_ * rx.cheb_par[r_][c_]))
        else:
            cheb_red.append([0.0] * 12)
            cheb_dim.append([0, 0])
        cheb_off.append(len(cheb_c8))
        if rx.rev:
            fl |= F_REV
            rev_idx[i] = rev_reacs.index(i)
        if rx.thd_body:
            fl |= F_THD
        if rx.pdep:
            fl |= F_PDEP
            if rx.low:
                fl |= F_LOW
            elif not rx.high:
                raise NotImplementedError('fall-off reaction without LOW or HIGH')
        if rx.thd_body or rx.pdep:
            pm_idx[i] = pdep_reacs.index(i)
        if rx.troe:
            fl |= F_TROE
        if rx.sri:
            fl |= F_SRI
        if rx.thd_body_eff:
            fl |= F_EFF
        if rx.pdep_sp is not None:
            pdep_sp[i] = rx.pdep_sp
        if rx.pdep_sp:
            fl |= F_PDEPSP_TRUTHY

        reac.add(sp=rx.reac, nu=[int(v) for v in rx.reac_nu])
        prod.add(sp=rx.prod, nu=[int(v) for v in rx.prod_nu])
        order = [k for k in set(rx.reac + rx.prod) if rx.net_nu(k) != 0]
        net.add(sp=order, nu=[int(rx.net_nu(k)) for k in order])
        eff.add(sp=[s for s, _ in rx.thd_body_eff], alpha=[a for _, a in rx.thd_body_eff])

        arr_main.append(arrhenius_form(rx.A, rx.b, rx.E))

        # Kc constants (both printings)
        if rx.rev:
            c, pref = kc_rates(specs, rx)
            kcr.add(tmid=list(c), lo=[v for t in c for v in c[t][0]], hi=[v for t in c for v in c[t][1]])
            kcr.off[-1] = kcr.off[-2] + len(c)
            kcr_pref.append(q('{:.16e}', pref))
            c, pref = kc_jac(specs, rx)
            kcj.add(tmid=list(c), lo=[v for t in c for v in c[t][0]], hi=[v for t in c for v in c[t][1]])
            kcj.off[-1] = kcj.off[-2] + len(c)
            if pref is not None:
                fl |= F_KCJ_PREF
                kcj_pref.append(q('{:.16e}', pref))
            else:
                kcj_pref.append(1.0)
        else:
            kcr.add(tmid=[], lo=[], hi=[])
            kcj.add(tmid=[], lo=[], hi=[])
            kcr_pref.append(1.0)
            kcj_pref.append(1.0)

        # dBdT sum in the order get_db_dt prints it (cj:888-950)
        dsp, dnu = [], []
        for k in rx.prod:
            nu = rx.net_nu(k) if k in rx.reac else rx.prod_nu[rx.prod.index(k)]
            if nu == 0:
                continue
            dsp.append(k)
            dnu.append(int(nu))
        for k in rx.reac:
            if k in rx.prod:
                continue
            dsp.append(k)
            dnu.append(-int(rx.reac_nu[rx.reac.index(k)]))
        dbl.add(sp=dsp, nu=dnu)
        if dsp:
            fl |= F_HAS_DBDT

        # temperature-derivative pieces (cj:724-758, 1398-1529)
        b_on, E_on = abs(rx.b) > 1.0e-90, abs(rx.E) > 1.0e-90
        dk_form = (1 if b_on else 0) | (2 if E_on else 0)
        rnu, pnu = sum(rx.reac_nu), sum(rx.prod_nu)
        dt[i] = [dk_form, q('{:.16e}', rx.b), q('{:.16e}', rx.E),
                 float(str(1. - float(rnu))), 1.0 if rnu != 1.0 else 0.0,
                 float(str(1. - float(pnu))), 1.0 if pnu != 1.0 else 0.0, 0.0]
        if not rx.rev and not dk_form and rnu == 1.0 and not rx.plog and not rx.cheb:
            fl |= F_NO_T

        if rx.pdep:
            k0 = rx.low if rx.low else [rx.A, rx.b, rx.E]
            kinf = [rx.A, rx.b, rx.E] if rx.low else rx.high
            arr_k0[i] = arrhenius_form(*k0)
            arr_kinf[i] = arrhenius_form(*kinf)
            beta_0minf, E_0minf = k0[1] - kinf[1], k0[2] - kinf[2]      # cj:641-655
            arr_ratio[i] = arrhenius_form(k0[0] / kinf[0], beta_0minf, E_0minf)
            pdt[i] = [q('{:.4e}', beta_0minf), q('{:.16e}', beta_0minf),
                      q('{:.16e}', E_0minf), 0.0]
            if rx.troe:
                a, T3, T1 = rx.troe_par[:3]
                T2 = rx.troe_par[3] if len(rx.troe_par) == 4 else 0.0
                if len(rx.troe_par) == 4 and T2 != 0.0:
                    fl |= F_TROE_T2
                # rs:1189-1209  ({:.8e})
                troe_pm[i] = [q('{:.8e}', 1.0 - a), q('{:.8e}', abs(T3)), 1.0 if T3 > 0.0 else -1.0,
                              q('{:.8e}', a), q('{:.8e}', abs(T1)), 1.0 if T1 > 0.0 else -1.0,
                              q('{:.8e}', abs(T2)), 1.0 if T2 > 0.0 else -1.0]
                # cj:1083-1090, 1262-1282  ({:.16e})
                troe_j[i] = [q('{:.16e}', 1.0 - a), q('{:.16e}', -T3), q('{:.16e}', a),
                             q('{:.16e}', -T1), q('{:.16e}', -T2),
                             q('{:.16e}', -(1.0 - a) / T3), q('{:.16e}', a / T1),
                             q('{:.16e}', T2)]
            elif rx.sri:
                sp_ = rx.sri_par
                five = len(sp_) == 5
                if five and sp_[3] != 1.0 and sp_[4] != 0.0:
                    fl |= F_SRI5
                if five and sp_[4] != 0.0:
                    fl |= F_SRI5_DT
                d, e = (sp_[3], sp_[4]) if five else (1.0, 0.0)
                # rs:1239-1255 / cj:250-266  ({:.6}, {:.8e})
                sri_pm[i] = [q('{:.6}', sp_[0]), q('{:.6}', abs(sp_[1])), 1.0 if sp_[1] > 0.0 else -1.0,
                             q('{:.6}', abs(sp_[2])), 1.0 if sp_[2] > 0.0 else -1.0,
                             q('{:.8e}', d), q('{:.6}', e), 0.0]
                # cj:173-180 ({:.4}) and cj:1215-1235 ({:.16}, {:.16e})
                sri_j[i] = [q('{:.4}', sp_[0]), q('{:.4}', -sp_[1]), q('{:.4}', -sp_[2]),
                            q('{:.16}', sp_[0] * sp_[1]), q('{:.16}', -sp_[1]),
                            q('{:.16e}', 1.0 / sp_[2]), q('{:.16}', -sp_[2]),
                            q('{:.16}', sp_[0]), q('{:.16}', e), 0.0, 0.0, 0.0]

        # conc_temp recipe of write_pr, including the "collapsing" against the previous
        # fall-off reaction (cj:986-1052) and its reset at unroll boundaries (cj:2651-2653)
        if do_unroll and i % UNROLL == 0:
            last_conc_temp = None
        if rx.pdep:
            if rx.pdep_sp is not None:
                pr_mode[i] = 1
                prl.add(sp=[rx.pdep_sp], coef=[0.0])
                log = None
            elif not rx.thd_body_eff:
                pr_mode[i] = 2
                prl.add(sp=[], coef=[])
                log = None
            else:
                log = [(s, a - 1.0) for s, a in rx.thd_body_eff if a != 1.0]
                use, mode = log, 3
                if last_conc_temp is not None:
                    new = []
                    for s, a in log:
                        m = next((x for x in last_conc_temp if x[0] == s), None)
                        c = a - m[1] if m is not None else a
                        if c != 0.0:
                            new.append((s, c))
                    for s, a in last_conc_temp:
                        if next((x for x in log if x[0] == s), None) is None:
                            new.append((s, -a))
                    if len(new) < len(log):
                        use, mode = new, 4
                    if not len(use):
                        mode = 5
                pr_mode[i] = mode
                # coefficients go through str()/abs() in the generator: exact
                prl.add(sp=[s for s, _ in use], coef=[c for _, c in use])
            last_conc_temp = log
        else:
            prl.add(sp=[], coef=[])

        # species-independent dR/dY pieces (cj:153-230)
        if (rx.pdep or rx.thd_body) and (rx.thd_body_eff or rx.pdep_sp):
            fl |= F_PMT
        if rx.pdep and (rx.pdep_sp or rx.thd_body_eff):
            fl |= F_PMT_IN_JTEMP
        n_r, n_p = 0, 0
        if rx.thd_body_eff and not rx.pdep:
            n_r = 1
            if rx.rev:
                n_p = 1
        n_r += sum(rx.reac_nu)
        if rx.rev:
            n_p += sum(rx.prod_nu)
        drdy[i] = [float(n_r), float(n_p)]

        # alpha_ij terms (cj:379-400)
        for j in range(nsp - 1):
            mw_frac = specs[j].mw / mwN
            if ((rx.pdep and rx.pdep_sp is None) or rx.thd_body) and rx.thd_body_eff:
                aij = next((a for s, a in rx.thd_body_eff if s == j), 1.0)
                aiN = next((a for s, a in rx.thd_body_eff if s == last), 1.0)
                if aiN != 0:
                    aij -= aiN * mw_frac
                if aij != 0:
                    if aij == 1:
                        alpha_mode[i, j] = 1
                    elif aij == -1:
                        alpha_mode[i, j] = 2
                    else:
                        alpha_mode[i, j] = 3
                        alpha_val[i, j] = q('{:.16e}', aij)
            elif rx.pdep_sp == j or rx.pdep_sp == last:
                if rx.pdep_sp == j:
                    alpha_mode[i, j] = 1
                else:
                    alpha_mode[i, j] = 4
                    alpha_val[i, j] = q('{:.16e}', specs[j].mw / specs[rx.pdep_sp].mw)
        flags[i] = fl

    T['rx_flags'] = i32(flags)
    T['rx_pdep_sp'] = i32(pdep_sp)
    T['rx_rev_idx'] = i32(rev_idx)
    T['rx_pm_idx'] = i32(pm_idx)
    for nm, c in (('reac', reac), ('prod', prod), ('net', net), ('db', dbl)):
        T['rx_%s_off' % nm] = i32(c.off)
        T['rx_%s_sp' % nm] = i32(c.cols.get('sp', []))
        T['rx_%s_nu' % nm] = i32(c.cols.get('nu', []))
    T['rx_eff_off'] = i32(eff.off)
    T['rx_eff_sp'] = i32(eff.cols.get('sp', []))
    T['rx_eff_alpha'] = f64(eff.cols.get('alpha', []))
    T['rx_pr_mode'] = i32(pr_mode)
    T['rx_pr_off'] = i32(prl.off)
    T['rx_pr_sp'] = i32(prl.cols.get('sp', []))
    T['rx_pr_coef'] = f64(prl.cols.get('coef', []))
    for nm, c, pref in (('kcr', kcr, kcr_pref), ('kcj', kcj, kcj_pref)):
        T[nm + '_off'] = i32(c.off)
        T[nm + '_tmid'] = f64(c.cols.get('tmid', []))
        T[nm + '_lo'] = f64(c.cols.get('lo', []))
        T[nm + '_hi'] = f64(c.cols.get('hi', []))
        T[nm + '_pref'] = f64(pref)
    T['arr_main'] = f64(arr_main).ravel()
    T['arr_k0'] = arr_k0.ravel()
    T['arr_kinf'] = arr_kinf.ravel()
    T['arr_ratio'] = arr_ratio.ravel()
    T['troe_pm'] = troe_pm.ravel()
    T['troe_j'] = troe_j.ravel()
    T['sri_pm'] = sri_pm.ravel()
    T['sri_j'] = sri_j.ravel()
    T['cheb_off'] = i32(cheb_off)
    T['cheb_dim'] = i32(cheb_dim).ravel()
    T['cheb_red'] = f64(cheb_red).ravel()
    T['cheb_c8'] = f64(cheb_c8 or [0.0])
    T['cheb_c16'] = f64(cheb_c16 or [0.0])
    T['plog_off'] = i32(plog_off)
    T['plog_p4'] = f64(plog_p4 or [0.0])
    T['plog_arr'] = f64(plog_arr or [[0.0] * 4]).ravel()
    T['plog_lp'] = f64(plog_lp or [0.0])
    T['plog_dlp'] = f64(plog_dlp or [0.0])
    T['plog_dt'] = f64(plog_dt or [[0.0] * 3]).ravel()
    T['plog_mid'] = f64(plog_mid or [[0.0] * 8]).ravel()
    T['rx_dt'] = dt.ravel()
    T['rx_pdt'] = pdt.ravel()
    T['rx_drdy'] = drdy.ravel()
    T['alpha_mode'] = alpha_mode.ravel()
    T['alpha_val'] = alpha_val.ravel()
    # generator constants used at run time
    ln10 = math.log(10.0)
    T['consts'] = f64([
        q('{:.16}', 1.0 / ln10), q('{:.16}', 0.14 / ln10),            # cj:167-169  [0,1]
        q('{:.16}', 2.0 / ln10),                                       # cj:175,1224 [2]
        q('{:.16e}', 0.67 / ln10), q('{:.16e}', 1.1762 / ln10),       # cj:1265-1267 [3,4]
        q('{:.16e}', 1.0 / ln10), q('{:.16e}', 0.14 / ln10),          # cj:1286-1288 [5,6]
    ])
    return T
