"""TEST INFRASTRUCTURE -- a NumPy model of the *reformulated* algorithm the CUDA kernels run.

It interprets the product tables (pyjac_b200/tables.py) exactly the way
pyjac_b200/csrc/pyjac_b200.cu does (same per-reaction scalars, same dense + sparse
assembly), vectorised over states.  Checking it against the oracle on the CPU validates
the tables and the algebra before any GPU time is spent; the GPU tests then only have to
show that the kernel implements this model.
"""
import numpy as np

from pyjac_b200 import tables as tb

LN10 = np.log(10.0)


def evaluate(Tb, P, y):
    """Returns dict(conc, fwd, rev, pres_mod, spec_rates, dydt, jac); layouts as the oracle."""
    d = Tb['dims']
    nsp, nr, nrev, npd, nraw = (int(v) for v in d[:5])
    first_pm = int(d[8])
    RU8, ln_pa_ru = Tb['cst'][:2]
    n = y.shape[0]
    last = nsp - 1
    T = y[:, 0].copy()
    P = np.asarray(P, dtype=np.float64)
    logT, iT = np.log(T), 1.0 / T
    w, iw, ruw, mwf = Tb['sp_w'], Tb['sp_iw'], Tb['sp_ruw'], Tb['sp_mwf']
    Y = np.empty((n, nsp))
    Y[:, :last] = y[:, 1:]
    Y[:, last] = 1.0 - y[:, 1:].sum(axis=1)
    mw_avg = 1.0 / (Y * iw).sum(axis=1)
    rho = P * mw_avg / (RU8 * T)
    rho_inv = 1.0 / rho
    conc = np.ones((n, nsp + 1))
    conc[:, :nsp] = rho[:, None] * Y * iw
    m = P / (RU8 * T)

    nasa = Tb['sp_nasa'].reshape(nsp, 2, 16)
    cp = np.empty((n, nsp)); h = np.empty((n, nsp)); dcp = np.empty((n, nsp))
    B = np.zeros((n, nsp + 1)); dB = np.zeros((n, nsp + 1))
    for k in range(nsp):
        lo = T <= Tb['sp_tmid'][k]
        c = np.where(lo[:, None], nasa[k, 0][None, :], nasa[k, 1][None, :]).T
        cp[:, k] = ruw[k] * (c[0] + T * (c[1] + T * (c[2] + T * (c[3] + c[4] * T))))
        h[:, k] = ruw[k] * (c[5] + T * (c[0] + T * (c[6] + T * (c[7] + T * (c[8] + c[9] * T)))))
        dcp[:, k] = ruw[k] * (c[1] + T * (2.0 * c[2] + T * (3.0 * c[3] + 4.0 * c[4] * T)))
        B[:, k] = c[10] + c[11] * logT + T * (c[6] + T * (c[12] + T * (c[13] + c[14] * T))) - c[5] * iT
        dB[:, k] = (c[11] + c[5] * iT) * iT + c[6] + T * (c[7] + T * (c[8] + c[9] * T))
    cp_avg = (Y * cp).sum(axis=1)
    wdcp = (Y * dcp).sum(axis=1)

    flags, slots, arr = Tb['rx_flags'], Tb['rx_slots'].reshape(nr, 6), Tb['rx_arr'].reshape(nr, 4)
    par_all = Tb['pm_par'].reshape(-1, tb.NPAR)
    raw = np.zeros((n, nraw + 1))
    rx_dst = Tb['rx_dst'].reshape(nr, 8)
    R4 = np.zeros((nr, 4, n))
    RH = np.zeros((nr, n))
    hw = np.zeros((n, nsp + 1))
    hw[:, :nsp] = h * w[None, :]
    fwd = np.zeros((n, nr)); rev = np.zeros((n, max(nrev, 1))); pres_mod = np.zeros((n, max(npd, 1)))
    lg10 = lambda x: np.log10(np.maximum(x, 1.0e-300))
    for p in range(nr):
        fl = int(flags[p])
        s = slots[p]
        lnA, b, Ta, lnKc = arr[p]
        c = [conc[:, s[a]] for a in range(6)]
        nre = (fl >> tb.NRE_SHIFT) & 15
        npr = (fl >> tb.NPR_SHIFT) & 15
        lnkf = lnA + b * logT - Ta * iT
        kf = np.exp(lnkf)
        f = kf * c[0] * c[1] * c[2]
        isrev = bool(fl & tb.F_REV)
        if isrev:
            sB = (B[:, s[3]] + B[:, s[4]] + B[:, s[5]]) - (B[:, s[0]] + B[:, s[1]] + B[:, s[2]])
            kr = np.exp(lnkf - sB - lnKc)
            r = kr * c[3] * c[4] * c[5]
        else:
            kr = np.zeros(n)
            r = np.zeros(n)
        net = f - r
        fwd[:, Tb['rx_orig'][p]] = f
        if isrev:
            rev[:, Tb['rx_rev_idx'][p]] = r

        # --- pressure modification
        PM = np.ones(n)
        pmt = np.zeros(n)
        Xd = np.zeros(n)
        has_pm = bool(fl & (tb.F_THD | tb.F_PDEP))
        if has_pm:
            mi = p - first_pm
            par = par_all[mi]
            thd = m.copy()
            for e in range(Tb['pm_eff_off'][mi], Tb['pm_eff_off'][mi + 1]):
                thd = thd + Tb['pm_eff_am1'][e] * conc[:, Tb['pm_eff_sp'][e]]
            if fl & tb.F_PDEP:
                sp = int(Tb['pm_sp'][mi])
                ct = conc[:, sp] if sp >= 0 else thd
                e1 = np.exp(par[0] + par[1] * logT - par[2] * iT)
                Pr = ct * e1
                dpr4 = par[3] + par[2] * iT - 1.0
                dpr = par[1] + par[2] * iT - 1.0
                i1p = 1.0 / (1.0 + Pr)
                low = bool(fl & tb.F_LOW)
                Xd = dpr4 * iT * i1p if low else -Pr * dpr4 * iT * i1p
                g = i1p if low else -Pr * i1p
                F = np.ones(n)
                if fl & tb.F_TROE:
                    e3 = np.exp(T / par[7]); e1_ = np.exp(T / par[9])
                    Fc = par[6] * e3 + par[8] * e1_
                    dF = par[11] * e3 - par[12] * e1_
                    if fl & tb.F_TROE_T2:
                        e2 = np.exp(par[10] * iT)
                        Fc = Fc + e2
                        dF = dF + par[13] * iT * iT * e2
                    lnFc = np.log(np.maximum(Fc, 1.0e-300))
                    lF, lP = lnFc / LN10, lg10(Pr)
                    A = lP - 0.67 * lF - 0.4
                    Bq = 0.806 - 1.1762 * lF - 0.14 * lP
                    q1 = 1.0 + A * A / (Bq * Bq)
                    lnF_AB = 2.0 * lnFc * A / (Bq * Bq * Bq * q1 * q1)
                    F = np.exp(lnFc / q1)
                    Xd = Xd + (1.0 / (Fc * q1) - lnF_AB * (-0.67 / LN10 * Bq + 1.1762 / LN10 * A) / Fc) * dF \
                        - lnF_AB * (Bq / LN10 + 0.14 / LN10 * A) * dpr * iT
                    g = g - lnF_AB * (Bq / LN10 + A * 0.14 / LN10)
                elif fl & tb.F_SRI:
                    lP = lg10(Pr)
                    X = 1.0 / (1.0 + lP * lP)
                    base = par[14] * np.exp(-par[15] * iT) + np.exp(-T / par[16])
                    F = base ** X
                    if fl & tb.F_SRI5:
                        F = F * par[17] * T ** par[18]
                    eb = np.exp(par[23] * iT); ec = np.exp(T / par[25])
                    den = par[26] * eb + ec
                    Xd = Xd + X * ((par[22] * iT * iT * eb - par[24] * ec) / den
                                   - X * (2.0 / LN10) * lP * dpr * np.log(den) * iT)
                    if fl & tb.F_SRI5_DT:
                        Xd = Xd + par[27] * iT
                    g = g - X * X * (2.0 / LN10) * lP * np.log(par[19] * np.exp(par[20] * iT) + np.exp(T / par[21]))
                Fi = F * i1p
                PM = Fi * Pr if low else Fi
                if fl & tb.F_PMT:
                    pmt = g * net
            else:
                PM = thd
                if fl & tb.F_PMT:
                    pmt = net
            pres_mod[:, Tb['rx_pm_idx'][p]] = PM

        # --- T column scalar
        if fl & tb.F_NO_T:
            tT = np.zeros(n)
        else:
            dk = b + Ta * iT
            if isrev:
                sdB = (dB[:, s[3]] + dB[:, s[4]] + dB[:, s[5]]) - (dB[:, s[0]] + dB[:, s[1]] + dB[:, s[2]])
                elem = net * dk + f * (1.0 - nre) - r * ((1.0 - npr) - T * sdB)
            else:
                elem = f * (dk + (1.0 - nre))
            if fl & tb.F_PDEP:
                tT = (PM * Xd * net + PM * iT * elem) * rho_inv
            elif fl & tb.F_THD:
                tT = (-PM * net * iT + PM * iT * elem) * rho_inv
            else:
                tT = iT * elem * rho_inv

        # --- Y columns: coefficient of 1 (X1), of W_j/W_N (X2), sparse values
        n1 = nre + (1 if fl & tb.F_EFFN1 else 0)
        n2 = (npr + (1 if fl & tb.F_EFFN1 else 0)) if isrev else 0
        inner = n1 * f - n2 * r
        if fl & tb.F_PMT_INJ:
            inner = inner + pmt
        jy = -mw_avg * rho_inv * PM * inner
        if fl & tb.F_PMT_INJ:
            pmt = pmt * e1 * Fi
        aN = par[4] if has_pm else 0.0
        adef = par[5] if has_pm else 0.0
        X1 = jy + adef * pmt
        X2 = -jy - aN * pmt
        dst = rx_dst[p]
        for a in range(3):
            if s[a] == nsp:
                continue
            dv = PM * kf * c[(a + 1) % 3] * c[(a + 2) % 3]
            if s[a] == last:
                X2 = X2 - dv
            else:
                assert dst[a] != 0xFFFF
                raw[:, dst[a]] = dv
        if isrev:
            for a in range(3):
                if s[3 + a] == nsp:
                    continue
                dv = -(PM * kr * c[3 + (a + 1) % 3] * c[3 + (a + 2) % 3])
                if s[3 + a] == last:
                    X2 = X2 - dv
                else:
                    assert dst[3 + a] != 0xFFFF
                    raw[:, dst[3 + a]] = dv
        if fl & tb.F_EFF_SLOTS:
            rb = int(dst[6])
            for e in range(Tb['pm_eff_off'][mi], Tb['pm_eff_off'][mi + 1]):
                if Tb['pm_eff_sp'][e] != last:
                    raw[:, rb] = pmt * Tb['pm_eff_am1'][e]
                    rb += 1
        if fl & tb.F_WANT_PMT:
            raw[:, dst[7]] = pmt
        R4[p, 0], R4[p, 1], R4[p, 2], R4[p, 3] = net * PM, tT, X1, X2
        RH[p] = (hw[:, s[3]] + hw[:, s[4]] + hw[:, s[5]]) - (hw[:, s[0]] + hw[:, s[1]] + hw[:, s[2]])

    # --- species reductions: packed (reaction | hi16(nu) << 16) lists per species
    def coef(word):
        return float(np.uint64((int(word) >> 16) << 48).view(np.float64))
    red_pk = Tb['red_pk'].view(np.uint32)
    wdot = np.zeros((n, nsp)); tcol = np.zeros((n, nsp)); Ak = np.zeros((n, nsp)); Bk = np.zeros((n, nsp))
    for k in range(nsp):
        for e in range(Tb['red_off'][k], Tb['red_off'][k + 1]):
            wd = red_pk[e]
            cf, rxn = coef(wd), int(wd) & 0xFFFF
            wdot[:, k] += cf * R4[rxn, 0]; tcol[:, k] += cf * R4[rxn, 1]
            Ak[:, k] += cf * R4[rxn, 2]; Bk[:, k] += cf * R4[rxn, 3]
    comp = wdot * (mw_avg * rho_inv)[:, None]
    Ak = Ak + comp                       # unscaled a_k, b_k, as the kernel keeps them
    Bk = Bk - comp

    # --- sparse gather into the dense tile: fixed-length classes, then quad entries
    nfix, nq, nq_j = (int(v) for v in Tb['dims3'][:3])
    tile = np.zeros((n, nsp * nsp))
    raw[:, nraw] = 0.0
    d_con, q_con = Tb['d_con'].view(np.uint32), Tb['q_con'].view(np.uint32)
    for c, ln in enumerate((8, 4, 2, 1)):
        for e in range(Tb['d_cls'][c], Tb['d_cls'][c + 1]):
            o = Tb['d_ccon'][c] + (e - Tb['d_cls'][c]) * ln
            acc = np.zeros(n)
            for wd in d_con[o:o + ln]:
                acc += coef(wd) * raw[:, int(wd) & 0xFFFF]
            assert not tile[:, Tb['d_dst'][e]].any()
            tile[:, Tb['d_dst'][e]] = acc
    for e in range(nq):
        acc = np.zeros(n)
        assert (Tb['q_off'][e + 1] - Tb['q_off'][e]) % 16 == 0
        for wd in q_con[Tb['q_off'][e]:Tb['q_off'][e + 1]]:
            if e < nq_j:
                acc += coef(wd) * raw[:, int(wd) & 0xFFFF]
            else:
                acc += RH[int(wd) >> 16] * raw[:, int(wd) & 0xFFFF]
        assert not tile[:, Tb['q_dst'][e]].any()
        tile[:, Tb['q_dst'][e]] = acc
    tile = tile.reshape(n, nsp, nsp)       # [state, col, row]

    # --- assembly: row k+1 of column j+1 is iw_j (W_k a_k + W_k b_k mwf_j + W_k tile);
    # row 0 (energy equation) uses -1/cp_avg times the dH-weighted sums
    jac = np.zeros((n, nsp, nsp))          # [state, col, row]
    wt = 1.0 / cp_avg
    hwk = hw[:, :nsp]
    H1 = (hwk * wdot).sum(axis=1)
    HA = (hwk * Ak).sum(axis=1)
    HB = (hwk * Bk).sum(axis=1)
    HT = (hwk * tcol).sum(axis=1)
    SCP = (cp * w[None, :] * wdot).sum(axis=1)
    XT = H1 / (rho * cp_avg * cp_avg)
    WA, WB = w[None, :last] * Ak[:, :last], w[None, :last] * Bk[:, :last]
    for j in range(nsp - 1):
        v = np.empty((n, nsp))
        v[:, 1:] = iw[j] * (WA + WB * mwf[j] + w[None, :last] * tile[:, j + 1, 1:])
        v[:, 0] = iw[j] * (-wt * HA + -wt * HB * mwf[j] + -wt * tile[:, j + 1, 0]) \
            + XT * (cp[:, j] - cp[:, last])
        jac[:, j + 1, :] = v
    jac[:, 0, 1:] = w[None, :last] * tcol[:, :last]
    s0 = -wdcp / cp_avg * H1 + SCP + HT * rho
    jac[:, 0, 0] = -s0 / (rho * cp_avg)

    dydt = np.empty((n, nsp))
    dydt[:, 0] = -1.0 / (rho * cp_avg) * H1
    dydt[:, 1:] = wdot[:, :last] * w[None, :last] / rho[:, None]
    return dict(conc=conc[:, :nsp], fwd=fwd, rev=rev[:, :nrev], pres_mod=pres_mod[:, :npd],
                spec_rates=wdot, dydt=dydt, jac=jac.reshape(n, nsp * nsp))
