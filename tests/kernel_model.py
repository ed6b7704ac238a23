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
        rb = int(Tb['rx_raw_base'][p])
        for a in range(3):
            if s[a] == nsp:
                continue
            dv = PM * kf * c[(a + 1) % 3] * c[(a + 2) % 3]
            if s[a] == last:
                X2 = X2 - dv
            else:
                raw[:, rb] = dv
                rb += 1
        if isrev:
            for a in range(3):
                if s[3 + a] == nsp:
                    continue
                dv = -(PM * kr * c[3 + (a + 1) % 3] * c[3 + (a + 2) % 3])
                if s[3 + a] == last:
                    X2 = X2 - dv
                else:
                    raw[:, rb] = dv
                    rb += 1
        if fl & tb.F_EFF_SLOTS:
            for e in range(Tb['pm_eff_off'][mi], Tb['pm_eff_off'][mi + 1]):
                if Tb['pm_eff_sp'][e] != last:
                    raw[:, rb] = pmt * Tb['pm_eff_am1'][e]
                    rb += 1
        if fl & tb.F_WANT_PMT:
            raw[:, rb] = pmt
        R4[p, 0], R4[p, 1], R4[p, 2], R4[p, 3] = net * PM, tT, X1, X2
        RH[p] = (hw[:, s[3]] + hw[:, s[4]] + hw[:, s[5]]) - (hw[:, s[0]] + hw[:, s[1]] + hw[:, s[2]])

    # --- species reductions (chunked two-level, as the kernel does)
    nchunk = int(d[14])
    part = np.zeros((nchunk, 4, n))
    crx, cnu = Tb['chk_rx'].reshape(-1, tb.RCH), Tb['chk_nu'].reshape(-1, tb.RCH)
    for c in range(nchunk):
        for e in range(tb.RCH):
            part[c] += cnu[c, e] * R4[crx[c, e]]
    wdot = np.zeros((n, nsp)); tcol = np.zeros((n, nsp)); Ak = np.zeros((n, nsp)); Bk = np.zeros((n, nsp))
    for k in range(nsp):
        for c in range(Tb['sp_chk_off'][k], Tb['sp_chk_off'][k + 1]):
            wdot[:, k] += part[c, 0]; tcol[:, k] += part[c, 1]; Ak[:, k] += part[c, 2]; Bk[:, k] += part[c, 3]
    comp = wdot * (mw_avg * rho_inv)[:, None]
    Ak = w[None, :] * (Ak + comp)
    Bk = w[None, :] * (Bk - comp)
    tc = w[None, :] * tcol

    # --- sparse gather: class-padded sub-entries, then the combine lists
    nsub, nsub_j, nsplit, zero_slot = int(d[5]), int(d[12]), int(d[13]), int(d[15])
    sval = np.zeros((n, zero_slot + 1))
    con = Tb['con'].view(np.uint32)
    raw[:, nraw] = 0.0
    wt = 1.0 / cp_avg
    lens = [8, 4, 2, 1, 8, 4, 2, 1]
    for c in range(8):
        for e in range(Tb['cls_sub'][c], Tb['cls_sub'][c + 1]):
            o = Tb['cls_con'][c] + (e - Tb['cls_sub'][c]) * lens[c]
            for cc in con[o:o + lens[c]]:
                cc = int(cc)
                if c < 4:
                    cf = float(np.uint32(cc & 0xFFFF0000).view(np.float32))
                    sval[:, e] += cf * raw[:, cc & 0xFFFF]
                else:
                    sval[:, e] += RH[cc >> 16] * raw[:, cc & 0xFFFF]
            sval[:, e] *= Tb['sub_w'][e] if c < 4 else -wt
    for t in range(nsplit):
        for cidx in range(Tb['cmb_off'][t], Tb['cmb_off'][t + 1]):
            sval[:, nsub + t] += sval[:, Tb['cmb_idx'][cidx]]

    # --- assembly: output row r of column j+1 is iw_j (rowA[r] + rowB[r] mwf_j + sval[jmap[j][r]])
    jmap = Tb['jmap'].reshape(nsp - 1, nsp)
    jac = np.zeros((n, nsp, nsp))          # [state, col, row]
    H1 = (hw[:, :nsp] * wdot).sum(axis=1)
    HA = (h * Ak).sum(axis=1)
    HB = (h * Bk).sum(axis=1)
    HT = (h * tc).sum(axis=1)
    SCP = (cp * w[None, :] * wdot).sum(axis=1)
    XT = H1 / (rho * cp_avg * cp_avg)
    rowA = np.concatenate([(-wt * HA)[:, None], Ak[:, :last]], axis=1)
    rowB = np.concatenate([(-wt * HB)[:, None], Bk[:, :last]], axis=1)
    for j in range(nsp - 1):
        v = iw[j] * (rowA + rowB * mwf[j] + sval[:, jmap[j]])
        v[:, 0] += XT * (cp[:, j] - cp[:, last])
        jac[:, j + 1, :] = v
    jac[:, 0, 1:] = tc[:, :last]
    s0 = -wdcp / cp_avg * H1 + SCP + HT * rho
    jac[:, 0, 0] = -s0 / (rho * cp_avg)

    dydt = np.empty((n, nsp))
    dydt[:, 0] = -1.0 / (rho * cp_avg) * H1
    dydt[:, 1:] = wdot[:, :last] * w[None, :last] / rho[:, None]
    return dict(conc=conc[:, :nsp], fwd=fwd, rev=rev[:, :nrev], pres_mod=pres_mod[:, :npd],
                spec_rates=wdot, dydt=dydt, jac=jac.reshape(n, nsp * nsp))
