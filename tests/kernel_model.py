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


def evaluate(Tb, P, y, plan=5):
    """Returns dict(conc, fwd, rev, pres_mod, spec_rates, dydt, jac); layouts as the oracle.
    plan = 5: the schedule tables of k_eval (p5_*); plan = 6: the record streams of k_jac6 (p6_*)."""
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
        dk_ = b + Ta * iT
        if fl & tb.F_PLOG:
            # rs:598-632, cj:1687-1850: Arrhenius set of the end ranges, log-P interpolation between
            pq = Tb['plog_par'].reshape(-1, 8)[Tb['plog_off'][p]:Tb['plog_off'][p + 1]]
            idx = (P[:, None] > pq[None, :, 0]).sum(axis=1)
            lo = np.clip(idx - 1, 0, len(pq) - 1)
            mid = (idx > 0) & (idx < len(pq))
            k1 = pq[lo, 1] + pq[lo, 2] * logT - pq[lo, 3] * iT
            d1 = pq[lo, 2] + pq[lo, 3] * iT
            hi_ = np.clip(lo + 1, 0, len(pq) - 1)
            k2 = pq[hi_, 1] + pq[hi_, 2] * logT - pq[hi_, 3] * iT
            wgt = np.where(mid, (np.log(P) - pq[lo, 4]) * pq[lo, 5], 0.0)
            lnkf = k1 + (k2 - k1) * wgt
            dk_ = d1 + (pq[lo, 6] + pq[lo, 7] * iT) * wgt
        rat = 1.0
        if fl & tb.F_CHEB:
            # rs:149-251 (rate, {:.8e} constants) and cj:1532-1684 (temperature derivative and the rate
            # constant of the species part, {:.16e} reduced variables)
            cq = Tb['cheb_par'][Tb['cheb_off'][p]:Tb['cheb_off'][p + 1]]
            n_t, n_p = int(cq[0]), int(cq[1])
            c8 = cq[12:12 + n_t * n_p].reshape(n_t, n_p)
            c16 = cq[12 + n_t * n_p:].reshape(n_t, n_p)
            l10p = np.log10(P)
            cheb_T = np.polynomial.chebyshev.chebvander

            def series(co, tr, pr, second_kind):
                Pv = cheb_T(pr, n_p - 1)                      # (n, n_p)
                if second_kind:                               # sum_i co[i] U_{i-1}(tr), row 0 unused
                    U = np.zeros((len(tr), n_t))
                    U[:, 1] = 1.0
                    if n_t > 2:
                        U[:, 2] = 2.0 * tr
                    for i_ in range(3, n_t):
                        U[:, i_] = 2.0 * tr * U[:, i_ - 1] - U[:, i_ - 2]
                    Tv = U
                else:
                    Tv = cheb_T(tr, n_t - 1)
                return np.einsum('ni,ij,nj->n', Tv, co, Pv)
            tr8, pr8 = (2.0 * iT - cq[2]) / cq[3], (2.0 * l10p - cq[4]) / cq[5]
            tr16, pr16 = (2.0 * iT - cq[6]) / cq[7], (2.0 * l10p - cq[8]) / cq[9]
            lnkf_r = LN10 * series(c8, tr8, pr8, False)
            lnkf = LN10 * series(c8, tr16, pr16, False)
            dk_ = series(c16, tr16, pr16, True) * cq[10] * iT
            rat = np.exp(lnkf_r - lnkf)
        sg = -1.0 if fl & tb.F_NEGA else 1.0                  # rs:108-141
        kf = sg * np.exp(lnkf)
        f = kf * c[0] * c[1] * c[2] * rat
        isrev = bool(fl & tb.F_REV)
        if isrev:
            sB = (B[:, s[3]] + B[:, s[4]] + B[:, s[5]]) - (B[:, s[0]] + B[:, s[1]] + B[:, s[2]])
            kr = sg * np.exp(lnkf - sB - lnKc)
            r = kr * c[3] * c[4] * c[5] * rat
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
            dk = dk_
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

    if plan == 6:
        return _assemble6(dict(locals()))

    # --- phase C of the plan: per species, +1 and -1 lists of reaction-row byte offsets, two
    # entries per unit and sub-group (padding entries point at the all-zero reaction row nr)
    cfg = Tb['p5_cfg']
    gs, nt, nw, nsub = (int(v) for v in cfg[:4])
    assert nsub * gs == 64 and nt == nw * 32
    RB = gs * 8
    SPB, RXB = 8 * RB, 5 * RB
    R4z = np.concatenate([R4, np.zeros((2, 4, n))], axis=0)
    wdot = np.zeros((n, nsp)); tcol = np.zeros((n, nsp)); Ak = np.zeros((n, nsp)); Bk = np.zeros((n, nsp))
    def sp_slot0(off):
        k_, chunk = off // SPB, (off % SPB) // RB
        return k_, chunk ^ (k_ & 1)
    c_item, c_str = Tb['p5_c_item'].reshape(-1, 4), Tb['p5_c_str'].view(np.uint32)
    coop = int(cfg[11])
    assert coop >= 1 and nsub % coop == 0
    seen_k = set()
    for wp in range(nw):
        for it in range(Tb['p5_c_off'][wp], Tb['p5_c_off'][wp + 1]):
            u, n_p, n_m = (int(v) for v in c_item[it][:3])
            hdr = [(int(c_str[(u * nsub + sub) * 2]), int(c_str[(u * nsub + sub) * 2 + 1])) for sub in range(nsub)]
            u += 1
            part = np.zeros((nsub, 4, n))
            for sign, cnt in ((1.0, n_p), (-1.0, n_m)):
                for i in range(cnt):
                    for sub in range(nsub):
                        for h in range(2):
                            off = int(c_str[((u + i) * nsub + sub) * 2 + h])
                            assert off % RXB == 0
                            part[sub] += sign * R4z[off // RXB]
                u += cnt
            for g0 in range(0, nsub, coop):
                spoff, lead = hdr[g0]
                assert all(hdr[g0 + d][0] == spoff for d in range(coop))
                assert lead == (0 if spoff == 0xFFFFFFFF else 1) and all(hdr[g0 + d][1] == 0 for d in range(1, coop))
                tot = part[g0:g0 + coop].sum(axis=0)
                if spoff == 0xFFFFFFFF:
                    assert not tot.any()
                    continue
                k, sl0 = sp_slot0(spoff)
                assert sl0 == 0
                assert k not in seen_k
                seen_k.add(k)
                wdot[:, k], tcol[:, k], Ak[:, k], Bk[:, k] = tot
    assert seen_k == set(range(nsp))
    comp = wdot * (mw_avg * rho_inv)[:, None]
    Ak = Ak + comp
    Bk = Bk - comp

    # --- phase DE of the plan: elements in steps of NSUB with one padded length L per step;
    # unit u of a warp's stream is e_str[(u * NSUB + sub) * 2 + {0, 1}]
    raw[:, nraw] = 0.0
    RHz = np.concatenate([RH, np.zeros((2, n))], axis=0)
    wt = 1.0 / cp_avg
    hwk = hw[:, :nsp]
    H1 = (hwk * wdot).sum(axis=1)
    HA = (hwk * Ak).sum(axis=1)
    HB = (hwk * Bk).sum(axis=1)
    HT = (hwk * tcol).sum(axis=1)
    SCP = (cp * w[None, :] * wdot).sum(axis=1)
    XT = H1 / (rho * cp_avg * cp_avg)
    slots8 = np.zeros((n, nsp, 8))                 # the species rows as the kernel keeps them
    slots8[:, :, 4], slots8[:, :, 5], slots8[:, :, 6] = w[None, :] * Ak, w[None, :] * Bk, w[None, :] * tcol
    slots8[:, :, 7] = cp
    colfac = Tb['p5_colfac'].reshape(nsp, 2)
    jac = np.full((n, nsp * nsp), np.nan)          # [state, col * nsp + row]; every element once
    NULL_E = 0x3FFFFF
    # the factored record k_eval<.., M_FACT> writes instead (include/pyjac_b200.h)
    fac_map, fac_nnz = Tb['p5_fac_map'], int(cfg[15])
    fac = np.full((n, nsp + 3 * last + fac_nnz), np.nan)

    def rawrow(off):
        assert off % RB == 0 and off // RB <= nraw + 1      # rows nraw, nraw + 1: zeros
        return raw[:, min(off // RB, nraw)]

    def sp_slot(off):
        """(species, logical slot) of a byte offset into the species rows; the slot pair of odd
        species is swapped (plan.sp_even)."""
        assert off % RB == 0
        k, chunk = off // SPB, (off % SPB) // RB
        return k, chunk ^ (k & 1)

    def put(eidx, y, extra):
        """dense part of element eidx from the species-row offset / column word y, plus extra."""
        off, col = y & 0xFFFFF, y >> 20
        k, sl = sp_slot(off)
        k2, sl2 = sp_slot(off ^ RB)
        assert k2 == k and sl == 4 and sl2 == 5 and col >= 1
        assert np.isnan(jac[:, eidx]).all() and eidx == col * nsp + k + 1
        jac[:, eidx] = colfac[col, 0] * slots8[:, k, sl] + colfac[col, 1] * slots8[:, k, sl2] + extra
        assert fac_map[eidx] >= nsp + 3 * last and np.isnan(fac[:, fac_map[eidx]]).all()
        fac[:, fac_map[eidx]] = extra

    d_str, d_item = Tb['p5_d_str'].view(np.uint32), Tb['p5_d_item'].reshape(-1, 2)
    s_str, o_str = Tb['p5_s_str'].view(np.uint32), Tb['p5_o_str'].view(np.uint32)
    t_str, t_item = Tb['p5_t_str'].view(np.uint32), Tb['p5_t_item'].reshape(-1, 2)
    tcoop = int(cfg[12])
    for wp in range(nw):
        # class D: dense-only elements by row
        for it in range(int(Tb['p5_d_off'][wp]), int(Tb['p5_d_off'][wp + 1])):
            u, ncol = int(d_item[it][0]), int(d_item[it][1])
            assert ncol % 4 == 0
            for sub in range(nsub):
                base, e0 = int(d_str[(u * nsub + sub) * 2]), int(d_str[(u * nsub + sub) * 2 + 1])
                if base == 0xFFFFFFFF:
                    assert all(int(d_str[((u + 1 + i) * nsub + sub) * 2]) == NULL_E for i in range(ncol))
                    continue
                k, sl = sp_slot(base)
                assert sl == 0 and e0 in (k + 1, NULL_E)
                if e0 != NULL_E:                                   # temperature column: W_k * T-term
                    assert np.isnan(jac[:, e0]).all()
                    jac[:, e0] = slots8[:, k, 6]
                    fac[:, nsp + k], fac[:, nsp + last + k], fac[:, nsp + 2 * last + k] = slots8[:, k, 6], slots8[:, k, 4], slots8[:, k, 5]
                for i in range(ncol):
                    eidx, col = int(d_str[((u + 1 + i) * nsub + sub) * 2]), int(d_str[((u + 1 + i) * nsub + sub) * 2 + 1])
                    if eidx == NULL_E:
                        continue
                    assert eidx == col * nsp + k + 1 and col >= 1 and np.isnan(jac[:, eidx]).all() and fac_map[eidx] == -1
                    jac[:, eidx] = colfac[col, 0] * slots8[:, k, 4] + colfac[col, 1] * slots8[:, k, 5]
        # class S: two uint4 per element and step, overflow units beyond the first two
        lo, hi = int(Tb['p5_s_off'][wp]), int(Tb['p5_s_off'][wp + 1])
        assert (hi - lo) % 2 == 0
        ou = int(Tb['p5_o_off'][wp])
        for st in range(lo, hi):
            A = [s_str[((st * 2) * nsub + sub) * 4:((st * 2) * nsub + sub) * 4 + 4] for sub in range(nsub)]
            B = [s_str[((st * 2 + 1) * nsub + sub) * 4:((st * 2 + 1) * nsub + sub) * 4 + 4] for sub in range(nsub)]
            L = int(A[0][0]) >> 22
            assert all(int(a_[0]) >> 22 == L for a_ in A)
            n_ovf = (max(L - 2, 0) + 3) // 4 * 4
            for sub in range(nsub):
                eidx, y = int(A[sub][0]) & 0x3FFFFF, int(A[sub][1])
                pw_ = A[sub][2:4].copy().view(np.float64)[0]
                acc = np.zeros(n)
                ents = [int(v) for v in B[sub]]
                for i in range(n_ovf):
                    ents += [int(o_str[((ou + i) * nsub + sub) * 2]), int(o_str[((ou + i) * nsub + sub) * 2 + 1])]
                for c in ents:
                    acc = acc + (-1.0 if c & 1 else 1.0) * rawrow(c & ~1)
                if L == 1:                                 # the kernel skips these two
                    assert ents[2] // RB >= nraw and ents[3] // RB >= nraw
                if eidx == NULL_E:
                    assert not acc.any()
                    continue
                put(eidx, y, pw_ * acc)
            ou += n_ovf
        assert ou == Tb['p5_o_off'][wp + 1]
        # class T: energy-equation row, tcoop sub-groups per column
        for it in range(int(Tb['p5_t_off'][wp]), int(Tb['p5_t_off'][wp + 1])):
            u, nun = int(t_item[it][0]), int(t_item[it][1])
            assert nun % 4 == 0
            part = np.zeros((nsub, n))
            hdr = [(int(t_str[(u * nsub + sub) * 2]), int(t_str[(u * nsub + sub) * 2 + 1])) for sub in range(nsub)]
            for i in range(nun):
                for sub in range(nsub):
                    ro, xo = int(t_str[((u + 1 + i) * nsub + sub) * 2]), int(t_str[((u + 1 + i) * nsub + sub) * 2 + 1])
                    assert xo % RXB == 0
                    part[sub] += RHz[xo // RXB] * rawrow(ro)
            for g0 in range(0, nsub, tcoop):
                col, lead = hdr[g0][0] & 0xFFFF, hdr[g0][0] >> 16
                E0 = part[g0:g0 + tcoop].sum(axis=0)
                if col == 0:
                    assert not E0.any() and lead == 0
                    continue
                assert lead == 1 and all(hdr[g0 + d][0] == col for d in range(1, tcoop))
                kcp, slcp = sp_slot(hdr[g0][1])
                assert kcp == col - 1 and slcp == 7
                eidx = col * nsp
                assert np.isnan(jac[:, eidx]).all()
                pj, qj = colfac[col]
                jac[:, eidx] = pj * (-wt * HA) + qj * (-wt * HB) + pj * (-wt) * E0 \
                    + XT * (cp[:, col - 1] - cp[:, last])
                fac[:, col] = jac[:, eidx]
    s0 = -wdcp / cp_avg * H1 + SCP + HT * rho
    jac[:, 0] = -s0 / (rho * cp_avg)
    fac[:, 0] = jac[:, 0]
    assert not np.isnan(jac).any() and not np.isnan(fac).any()
    jac = jac.reshape(n, nsp, nsp)

    dydt = np.empty((n, nsp))
    dydt[:, 0] = -1.0 / (rho * cp_avg) * H1
    dydt[:, 1:] = wdot[:, :last] * w[None, :last] / rho[:, None]
    return dict(conc=conc[:, :nsp], fwd=fwd, rev=rev[:, :nrev], pres_mod=pres_mod[:, :npd],
                spec_rates=wdot, dydt=dydt, jac=jac.reshape(n, nsp * nsp), fac=fac)


def _assemble6(v):
    """Phases C and DE of k_jac6: interprets the per-warp record streams (pyjac_b200/plan6.py) record
    by record, exactly in the order the kernel consumes them."""
    from pyjac_b200 import plan6
    Tb, n, nsp, nr, nraw, last = v['Tb'], v['n'], v['nsp'], v['nr'], v['nraw'], v['last']
    R4, RH, raw, hw, cp, w = v['R4'], v['RH'], v['raw'], v['hw'], v['cp'], v['w']
    mw_avg, rho, rho_inv, cp_avg, wdcp = v['mw_avg'], v['rho'], v['rho_inv'], v['cp_avg'], v['wdcp']
    cfg = [int(x) for x in Tb['p6_cfg']]
    gs, nt, nw, nsub = cfg[:4]
    coop, tcoop, p_c0, ncorr, chb, nslot, chr_ = cfg[16:23]
    assert nsub * gs == 64 and nt == nw * 32 and chr_ == chb // (nsub * 16)
    L = plan6.layout6(nsp, nr, ncorr, nraw, gs, nw)
    assert [L[k] for k in ('SP', 'RX', 'XC', 'RAW', 'ET', 'SC', 'PA', 'CF', 'ring', 'mbar', 'bytes')] == cfg[4:15]
    assert L['bytes'] <= plan6.SMEM_LIMIT
    RB = gs * 8
    SPB = plan6.SP_SLOTS * RB
    hdr = Tb['p6_hdr'].reshape(nw, 8)
    st = Tb['p6_str'].view(np.uint32)
    rec6 = None
    # reaction rows as the kernel keeps them: row index v = p * 4 + (p & 1) is the even-slot base
    # (net, X1 at +0, +2), the odd slots (tT, dH) sit at (v ^ 1) + {0, 2}
    RXrows = {}
    for p in range(nr):
        b = p * 4 + (p & 1)
        RXrows[b] = R4[p, 0]; RXrows[b + 2] = R4[p, 2]
        RXrows[b ^ 1] = R4[p, 1]; RXrows[(b ^ 1) + 2] = RH[p]
    for p in (nr, nr + 1):
        b = p * 4 + (p & 1)
        for r_ in (b, b + 2, b ^ 1, (b ^ 1) + 2):
            RXrows[r_] = np.zeros(n)
    XC = np.zeros((ncorr + 2, n))
    for p in range(p_c0, nr):
        XC[p - p_c0] = R4[p, 2] + R4[p, 3]
    rawz = np.concatenate([raw[:, :nraw], np.zeros((n, 2))], axis=1)

    def sp_k(off):
        assert off % RB == 0
        k, chunk = off // SPB, (off % SPB) // RB
        assert chunk == (k & 1)                       # even-slot base
        return k

    wdot = np.zeros((n, nsp)); tcol = np.zeros((n, nsp)); Ak = np.zeros((n, nsp)); Bk = np.zeros((n, nsp))
    ET = np.zeros((n, max(last, 1)))
    seen_k, seen_j = set(), set()
    pos = [0] * nw
    rx_seen = set()

    def get(wp):
        base = int(hdr[wp, 0]) * (chb // 4) + pos[wp]
        pos[wp] += nsub * 4
        assert pos[wp] <= int(hdr[wp, 1]) * (chb // 4)
        return [[int(x) for x in st[base + 4 * sb:base + 4 * sb + 4]] for sb in range(nsub)]

    def f64(lo, hi):
        return float(np.array([lo, hi], dtype=np.uint32).view(np.float64)[0])

    # ---- phase B records: every reaction once, the padding sub-groups flagged
    rx6 = {}
    for wp in range(nw):
        for r in range(int(hdr[wp, 2]) + int(hdr[wp, 3])):
            q = [get(wp) for _ in range(4)]
            for sb in range(nsub):
                words = q[0][sb] + q[1][sb] + q[2][sb] + q[3][sb]
                fl, p = words[8], words[15] & 0xFFFF
                if fl & plan6.F_NULL:
                    continue
                assert p not in rx6
                rx6[p] = words
                assert (p >= Tb['dims'][8]) == (r < int(hdr[wp, 2]))
                assert bool(fl & plan6.F_CORR) == (p >= p_c0)
    assert sorted(rx6) == list(range(nr))
    v['rx6'] = rx6

    # ---- phase C
    for wp in range(nw):
        for r in range(int(hdr[wp, 4])):
            h = get(wp)
            nm, nx = (h[0][1] >> 8) & 0xFF, (h[0][1] >> 16) & 0xFF
            assert all(((x[1] >> 8) & 0xFF, (x[1] >> 16) & 0xFF) == (nm, nx) for x in h)
            part = np.zeros((nsub, 4, n))
            for u in range(nm):
                e = get(wp)
                for sb in range(nsub):
                    for i in range(8):
                        x = (e[sb][i // 2] >> (16 * (i & 1))) & 0xFFFF
                        sg = -1.0 if x & 0x8000 else 1.0
                        b = x & 0x7FFF
                        part[sb, 0] += sg * RXrows[b]
                        part[sb, 1] += sg * RXrows[b ^ 1]
                        part[sb, 2] += sg * RXrows[b + 2]
            for u in range(nx):
                e = get(wp)
                for sb in range(nsub):
                    for i in range(8):
                        x = (e[sb][i // 2] >> (16 * (i & 1))) & 0xFFFF
                        part[sb, 3] += (-1.0 if x & 0x8000 else 1.0) * XC[x & 0x7FFF]
            for g0 in range(0, nsub, coop):
                spoff, lead = h[g0][0], h[g0][1] & 1
                tot = part[g0:g0 + coop].sum(axis=0)
                if spoff == 0xFFFFFFFF:
                    assert not tot.any() and not lead
                    continue
                assert lead and all(h[g0 + d][0] == spoff and not (h[g0 + d][1] & 1) for d in range(1, coop))
                k = sp_k(spoff)
                assert k not in seen_k and f64(h[g0][2], h[g0][3]) == w[k]
                seen_k.add(k)
                wdot[:, k], tcol[:, k] = tot[0], tot[1]
                comp = tot[0] * mw_avg * rho_inv
                Ak[:, k] = tot[2] + comp
                Bk[:, k] = -Ak[:, k] + tot[3]
        for r in range(int(hdr[wp, 5])):
            h = get(wp)
            nrec = h[0][1]
            assert all(x[1] == nrec for x in h)
            part = np.zeros((nsub, n))
            for u in range(nrec):
                e = get(wp)
                for sb in range(nsub):
                    for i in range(4):
                        ro, b = e[sb][i] & 0xFFFF, e[sb][i] >> 16
                        part[sb] += RXrows[(b ^ 1) + 2] * rawz[:, ro]
            for g0 in range(0, nsub, tcoop):
                col, lead = h[g0][0] & 0xFFFF, h[g0][0] >> 16
                E0 = part[g0:g0 + tcoop].sum(axis=0)
                if col == 0:
                    assert not E0.any() and not lead
                    continue
                assert lead == 1 and all(h[g0 + d][0] == col for d in range(1, tcoop))
                assert col - 1 not in seen_j
                seen_j.add(col - 1)
                ET[:, col - 1] = E0
    assert seen_k == set(range(nsp)) and seen_j == set(range(last))

    # ---- phase DE
    wt = 1.0 / cp_avg
    hwk = hw[:, :nsp]
    H1 = (hwk * wdot).sum(axis=1)
    HA = (hwk * Ak).sum(axis=1)
    HB = (hwk * Bk).sum(axis=1)
    HT = (hwk * tcol).sum(axis=1)
    SCP = (cp * w[None, :] * wdot).sum(axis=1)
    XT = H1 / (rho * cp_avg * cp_avg)
    WA, WB, WT = w[None, :] * Ak, w[None, :] * Bk, w[None, :] * tcol
    colfac = Tb['p6_colfac'].reshape(nsp, 2)
    jac = np.full((n, nsp * nsp), np.nan)
    wpc = chb // 4                                   # words per chunk
    PER_REC = {0: 16, 1: 5, 2: 3, 4: 1, 6: 1, 7: 1}

    def ents_of(words):
        return [(wd >> (16 * h)) & 0xFFFF for wd in words for h in (0, 1)]

    def gather(ents):
        acc_ = np.zeros(n)
        for x in ents:
            acc_ = acc_ + (-1.0 if x & 0x8000 else 1.0) * rawz[:, x & 0x7FFF]
        return acc_

    for wp in range(nw):
        for sgm in range(int(hdr[wp, 6])):
            h = get(wp)
            assert all(x[0] & plan6.D_HDR for x in h)
            rows_k = []
            for sb in range(nsub):
                if not h[sb][0] & plan6.D_VALID:
                    rows_k.append(None)
                    continue
                k = sp_k(h[sb][1])
                assert k == (h[sb][0] >> 8) & 0xFF and f64(h[sb][2], h[sb][3]) == w[k]
                rows_k.append(k)
                if h[sb][0] & 1:
                    assert np.isnan(jac[:, k + 1]).all()
                    jac[:, k + 1] = WT[:, k]
            cw = get(wp)
            assert all(x == cw[0] for x in cw)
            counts = {7: cw[0][0] & 0xFF, 6: (cw[0][0] >> 8) & 0xFF, 4: (cw[0][0] >> 16) & 0xFF, 2: cw[0][0] >> 24,
                      1: cw[0][1] & 0xFF, 0: (cw[0][1] >> 8) & 0xFF}
            carry = [None] * nsub

            def put_el(sb, col, acc_):
                k = rows_k[sb]
                assert k is not None and 1 <= col < nsp
                e = col * nsp + k + 1
                assert np.isnan(jac[:, e]).all()
                tt = w[k] * acc_ + WA[:, k]
                jac[:, e] = colfac[col, 1] * WB[:, k] + colfac[col, 0] * tt

            for kd in (7, 6, 4, 2, 1, 0):
                for i in range(counts[kd]):
                    r = get(wp)
                    for sb in range(nsub):
                        x = r[sb]
                        if kd in (7, 6, 4):
                            ents = ents_of(x[1:4])
                            nread = 4 if kd == 4 else 6
                            assert all((v_ & 0x7FFF) >= nraw for v_ in ents[nread:])      # not read by the kernel: padding
                            acc_ = gather(ents[:nread])
                            col = x[0] & 0xFF
                            if kd == 7 and x[0] & plan6.D_CIN:
                                assert carry[sb] is not None and carry[sb][0] == col
                                acc_ = acc_ + carry[sb][1]
                            else:
                                assert carry[sb] is None
                            if not x[0] & plan6.D_VALID:
                                assert not acc_.any() and col == 0 and not x[0] & (plan6.D_CIN | plan6.D_COUT)
                                continue
                            if kd == 7 and x[0] & plan6.D_COUT:
                                carry[sb] = (col, acc_)
                            else:
                                assert not x[0] & plan6.D_COUT
                                carry[sb] = None
                                put_el(sb, col, acc_)
                        elif kd == 2:
                            cols = [x[0] & 0xFF, (x[0] >> 8) & 0xFF, (x[0] >> 16) & 0xFF]
                            assert x[0] >> 24 == 0
                            for j in range(3):
                                acc_ = gather(ents_of([x[1 + j]]))
                                if cols[j]:
                                    put_el(sb, cols[j], acc_)
                                else:
                                    assert not acc_.any()
                        elif kd == 1:
                            cols = [x[0] & 0xFF, (x[0] >> 8) & 0xFF, (x[0] >> 16) & 0xFF, x[0] >> 24, x[1] & 0xFF]
                            assert (x[1] >> 8) & 0xFF == 0
                            ents = [x[1] >> 16] + ents_of(x[2:4])
                            for j in range(5):
                                acc_ = gather([ents[j]])
                                if cols[j]:
                                    put_el(sb, cols[j], acc_)
                                else:
                                    assert not acc_.any()
                        else:
                            for j in range(16):
                                col = (x[j // 4] >> (8 * (j % 4))) & 0xFF
                                if col:
                                    put_el(sb, col, np.zeros(n))
            assert all(c_ is None for c_ in carry)
        for r_ in range(int(hdr[wp, 7])):
            r = get(wp)
            for sb in range(nsub):
                col = r[sb][0]
                if col == 0:
                    continue
                eidx = col * nsp
                assert np.isnan(jac[:, eidx]).all()
                pj, qj = colfac[col]
                jac[:, eidx] = pj * (-wt * ET[:, col - 1] + (-wt * HA)) + qj * (-wt * HB) \
                    + XT * (cp[:, col - 1] - cp[:, last])
        assert pos[wp] <= int(hdr[wp, 1]) * (chb // 4) and pos[wp] > (int(hdr[wp, 1]) - 1) * (chb // 4)
    s0 = -wdcp / cp_avg * H1 + SCP + HT * rho
    jac[:, 0] = -s0 / (rho * cp_avg)
    assert not np.isnan(jac).any()
    dydt = np.empty((n, nsp))
    dydt[:, 0] = -1.0 / (rho * cp_avg) * H1
    dydt[:, 1:] = wdot[:, :last] * w[None, :last] / rho[:, None]
    return dict(conc=v['conc'][:, :nsp], fwd=v['fwd'], rev=v['rev'][:, :v['nrev']], pres_mod=v['pres_mod'][:, :v['npd']],
                spec_rates=wdot, dydt=dydt, jac=jac)


def factored_jvp(fac, v, nsp, rows, cols, ca, cb):
    """J v per state from the factored records without forming J (numpy statement of csrc/consumer.cuh k_jvp; test
    reference only)."""
    fac, v = np.asarray(fac), np.asarray(v)
    last = nsp - 1
    out = np.empty_like(v)
    out[:, 0] = (fac[:, :nsp] * v).sum(axis=1)
    sa, sb = (ca[None, 1:] * v[:, 1:]).sum(axis=1), (cb[None, 1:] * v[:, 1:]).sum(axis=1)
    out[:, 1:] = fac[:, nsp:nsp + last] * v[:, :1] + fac[:, nsp + last:nsp + 2 * last] * sa[:, None] \
        + fac[:, nsp + 2 * last:nsp + 3 * last] * sb[:, None]
    np.add.at(out, (slice(None), rows), fac[:, nsp + 3 * last:] * v[:, cols])
    return out
