/* TEST INFRASTRUCTURE -- CPU restatement of pyJac's emitted-library hot path.
 *
 * This file is the parity ORACLE.  It is never linked into, imported by or called from
 * the shipped product path (pyjac_b200/); only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may execute it.
 *
 * It restates, data-driven, the arithmetic that the reference generator unrolls into C,
 * in the emitted code's own evaluation order (left-to-right, no FMA contraction:
 * compile with -std=c99 / -ffp-contract=off exactly like the reference, libgen.py:43).
 * All literal constants come from oracle/ref_tables.py, which quantises them with the
 * generator's format strings.  Pinned against the reference's own generated C (see
 * tests/test_oracle_pinning.py and tests/golden/).
 *
 * Citations: rs = pyjac/core/rate_subs.py, cj = pyjac/core/create_jacobian.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { F_REV = 1, F_THD = 2, F_PDEP = 4, F_LOW = 8, F_TROE = 16, F_SRI = 32, F_EFF = 64,
       F_PDEPSP_TRUTHY = 128, F_NO_T = 256, F_TROE_T2 = 512, F_SRI5 = 1024, F_SRI5_DT = 2048,
       F_PMT = 4096, F_PMT_IN_JTEMP = 8192, F_HAS_DBDT = 16384, F_KCJ_PREF = 32768, F_PLOG = 65536, F_CHEB = 131072 };

typedef struct {
    int nsp, nr, nrev, npd;
    double ru8;
    const double *sp_mw, *sp_mw_inv, *sp_mw8, *sp_ru_mw, *sp_tmid, *sp_lo, *sp_hi;
    const double *sp_h_lo, *sp_h_hi, *sp_db_lo, *sp_db_hi, *sp_dcp_lo, *sp_dcp_hi;
    const int *dcp_off, *dcp_sp; const double *dcp_tmid; int n_dcp;
    const int *sp_seen, *sp_dbdt_flag;
    const double *j_mwfrac, *j_cj;
    const int *rx_flags, *rx_pdep_sp, *rx_rev_idx, *rx_pm_idx;
    const int *reac_off, *reac_sp, *reac_nu, *prod_off, *prod_sp, *prod_nu;
    const int *net_off, *net_sp, *net_nu, *db_off, *db_sp, *db_nu;
    const int *eff_off, *eff_sp; const double *eff_alpha;
    const int *pr_mode, *pr_off, *pr_sp; const double *pr_coef;
    const int *kcr_off, *kcj_off;
    const double *kcr_tmid, *kcr_lo, *kcr_hi, *kcr_pref, *kcj_tmid, *kcj_lo, *kcj_hi, *kcj_pref;
    const double *arr_main, *arr_k0, *arr_kinf, *arr_ratio;
    const double *troe_pm, *troe_j, *sri_pm, *sri_j, *rx_dt, *rx_pdt, *rx_drdy;
    const int *alpha_mode; const double *alpha_val;
    const int *cheb_off, *cheb_dim;
    const double *cheb_red, *cheb_c8, *cheb_c16;
    const int *plog_off;
    const double *plog_p4, *plog_arr, *plog_lp, *plog_dlp, *plog_dt, *plog_mid;
    const double *consts;
    void* blob;
} OracleMech;

/* ---- PJB200T1 container (pyjac_b200/blob.py) ---- */
typedef struct { char name[24]; int32_t dtype, pad; int64_t count, offset; } Entry;

static const void* find(const void* blob, const char* name, int64_t* count)
{
    const char* b = (const char*)blob;
    int64_t n = *(const int64_t*)(b + 8);
    const Entry* e = (const Entry*)(b + 16);
    for (int64_t k = 0; k < n; ++k)
        if (strncmp(e[k].name, name, 24) == 0) {
            if (count) *count = e[k].count;
            return b + e[k].offset;
        }
    return NULL;
}

OracleMech* oracle_load(const void* src, size_t len)
{
    if (len < 16 || memcmp(src, "PJB200T1", 8) != 0) return NULL;
    OracleMech* m = (OracleMech*)calloc(1, sizeof(OracleMech));
    m->blob = malloc(len);
    memcpy(m->blob, src, len);
    const void* b = m->blob;
    int64_t cnt;
    const int* dims = (const int*)find(b, "dims", NULL);
    m->nsp = dims[0]; m->nr = dims[1]; m->nrev = dims[2]; m->npd = dims[3];
    m->ru8 = *(const double*)find(b, "ru8", NULL);
#define D(field, name) m->field = (const double*)find(b, name, NULL)
#define I(field, name) m->field = (const int*)find(b, name, NULL)
    D(sp_mw, "sp_mw"); D(sp_mw_inv, "sp_mw_inv"); D(sp_mw8, "sp_mw8"); D(sp_ru_mw, "sp_ru_mw");
    D(sp_tmid, "sp_tmid"); D(sp_lo, "sp_lo"); D(sp_hi, "sp_hi");
    D(sp_h_lo, "sp_h_lo"); D(sp_h_hi, "sp_h_hi"); D(sp_db_lo, "sp_db_lo"); D(sp_db_hi, "sp_db_hi");
    D(sp_dcp_lo, "sp_dcp_lo"); D(sp_dcp_hi, "sp_dcp_hi");
    I(dcp_off, "dcp_off"); I(dcp_sp, "dcp_sp");
    m->dcp_tmid = (const double*)find(b, "dcp_tmid", &cnt); m->n_dcp = (int)cnt;
    I(sp_seen, "sp_seen"); I(sp_dbdt_flag, "sp_dbdt_flag");
    D(j_mwfrac, "j_mwfrac"); D(j_cj, "j_cj");
    I(rx_flags, "rx_flags"); I(rx_pdep_sp, "rx_pdep_sp"); I(rx_rev_idx, "rx_rev_idx"); I(rx_pm_idx, "rx_pm_idx");
    I(reac_off, "rx_reac_off"); I(reac_sp, "rx_reac_sp"); I(reac_nu, "rx_reac_nu");
    I(prod_off, "rx_prod_off"); I(prod_sp, "rx_prod_sp"); I(prod_nu, "rx_prod_nu");
    I(net_off, "rx_net_off"); I(net_sp, "rx_net_sp"); I(net_nu, "rx_net_nu");
    I(db_off, "rx_db_off"); I(db_sp, "rx_db_sp"); I(db_nu, "rx_db_nu");
    I(eff_off, "rx_eff_off"); I(eff_sp, "rx_eff_sp"); D(eff_alpha, "rx_eff_alpha");
    I(pr_mode, "rx_pr_mode"); I(pr_off, "rx_pr_off"); I(pr_sp, "rx_pr_sp"); D(pr_coef, "rx_pr_coef");
    I(kcr_off, "kcr_off"); D(kcr_tmid, "kcr_tmid"); D(kcr_lo, "kcr_lo"); D(kcr_hi, "kcr_hi"); D(kcr_pref, "kcr_pref");
    I(kcj_off, "kcj_off"); D(kcj_tmid, "kcj_tmid"); D(kcj_lo, "kcj_lo"); D(kcj_hi, "kcj_hi"); D(kcj_pref, "kcj_pref");
    D(arr_main, "arr_main"); D(arr_k0, "arr_k0"); D(arr_kinf, "arr_kinf"); D(arr_ratio, "arr_ratio");
    D(troe_pm, "troe_pm"); D(troe_j, "troe_j"); D(sri_pm, "sri_pm"); D(sri_j, "sri_j");
    D(rx_dt, "rx_dt"); D(rx_pdt, "rx_pdt"); D(rx_drdy, "rx_drdy");
    I(alpha_mode, "alpha_mode"); D(alpha_val, "alpha_val"); D(consts, "consts");
    I(cheb_off, "cheb_off"); I(cheb_dim, "cheb_dim"); D(cheb_red, "cheb_red"); D(cheb_c8, "cheb_c8"); D(cheb_c16, "cheb_c16");
    I(plog_off, "plog_off"); D(plog_p4, "plog_p4"); D(plog_arr, "plog_arr"); D(plog_lp, "plog_lp");
    D(plog_dlp, "plog_dlp"); D(plog_dt, "plog_dt"); D(plog_mid, "plog_mid");
#undef D
#undef I
    return m;
}

void oracle_free(OracleMech* m) { if (m) { free(m->blob); free(m); } }
int oracle_nsp(const OracleMech* m) { return m->nsp; }
int oracle_nr(const OracleMech* m) { return m->nr; }
int oracle_nrev(const OracleMech* m) { return m->nrev; }
int oracle_npd(const OracleMech* m) { return m->npd; }

/* rs:27-146: forms 0-3 for A > 0 (a[1] = log A), 4-7 for A < 0 (rs:108-141, a[1] = A) */
static double arrhenius(const double* a, double T, double logT)
{
    switch ((int)a[0]) {
    case 0: return a[1];
    case 1: return exp(a[1] + a[2] * logT);
    case 2: return exp(a[1] - (a[3] / T));
    case 3: return exp(a[1] + a[2] * logT - (a[3] / T));
    case 4: {
        double k = a[1];
        for (int i = 0; i < (int)a[2]; ++i) k = k * T;
        return k;
    }
    case 5: return a[1] * exp(a[2] * logT);
    case 6: return a[1] * exp(-(a[3] / T));
    default: return a[1] * exp(a[2] * logT - (a[3] / T));
    }
}

/* PLOG pressure range of reaction i (rs:598-632): -1 below the first pressure, e in
 * [o0, o1 - 1) between pressures e and e + 1, o1 - 1 above the last; -2 if no branch of the
 * emitted if / else-if chain is taken (NaN pressure) */
static int plog_range(const OracleMech* m, int i, double pres)
{
    const int o0 = m->plog_off[i], o1 = m->plog_off[i + 1];
    if (pres <= m->plog_p4[o0]) return -1;
    for (int e = o0; e + 1 < o1; ++e)
        if ((pres > m->plog_p4[e]) && (pres <= m->plog_p4[e + 1])) return e;
    if (pres > m->plog_p4[o1 - 1]) return o1 - 1;
    return -2;
}

/* kf of a PLOG reaction as emitted (rs:598-632, cj:293-327); `kf` keeps its old value when
 * no branch is taken */
static double plog_kf(const OracleMech* m, int i, double T, double logT, double pres, double kf)
{
    const int o0 = m->plog_off[i], o1 = m->plog_off[i + 1];
    const int e = plog_range(m, i, pres);
    if (e == -1) return arrhenius(m->plog_arr + 4 * o0, T, logT);
    if (e == o1 - 1) return arrhenius(m->plog_arr + 4 * (o1 - 1), T, logT);
    if (e < 0) return kf;
    kf = log(arrhenius(m->plog_arr + 4 * e, T, logT));
    double kf2 = log(arrhenius(m->plog_arr + 4 * (e + 1), T, logT));
    return exp(kf + (kf2 - kf) * (log(pres) - m->plog_lp[e]) / m->plog_dlp[e]);
}

/* get_cheb_rate (rs:149-251): kf of a Chebyshev reaction from the reduced temperature and
 * pressure, in the emitted statement order */
static double cheb_kf(const OracleMech* m, int i, double Tred, double Pred)
{
    const int nt = m->cheb_dim[2 * i], np = m->cheb_dim[2 * i + 1];
    const double* c = m->cheb_c8 + m->cheb_off[i];
    double dot_prod[64];
    double cheb_temp_0 = 1, cheb_temp_1 = Pred, kf;
    for (int r = 0; r < nt; ++r) dot_prod[r] = c[r * np] + Pred * c[r * np + 1];
    int update_one = 1;
    for (int j = 2; j < np; ++j) {
        if (update_one) {
            cheb_temp_0 = 2 * Pred * cheb_temp_1 - cheb_temp_0;
            for (int r = 0; r < nt; ++r) dot_prod[r] += c[r * np + j] * cheb_temp_0;
        } else {
            cheb_temp_1 = 2 * Pred * cheb_temp_0 - cheb_temp_1;
            for (int r = 0; r < nt; ++r) dot_prod[r] += c[r * np + j] * cheb_temp_1;
        }
        update_one = !update_one;
    }
    cheb_temp_0 = 1;
    cheb_temp_1 = Tred;
    kf = dot_prod[0] + Tred * dot_prod[1];
    update_one = 1;
    for (int r = 2; r < nt; ++r) {
        if (update_one) {
            cheb_temp_0 = 2 * Tred * cheb_temp_1 - cheb_temp_0;
            kf += dot_prod[r] * cheb_temp_0;
        } else {
            cheb_temp_1 = 2 * Tred * cheb_temp_0 - cheb_temp_1;
            kf += dot_prod[r] * cheb_temp_1;
        }
        update_one = !update_one;
    }
    return pow(10.0, kf);
}

/* write_cheb_ut (cj:1532-1606): the sum over second-kind polynomials of the temperature derivative */
static double cheb_ut(const OracleMech* m, int i, double Tred, double Pred)
{
    const int nt = m->cheb_dim[2 * i], np = m->cheb_dim[2 * i + 1];
    const double* c = m->cheb_c16 + m->cheb_off[i];
    double dot_prod[64];
    double cheb_temp_0 = 1, cheb_temp_1 = Pred, kf;
    for (int r = 1; r < nt; ++r) dot_prod[r] = c[r * np] + Pred * c[r * np + 1];
    int update_one = 1;
    for (int j = 2; j < np; ++j) {
        if (update_one) {
            cheb_temp_0 = 2 * Pred * cheb_temp_1 - cheb_temp_0;
            for (int r = 1; r < nt; ++r) dot_prod[r] += c[r * np + j] * cheb_temp_0;
        } else {
            cheb_temp_1 = 2 * Pred * cheb_temp_0 - cheb_temp_1;
            for (int r = 1; r < nt; ++r) dot_prod[r] += c[r * np + j] * cheb_temp_1;
        }
        update_one = !update_one;
    }
    cheb_temp_0 = 1.0;
    cheb_temp_1 = 2.0 * Tred;
    kf = dot_prod[1] + 2.0 * Tred * dot_prod[2];
    update_one = 1;
    for (int r = 3; r < nt; ++r) {
        if (update_one) {
            cheb_temp_0 = 2.0 * Tred * cheb_temp_1 - cheb_temp_0;
            kf += dot_prod[r] * cheb_temp_0;
        } else {
            cheb_temp_1 = 2.0 * Tred * cheb_temp_0 - cheb_temp_1;
            kf += dot_prod[r] * cheb_temp_1;
        }
        update_one = !update_one;
    }
    return kf;
}

/* C[a]*C[a]*C[b]*...: returns the left-assoc product with `tail` multiplied last.
 * rs:636-658, 815-840.  first==1 means the product starts with the first factor. */
static double conc_prod_times(const int* sp, const int* nu, int n, const double* C, double tail)
{
    int first = 1;
    double p = 0.0;
    for (int k = 0; k < n; ++k)
        for (int r = 0; r < nu[k]; ++r) {
            if (first) { p = C[sp[k]]; first = 0; } else p = p * C[sp[k]];
        }
    return first ? tail : p * tail;
}

/* rs:652-655: an irreversible reaction has its rate-constant expression printed in line after the
 * concentrations, `C[a] * C[b] * <expr>`; for A < 0 <expr> is itself a product (rs:108-141) and C
 * evaluates the whole line left to right */
static double conc_prod_times_inline(const int* sp, const int* nu, int n, const double* C,
                                     const double* a, double T, double logT)
{
    int form = (int)a[0];
    int first = 1;
    double p = 0.0;
    if (form < 4) return conc_prod_times(sp, nu, n, C, arrhenius(a, T, logT));
    for (int k = 0; k < n; ++k)
        for (int r = 0; r < nu[k]; ++r) {
            if (first) { p = C[sp[k]]; first = 0; } else p = p * C[sp[k]];
        }
    if (first) return arrhenius(a, T, logT);
    p = p * a[1];
    if (form == 4) {
        for (int i = 0; i < (int)a[2]; ++i) p = p * T;
        return p;
    }
    if (form == 5) return p * exp(a[2] * logT);
    if (form == 6) return p * exp(-(a[3] / T));
    return p * exp(a[2] * logT - (a[3] / T));
}

static double kc_exponent(const int* off, const double* tmid, const double* lo, const double* hi,
                          int i, double T, double logT)
{
    double Kc = 0.0;
    for (int bkt = off[i]; bkt < off[i + 1]; ++bkt) {
        const double* c = (T <= tmid[bkt]) ? lo + 7 * bkt : hi + 7 * bkt;
        double v = (c[0] + c[1] * logT + T * (c[2] + T * (c[3] + T * (c[4] + c[5] * T))) - c[6] / T);
        if (bkt == off[i]) Kc = v; else Kc += v;
    }
    return Kc;
}

/* rs:1626-1706 */
void oracle_eval_conc(const OracleMech* m, double T, double pres, const double* y,
                      double* y_N, double* mw_avg, double* rho, double* conc)
{
    int n = m->nsp;
    double s = 0.0;
    for (int k = 0; k < n - 1; ++k) s = (k == 0) ? y[0] : s + y[k];
    *y_N = 1.0 - (s);
    double w = 0.0;
    for (int k = 0; k < n - 1; ++k) {
        double t = (y[k] * m->sp_mw_inv[k]);
        w = (k == 0) ? t : w + t;
    }
    if (n > 1) w = w + ((*y_N) * m->sp_mw_inv[n - 1]); else w = ((*y_N) * m->sp_mw_inv[n - 1]);
    *mw_avg = 1.0 / w;
    *rho = pres * (*mw_avg) / (m->ru8 * T);
    for (int k = 0; k < n - 1; ++k) conc[k] = (*rho) * y[k] * m->sp_mw_inv[k];
    conc[n - 1] = (*rho) * (*y_N) * m->sp_mw_inv[n - 1];
}

/* rs:563-842 */
void oracle_eval_rxn_rates(const OracleMech* m, double T, double pres, const double* C,
                           double* fwd, double* rev)
{
    double logT = log(T);
    double kf = 0.0;
    for (int i = 0; i < m->nr; ++i) {
        if (m->rx_flags[i] & F_PLOG) kf = plog_kf(m, i, T, logT, pres, kf);
        else if (m->rx_flags[i] & F_CHEB) {
            const double* cr = m->cheb_red + 12 * i;
            double Tred = ((2.0 / T) - cr[0]) / cr[1];
            double Pred = (2.0 * log10(pres) - cr[2]) / cr[3];
            kf = cheb_kf(m, i, Tred, Pred);
        } else if (!(m->rx_flags[i] & F_REV)) {
            fwd[i] = conc_prod_times_inline(m->reac_sp + m->reac_off[i], m->reac_nu + m->reac_off[i],
                                            m->reac_off[i + 1] - m->reac_off[i], C,
                                            m->arr_main + 4 * i, T, logT);
            continue;
        } else kf = arrhenius(m->arr_main + 4 * i, T, logT);
        fwd[i] = conc_prod_times(m->reac_sp + m->reac_off[i], m->reac_nu + m->reac_off[i],
                                 m->reac_off[i + 1] - m->reac_off[i], C, kf);
        if (m->rx_flags[i] & F_REV) {
            double Kc = kc_exponent(m->kcr_off, m->kcr_tmid, m->kcr_lo, m->kcr_hi, i, T, logT);
            Kc = m->kcr_pref[i] * exp(Kc);
            rev[m->rx_rev_idx[i]] =
                conc_prod_times(m->prod_sp + m->prod_off[i], m->prod_nu + m->prod_off[i],
                                m->prod_off[i + 1] - m->prod_off[i], C, kf) / Kc;
        }
    }
}

/* m + (a-1)*C ... (rs:1122-1148) */
static double third_body(const OracleMech* m, int i, double mm, const double* C)
{
    double thd = mm;
    for (int e = m->eff_off[i]; e < m->eff_off[i + 1]; ++e) {
        double a = m->eff_alpha[e];
        if (a == 1.0) continue;
        thd = thd + (a - 1.0) * C[m->eff_sp[e]];
    }
    return thd;
}

static double lg10c(double x) { return log10(fmax(x, 1.0e-300)); }

/* rs:1076-1283 */
void oracle_get_rxn_pres_mod(const OracleMech* m, double T, double pres, const double* C,
                             double* pres_mod)
{
    double logT = log(T);
    double mm = pres / (m->ru8 * T);
    for (int i = 0; i < m->nr; ++i) {
        int fl = m->rx_flags[i];
        if (!(fl & (F_THD | F_PDEP))) continue;
        int p = m->rx_pm_idx[i];
        if (fl & F_THD) pres_mod[p] = third_body(m, i, mm, C);
        if (fl & F_PDEP) {
            double thd = 0.0;
            if (m->rx_pdep_sp[i] < 0) thd = third_body(m, i, mm, C);
            double k0 = arrhenius(m->arr_k0 + 4 * i, T, logT);
            double kinf = arrhenius(m->arr_kinf + 4 * i, T, logT);
            double Pr = (m->rx_pdep_sp[i] >= 0) ? k0 * C[m->rx_pdep_sp[i]] / kinf : k0 * thd / kinf;
            double val;
            if (fl & F_TROE) {
                const double* t = m->troe_pm + 8 * i;
                double f = t[0] * ((t[2] > 0) ? exp(-T / t[1]) : exp(T / t[1]))
                         + t[3] * ((t[5] > 0) ? exp(-T / t[4]) : exp(T / t[4]));
                if (fl & F_TROE_T2) f = f + ((t[7] > 0) ? exp(-t[6] / T) : exp(t[6] / T));
                double logFcent = log10(fmax(f, 1.0e-300));
                double A = lg10c(Pr) - 0.67 * logFcent - 0.4;
                double B = 0.806 - 1.1762 * logFcent - 0.14 * lg10c(Pr);
                val = pow(10.0, logFcent / (1.0 + A * A / (B * B)));
                val = (fl & F_LOW) ? val * Pr / (1.0 + Pr) : val / (1.0 + Pr);
            } else if (fl & F_SRI) {
                const double* s = m->sri_pm + 8 * i;
                double X = 1.0 / (1.0 + lg10c(Pr) * lg10c(Pr));
                val = pow(s[0] * ((s[2] > 0) ? exp(-s[1] / T) : exp(s[1] / T))
                          + ((s[4] > 0) ? exp(-T / s[3]) : exp(T / s[3])), X);
                if (fl & F_SRI5) val = val * s[5] * pow(T, s[6]);
                val = (fl & F_LOW) ? val * Pr / (1.0 + Pr) : val / (1.0 + Pr);
            } else {
                val = (fl & F_LOW) ? Pr / (1.0 + Pr) : 1.0 / (1.0 + Pr);
            }
            pres_mod[p] = val;
        }
    }
}

/* rs:1425-1527.  sp_rates has NSP-1 entries; the last species goes to *dy_N. */
void oracle_eval_spec_rates(const OracleMech* m, const double* fwd, const double* rev,
                            const double* pres_mod, double* sp_rates, double* dy_N)
{
    int n = m->nsp;
    char* first = (char*)malloc(n);
    memset(first, 1, n);
    for (int i = 0; i < m->nr; ++i) {
        int fl = m->rx_flags[i];
        for (int e = m->net_off[i]; e < m->net_off[i + 1]; ++e) {
            int k = m->net_sp[e];
            int nu = m->net_nu[e];
            double r = (fl & F_REV) ? (fwd[i] - rev[m->rx_rev_idx[i]]) : fwd[i];
            double term;
            int a = nu < 0 ? -nu : nu;
            if (fl & (F_THD | F_PDEP)) {
                double pm = pres_mod[m->rx_pm_idx[i]];
                term = (a != 1) ? (double)a * r * pm : r * pm;
            } else {
                term = (a != 1) ? (double)a * r : r;
            }
            double* dst = (k == n - 1) ? dy_N : &sp_rates[k];
            if (first[k]) { *dst = (nu < 0) ? -term : term; first[k] = 0; }
            else if (nu < 0) *dst -= term; else *dst += term;
        }
    }
    for (int k = 0; k < n; ++k)
        if (first[k]) { if (k == n - 1) *dy_N = 0.0; else sp_rates[k] = 0.0; }
    free(first);
}

/* rs:1806-1874 */
void oracle_eval_h(const OracleMech* m, double T, double* h)
{
    for (int k = 0; k < m->nsp; ++k) {
        const double* c = (T <= m->sp_tmid[k]) ? m->sp_h_lo + 6 * k : m->sp_h_hi + 6 * k;
        h[k] = m->sp_ru_mw[k] * (c[0] + T * (c[1] + T * (c[2] + T * (c[3] + T * (c[4] + c[5] * T)))));
    }
}

/* rs:2021-2086 */
void oracle_eval_cp(const OracleMech* m, double T, double* cp)
{
    for (int k = 0; k < m->nsp; ++k) {
        const double* a = (T <= m->sp_tmid[k]) ? m->sp_lo + 7 * k : m->sp_hi + 7 * k;
        cp[k] = m->sp_ru_mw[k] * (a[0] + T * (a[1] + T * (a[2] + T * (a[3] + a[4] * T))));
    }
}

/* rs:2171-2335 (CONP).  y = [T, Y_0..Y_{NSP-2}], dy has NSP entries. */
void oracle_dydt(const OracleMech* m, double t, double pres, const double* y, double* dy)
{
    (void)t;
    int n = m->nsp;
    double* w = (double*)malloc(sizeof(double) * (3 * (size_t)n + 2 * (size_t)m->nr + m->npd + 4));
    double *conc = w, *cp = w + n, *h = w + 2 * n, *fwd = w + 3 * n, *rev = fwd + m->nr,
           *pm = rev + m->nr;
    double y_N, mw_avg, rho, dy_N;
    oracle_eval_conc(m, y[0], pres, &y[1], &y_N, &mw_avg, &rho, conc);
    oracle_eval_rxn_rates(m, y[0], pres, conc, fwd, rev);
    oracle_get_rxn_pres_mod(m, y[0], pres, conc, pm);
    oracle_eval_spec_rates(m, fwd, rev, pm, &dy[1], &dy_N);
    oracle_eval_cp(m, y[0], cp);
    double cp_avg = 0.0;
    for (int k = 0; k < n - 1; ++k) {
        double v = (cp[k] * y[k + 1]);
        cp_avg = (k == 0) ? v : cp_avg + v;
    }
    cp_avg = (n > 1) ? cp_avg + (cp[n - 1] * y_N) : (cp[n - 1] * y_N);
    oracle_eval_h(m, y[0], h);
    double s = 0.0;
    int first = 1;
    for (int k = 0; k < n; ++k) {
        if (!m->sp_seen[k]) continue;
        double v = ((k < n - 1 ? dy[k + 1] : dy_N) * h[k] * m->sp_mw[k]);
        if (first) { s = v; first = 0; } else s = s + v;
    }
    dy[0] = (-1.0 / (rho * cp_avg)) * (s);
    for (int k = 0; k < n - 1; ++k) dy[k + 1] *= (m->sp_mw[k] / rho);
    free(w);
}

/* rs:1708-1800: concentrations from the density; the pressure follows */
void oracle_eval_conc_rho(const OracleMech* m, double T, double rho, const double* y,
                          double* y_N, double* mw_avg, double* pres, double* conc)
{
    int n = m->nsp;
    double s = 0.0;
    for (int k = 0; k < n - 1; ++k) s = (k == 0) ? y[0] : s + y[k];
    *y_N = 1.0 - (s);
    double w = 0.0;
    for (int k = 0; k < n - 1; ++k) {
        double t = (y[k] * m->sp_mw_inv[k]);
        w = (k == 0) ? t : w + t;
    }
    if (n > 1) w = w + ((*y_N) * m->sp_mw_inv[n - 1]); else w = ((*y_N) * m->sp_mw_inv[n - 1]);
    *mw_avg = 1.0 / w;
    *pres = rho * m->ru8 * T / (*mw_avg);
    for (int k = 0; k < n - 1; ++k) conc[k] = rho * y[k] * m->sp_mw_inv[k];
    conc[n - 1] = rho * (*y_N) * m->sp_mw_inv[n - 1];
}

/* rs:1876-1945 */
void oracle_eval_u(const OracleMech* m, double T, double* u)
{
    for (int k = 0; k < m->nsp; ++k) {
        const double* c = (T <= m->sp_tmid[k]) ? m->sp_h_lo + 6 * k : m->sp_h_hi + 6 * k;
        u[k] = m->sp_ru_mw[k] * (c[0] + T * (c[1] - 1.0 + T * (c[2] + T * (c[3] + T * (c[4] + c[5] * T)))));
    }
}

/* rs:1947-2019 */
void oracle_eval_cv(const OracleMech* m, double T, double* cv)
{
    for (int k = 0; k < m->nsp; ++k) {
        const double* a = (T <= m->sp_tmid[k]) ? m->sp_lo + 7 * k : m->sp_hi + 7 * k;
        cv[k] = m->sp_ru_mw[k] * (a[0] - 1.0 + T * (a[1] + T * (a[2] + T * (a[3] + a[4] * T))));
    }
}

/* rs:2340-2485 (CONV, compiled out upstream by header.h's #define CONP, mech_auxiliary.py:464-466): the second
 * argument is the density */
void oracle_dydt_conv(const OracleMech* m, double t, double rho, const double* y, double* dy)
{
    (void)t;
    int n = m->nsp;
    double* w = (double*)malloc(sizeof(double) * (3 * (size_t)n + 2 * (size_t)m->nr + m->npd + 4));
    double *conc = w, *cv = w + n, *u = w + 2 * n, *fwd = w + 3 * n, *rev = fwd + m->nr,
           *pm = rev + m->nr;
    double y_N, mw_avg, pres, dy_N;
    oracle_eval_conc_rho(m, y[0], rho, &y[1], &y_N, &mw_avg, &pres, conc);
    oracle_eval_rxn_rates(m, y[0], pres, conc, fwd, rev);
    oracle_get_rxn_pres_mod(m, y[0], pres, conc, pm);
    oracle_eval_spec_rates(m, fwd, rev, pm, &dy[1], &dy_N);
    oracle_eval_cv(m, y[0], cv);
    double cv_avg = 0.0;
    for (int k = 0; k < n - 1; ++k) {
        double v = (cv[k] * y[k + 1]);
        cv_avg = (k == 0) ? v : cv_avg + v;
    }
    cv_avg = (n > 1) ? cv_avg + (cv[n - 1] * y_N) : (cv[n - 1] * y_N);
    oracle_eval_u(m, y[0], u);
    double s = 0.0;
    int first = 1;
    for (int k = 0; k < n; ++k) {
        if (!m->sp_seen[k]) continue;
        double v = ((k < n - 1 ? dy[k + 1] : dy_N) * u[k] * m->sp_mw[k]);
        if (first) { s = v; first = 0; } else s = s + v;
    }
    dy[0] = (-1.0 / (rho * cv_avg)) * (s);
    for (int k = 0; k < n - 1; ++k) dy[k + 1] *= (m->sp_mw[k] / rho);
    free(w);
}

/* cj:2189-3298.  jac is column-major NSP x NSP; entries the generated code never
 * assigns are left untouched (caller pre-zeroes, as the reference requires). */
void oracle_eval_jacob(const OracleMech* m, double t, double pres, const double* y, double* jac)
{
    (void)t;
    const int n = m->nsp, nr = m->nr;
    double T = y[0];
    size_t nw = 6 * (size_t)n + 2 * (size_t)nr + m->npd + 8;
    double* w = (double*)calloc(nw, sizeof(double));
    double *conc = w, *spec_rates = w + n, *dBdT = w + 2 * n, *h = w + 3 * n, *cp = w + 4 * n,
           *J_nplusjplus = w + 5 * n, *fwd_rates = w + 6 * n, *rev_rates = fwd_rates + nr,
           *pres_mod = rev_rates + nr;
    char* touched = (char*)calloc((size_t)n * n + n, 1);
    char* Jnpj_touched = touched + (size_t)n * n;
    int J_nplusone_touched = 0;
    double J_nplusone = 0;

    double mw_avg, rho, y_N;
    oracle_eval_conc(m, y[0], pres, &y[1], &y_N, &mw_avg, &rho, conc);
    oracle_eval_rxn_rates(m, T, pres, conc, fwd_rates, rev_rates);
    oracle_get_rxn_pres_mod(m, T, pres, conc, pres_mod);
    oracle_eval_spec_rates(m, fwd_rates, rev_rates, pres_mod, spec_rates, &spec_rates[n - 1]);

    double mm = pres / (m->ru8 * T);
    double conc_temp = 0.0;
    double logT = log(T);
    double j_temp = 0.0, kf = 0.0, pres_mod_temp = 0.0, Kc = 0.0, kr = 0, Pr = 0.0;
    double Fcent = 0.0, A = 0.0, B = 0.0, lnF_AB = 0.0, X = 0.0;
    double Tred = 0.0, Pred = 0.0;
    double rho_inv = 1.0 / rho;
    const double* K = m->consts;

    /* cj:761-865 */
    for (int k = 0; k < n; ++k) {
        if (!m->sp_dbdt_flag[k]) continue;
        const double* c = (T <= m->sp_tmid[k]) ? m->sp_db_lo + 6 * k : m->sp_db_hi + 6 * k;
        dBdT[k] = (c[0] + c[1] / T) / T + c[2] + T * (c[3] + T * (c[4] + c[5] * T));
    }

    for (int i = 0; i < nr; ++i) {
        const int fl = m->rx_flags[i];
        const int rev = fl & F_REV;
        const int p = m->rx_pm_idx[i];
        const int ri = m->rx_rev_idx[i];
        const double f = fwd_rates[i];
        const double r = rev ? rev_rates[ri] : 0.0;
        const double* dt = m->rx_dt + 8 * i;

        /* ---------------- partial wrt T (cj:2728-2845) */
        if (fl & F_PDEP) {
            /* write_pr (cj:986-1061) */
            int mode = m->pr_mode[i];
            int o0 = m->pr_off[i], o1 = m->pr_off[i + 1];
            if (mode == 1) conc_temp = conc[m->pr_sp[o0]];
            else if (mode == 2) conc_temp = mm;
            else if (mode == 3) {
                double v = mm;
                for (int e = o0; e < o1; ++e) v = v + m->pr_coef[e] * conc[m->pr_sp[e]];
                conc_temp = (v);
            } else if (mode == 4) {
                double v = 0.0;
                for (int e = o0; e < o1; ++e) {
                    double tt = m->pr_coef[e] * conc[m->pr_sp[e]];
                    v = (e == o0) ? tt : v + tt;
                }
                conc_temp += (v);
            }
            Pr = conc_temp * (arrhenius(m->arr_ratio + 4 * i, T, logT));
            if (fl & F_TROE) {      /* cj:1083-1111 */
                const double* q = m->troe_j + 8 * i;
                Fcent = q[0] * exp(T / q[1]) + q[2] * exp(T / q[3]);
                if (fl & F_TROE_T2) Fcent = Fcent + exp(q[4] / T);
                A = lg10c(Pr) - 0.67 * lg10c(Fcent) - 0.4;
                B = 0.806 - 1.1762 * lg10c(Fcent) - 0.14 * lg10c(Pr);
                lnF_AB = 2.0 * log(fmax(Fcent, 1.0e-300)) * A /
                         (B * B * B * (1.0 + A * A / (B * B)) * (1.0 + A * A / (B * B)));
            } else if (fl & F_SRI) {
                X = 1.0 / (1.0 + lg10c(Pr) * lg10c(Pr));
            }
        }

        /* PLOG (cj:1687-1850): below the first / above the last pressure the elementary form
         * with that pressure's parameters; between two pressures the interpolated form */
        int plog_e = -3;
        double dtp[8];
        if (fl & F_PLOG) {
            const int o0 = m->plog_off[i], o1 = m->plog_off[i + 1];
            plog_e = plog_range(m, i, pres);
            if (plog_e == -1 || plog_e == o1 - 1) {
                const double* pdt_ = m->plog_dt + 3 * (plog_e == -1 ? o0 : o1 - 1);
                memcpy(dtp, dt, sizeof dtp);
                dtp[0] = pdt_[0]; dtp[1] = pdt_[1]; dtp[2] = pdt_[2];
                dt = dtp;
                plog_e = -3;
            }
        }
        if (fl & F_CHEB) {      /* write_cheb_rxn_dt (cj:1609-1684) */
            const double* cr = m->cheb_red + 12 * i;
            Tred = ((2.0 / T) - cr[4]) / cr[5];
            Pred = (2.0 * log10(pres) - cr[6]) / cr[7];
            kf = cheb_ut(m, i, Tred, Pred);
            double elem = kf * (cr[8] / T) * (rev ? (f - r) : f);
            if (dt[4] != 0.0) elem = elem + f * dt[3];
            if (rev) {
                double sdb = 0.0;
                int o0 = m->db_off[i], o1 = m->db_off[i + 1];
                for (int e = o0; e < o1; ++e) {
                    double v = (double)m->db_nu[e] * dBdT[m->db_sp[e]];
                    sdb = (e == o0) ? v : sdb + v;
                }
                double inner = (dt[6] != 0.0) ? dt[5] + -T * (sdb) : -T * (sdb);
                elem = elem - r * (inner);
            }
            j_temp = ((1.0 / T) * (elem)) * rho_inv;
            plog_e = -4;
        }
        if (plog_e >= 0) {
            const double* q = m->plog_mid + 8 * plog_e;
            double dk = 0.0;
            int have = 0;
            if (q[0] != 0.0) { dk = q[1]; have = 1; }
            if (q[2] != 0.0) { dk = have ? dk + q[3] / T : q[3] / T; have = 1; }
            {
                double lp = (q[7] != 0.0) ? (log(pres) - m->plog_lp[plog_e]) : (log(pres));
                double v = ((q[4] != 0.0) ? (q[5] + q[6] / T) : (q[6] / T)) * lp / m->plog_dlp[plog_e];
                dk = have ? dk + v : v;
            }
            double elem = (dk) * (rev ? (f - r) : f);
            if (dt[4] != 0.0) elem = elem + f * dt[3];
            if (rev && ((fl & F_HAS_DBDT) || dt[6] != 0.0)) {
                double inner;
                if (fl & F_HAS_DBDT) {
                    double sdb = 0.0;
                    int o0 = m->db_off[i], o1 = m->db_off[i + 1];
                    for (int e = o0; e < o1; ++e) {
                        double v = (double)m->db_nu[e] * dBdT[m->db_sp[e]];
                        sdb = (e == o0) ? v : sdb + v;
                    }
                    inner = (dt[6] != 0.0) ? dt[5] + -T * (sdb) : -T * (sdb);
                } else inner = dt[5];
                elem = elem - r * (inner);
            }
            j_temp = ((1.0 / T) * (elem)) * rho_inv;
        }
        if (!(fl & F_NO_T)) {
          if (plog_e == -3) {
            /* get_elementary_rxn_dt (cj:1426-1523) */
            double elem = 0.0;
            int dk_form = (int)dt[0];
            double dk = 0.0;
            if (dk_form == 3) dk = dt[1] + (dt[2] / T);
            else if (dk_form == 1) dk = dt[1];
            else if (dk_form == 2) dk = (dt[2] / T);
            if (rev) {
                int have = 0;
                if (dk_form) { elem = (f - r) * (dk); have = 1; }
                if (dt[4] != 0.0) {
                    double v = f * dt[3];
                    elem = have ? elem + v : v;
                    have = 1;
                }
                if ((fl & F_HAS_DBDT) || dt[6] != 0.0) {
                    double inner = 0.0;
                    if (fl & F_HAS_DBDT) {
                        /* left-to-right sum of nu*dBdT; +-1.0*x and a+(-x) are exact */
                        double sdb = 0.0;
                        int o0 = m->db_off[i], o1 = m->db_off[i + 1];
                        for (int e = o0; e < o1; ++e) {
                            double v = (double)m->db_nu[e] * dBdT[m->db_sp[e]];
                            sdb = (e == o0) ? v : sdb + v;
                        }
                        inner = (dt[6] != 0.0) ? dt[5] + -T * (sdb) : -T * (sdb);
                    } else inner = dt[5];
                    double v = r * (inner);
                    elem = have ? elem - v : -v;
                }
            } else {
                double inner;
                if (dk_form && dt[4] != 0.0) inner = dk + dt[3];
                else if (dk_form) inner = dk;
                else inner = +dt[3];
                elem = f * (inner);
            }

            if (fl & F_PDEP) {      /* get_pdep_dt (cj:1159-1191) */
                const double* pd = m->rx_pdt + 4 * i;
                double dpr = (pd[0] + (pd[2] / T) - 1.0);
                double Xd;
                if (fl & F_LOW) Xd = (dpr / (T * (1.0 + Pr)));
                else Xd = (-Pr * dpr / (T * (1.0 + Pr)));
                if (fl & F_TROE) {  /* cj:1262-1292 */
                    const double* q = m->troe_j + 8 * i;
                    double dF = (q[5] * exp(T / q[1]) - q[6] * exp(T / q[3]));
                    if (fl & F_TROE_T2)
                        dF = (q[5] * exp(T / q[1]) - q[6] * exp(T / q[3]) + (q[7] / (T * T)) * exp(q[4] / T));
                    Xd = Xd + (((1.0 / (Fcent * (1.0 + A * A / (B * B)))) - lnF_AB * (-K[3] * B + K[4] * A) / Fcent) * dF)
                         - lnF_AB * (K[5] * B + K[6] * A) * (pd[1] + (pd[2] / T) - 1.0) / T;
                } else if (fl & F_SRI) {  /* cj:1215-1235 */
                    const double* s = m->sri_j + 12 * i;
                    double v = X * ((((s[3] / (T * T)) * exp(s[4] / T) - s[5] * exp(T / s[6])) /
                                     (s[7] * exp(s[4] / T) + exp(T / s[6])) -
                                     X * K[2] * lg10c(Pr) * (pd[1] + (pd[2] / T) - 1.0) *
                                     log(s[7] * exp(s[4] / T) + exp(T / s[6])) / T));
                    Xd = Xd + v;
                    if (fl & F_SRI5_DT) Xd = Xd + (s[8] / T);
                }
                j_temp = (pres_mod[p] * (Xd) * (rev ? (f - r) : f) + (pres_mod[p] / T) * (elem)) * rho_inv;
            } else if (fl & F_THD) {
                j_temp = ((-pres_mod[p] * (rev ? (f - r) : f) / T) + (pres_mod[p] / T) * (elem)) * rho_inv;
            } else {
                j_temp = ((1.0 / T) * (elem)) * rho_inv;
            }
          }

            for (int e = m->net_off[i]; e < m->net_off[i + 1]; ++e) {
                int k = m->net_sp[e], nu = m->net_nu[e];
                double v;
                if (nu == 1) v = j_temp * m->sp_mw[k];
                else if (nu == -1) v = -j_temp * m->sp_mw[k];
                else v = j_temp * (double)nu * m->sp_mw[k];
                if (k + 1 == n) { if (J_nplusone_touched) J_nplusone += v; else J_nplusone = v; J_nplusone_touched = 1; }
                else { if (touched[k + 1]) jac[k + 1] += v; else jac[k + 1] = v; touched[k + 1] = 1; }
            }
        }

        /* ---------------- partial wrt species (cj:2850-2938) */
        if (rev) {
            Kc = kc_exponent(m->kcj_off, m->kcj_tmid, m->kcj_lo, m->kcj_hi, i, T, logT);
            Kc = (fl & F_KCJ_PREF) ? m->kcj_pref[i] * exp(Kc) : exp(Kc);
        }
        /* write_dr_dy (cj:153-269) */
        if (fl & F_PMT) {
            double net = rev ? (f - r) : f;
            if (fl & F_PDEP) {
                double g = (fl & F_LOW) ? (1.0 / (1.0 + Pr)) : (-Pr / (1.0 + Pr));
                if (fl & F_TROE)
                    g = g - log(fmax(Fcent, 1.0e-300)) * 2.0 * A * (B * K[0] + A * K[1]) /
                            (B * B * B * (1.0 + A * A / (B * B)) * (1.0 + A * A / (B * B)));
                else if (fl & F_SRI) {
                    const double* s = m->sri_j + 12 * i;
                    g = g - X * X * K[2] * lg10c(Pr) * log(s[0] * exp(s[1] / T) + exp(T / s[2]));
                }
                pres_mod_temp = (g) * net;
            } else pres_mod_temp = net;
        }
        {
            const double* dr = m->rx_drdy + 2 * i;
            double inner = 0.0;
            int have = 0;
            if (dr[0] != 0) { inner = (dr[0] != 1) ? dr[0] * f : f; have = 1; }
            if (dr[1] != 0) {
                double v = (dr[1] == 1) ? r : dr[1] * r;
                inner = have ? inner - v : -v;
                have = 1;
            }
            if (fl & F_PMT_IN_JTEMP) inner = have ? inner + pres_mod_temp : +pres_mod_temp;
            if (fl & (F_PDEP | F_THD)) j_temp = -mw_avg * rho_inv * pres_mod[p] * (inner);
            else j_temp = -mw_avg * rho_inv * (inner);
        }
        if (fl & F_PMT_IN_JTEMP) {
            double e1 = arrhenius(m->arr_ratio + 4 * i, T, logT);
            if (fl & F_TROE) pres_mod_temp *= e1 * pow(Fcent, 1.0 / (1 + A * A / (B * B))) / (1.0 + Pr);
            else if (fl & F_SRI) {
                const double* s = m->sri_pm + 8 * i;
                double v = pow(s[0] * ((s[2] > 0) ? exp(-s[1] / T) : exp(s[1] / T))
                               + ((s[4] > 0) ? exp(-T / s[3]) : exp(T / s[3])), X);
                if (fl & F_SRI5) pres_mod_temp *= e1 * v * s[5] * pow(T, s[6]) / (1.0 + Pr);
                else pres_mod_temp *= e1 * v / (1.0 + Pr);
            } else pres_mod_temp *= e1 / (1.0 + Pr);
        }
        /* write_rates (cj:290-338) */
        if (fl & F_PLOG) kf = plog_kf(m, i, T, logT, pres, kf);
        else if (fl & F_CHEB) kf = cheb_kf(m, i, Tred, Pred);      /* Tred, Pred of the T part ({:.16e}) */
        else kf = arrhenius(m->arr_main + 4 * i, T, logT);
        if (rev) kr = kf / Kc;

        const int* rs = m->reac_sp + m->reac_off[i];
        const int* rn = m->reac_nu + m->reac_off[i];
        const int nre = m->reac_off[i + 1] - m->reac_off[i];
        const int* ps = m->prod_sp + m->prod_off[i];
        const int* pn = m->prod_nu + m->prod_off[i];
        const int npr = m->prod_off[i + 1] - m->prod_off[i];
        const int last = n - 1;
        int last_in_reac = 0, last_in_prod = 0;
        for (int a = 0; a < nre; ++a) if (rs[a] == last) last_in_reac = 1;
        for (int a = 0; a < npr; ++a) if (ps[a] == last) last_in_prod = 1;

/* __get_s_term (cj:410-448): k * nu * C_j^(nu-1) * prod(others) */
#define S_TERM(out, kk, sp_, nu_, cnt_, jsp)                                         \
        do {                                                                         \
            double v_ = (kk);                                                        \
            int a0_ = 0;                                                             \
            for (int a_ = 0; a_ < (cnt_); ++a_) if ((sp_)[a_] == (jsp)) a0_ = a_;      \
            if ((nu_)[a0_] != 1) v_ = v_ * (double)(nu_)[a0_];                        \
            for (int r_ = 0; r_ < (nu_)[a0_] - 1; ++r_) v_ = v_ * conc[(jsp)];        \
            for (int a_ = 0; a_ < (cnt_); ++a_) {                                     \
                if ((sp_)[a_] == (jsp)) continue;                                     \
                for (int r_ = 0; r_ < (nu_)[a_]; ++r_) v_ = v_ * conc[(sp_)[a_]];     \
            }                                                                        \
            (out) = v_;                                                              \
        } while (0)

        for (int j = 0; j < n - 1; ++j) {
            /* write_dr_dy_species (cj:375-489) */
            double d = j_temp * m->j_cj[j];
            int am = (fl & (F_PDEP | F_THD)) ? m->alpha_mode[(size_t)i * (n - 1) + j] : 0;
            double av = m->alpha_val[(size_t)i * (n - 1) + j];
            if (am == 1) d = d + pres_mod_temp;
            else if (am == 2) d = d - pres_mod_temp;
            else if (am == 3) d = d + av * pres_mod_temp;
            else if (am == 4) d = d - pres_mod_temp * av;

            int j_in_reac = 0, j_in_prod = 0;
            for (int a = 0; a < nre; ++a) if (rs[a] == j) j_in_reac = 1;
            for (int a = 0; a < npr; ++a) if (ps[a] == j) j_in_prod = 1;
            int jr = j_in_reac, jp = rev && j_in_prod;
            int lr = last_in_reac, lp = rev && last_in_prod;
            if (jr || jp || lr || lp) {
                double tf = 0, tr = 0, lf = 0, lrv = 0;
                if (jr) S_TERM(tf, kf, rs, rn, nre, j);
                if (jp) S_TERM(tr, kr, ps, pn, npr, j);
                if (lr) S_TERM(lf, kf, rs, rn, nre, last);
                if (lp) S_TERM(lrv, kr, ps, pn, npr, last);
                /* '(-kr * ...)' == -(kr * ...) exactly */
                double lastgrp = 0.0;
                if (lr && lp) lastgrp = (lf - lrv);
                else if (lr) lastgrp = (lf);
                else if (lp) lastgrp = (-lrv);
                double mwf = m->j_mwfrac[j];
                if (fl & (F_PDEP | F_THD)) {
                    double S;
                    if (jr && jp) S = tf - tr;
                    else if (jr) S = tf;
                    else if (jp) S = -tr;
                    else S = 0.0;
                    if (lr || lp) {
                        if (jr || jp) S = S - mwf * (lastgrp);
                        else S = -mwf * (lastgrp);
                    }
                    d = d + pres_mod[p] * (S);
                } else {
                    if (jr) d = d + tf;
                    if (jp) d = d - tr;
                    if (lr || lp) d = d - mwf * (lastgrp);
                }
            }

            for (int e = m->net_off[i]; e < m->net_off[i + 1]; ++e) {
                int k = m->net_sp[e], nu = m->net_nu[e];
                double mw_frac = (m->sp_mw[k] / m->sp_mw[j]) * (double)nu;
                double v;
                if (mw_frac == -1.0) v = -(d);
                else if (mw_frac != 1.0) v = mw_frac * (d);
                else v = (d);
                if (k + 1 < n) {
                    size_t lin = (size_t)k + 1 + (size_t)n * (j + 1);
                    if (touched[lin]) jac[lin] += v; else jac[lin] = v;
                    touched[lin] = 1;
                } else {
                    if (Jnpj_touched[j]) J_nplusjplus[j] += v; else J_nplusjplus[j] = v;
                    Jnpj_touched[j] = 1;
                }
            }
        }
#undef S_TERM
    }

    /* ---------------- energy equation (cj:2984-3254) */
    oracle_eval_h(m, T, h);
    oracle_eval_cp(m, T, cp);
    double cp_avg = 0.0;
    for (int k = 0; k < n - 1; ++k) {
        double v = (y[k + 1] * cp[k]);
        cp_avg = (k == 0) ? v : cp_avg + v;
    }
    cp_avg = (n > 1) ? cp_avg + (y_N * cp[n - 1]) : (y_N * cp[n - 1]);
    jac[0] = 0.0;
    touched[0] = 1;
    double working_temp = (1.0 / cp_avg);
    j_temp = 1.0 / (rho * cp_avg * cp_avg);
    const double mwN = m->sp_mw[n - 1];
    for (int k = 0; k < n; ++k) {
        for (int j = 0; j < n - 1; ++j) {
            size_t lin = (size_t)k + 1 + (size_t)n * (j + 1);
            double cst = (m->sp_mw[k] / m->sp_mw[j]) * (1. - m->sp_mw[j] / mwN);
            int jt = 0;
            if (k + 1 < n && touched[lin]) {
                jac[lin] += (spec_rates[k] * mw_avg * cst * rho_inv);
                jt = 1;
            } else if (k + 1 == n && Jnpj_touched[j]) {
                J_nplusjplus[j] += (spec_rates[k] * mw_avg * cst * rho_inv);
                jt = 1;
            }
            if (!jt && !m->sp_seen[k]) continue;
            size_t my = (size_t)n * (j + 1);
            double sp_part = (j_temp * (cp[j] - cp[n - 1]) * spec_rates[k] * m->sp_mw8[k]);
            double val;
            if (jt) val = h[k] * (working_temp * ((k + 1 < n) ? jac[lin] : J_nplusjplus[j]) - sp_part);
            else val = h[k] * (-sp_part);
            if (touched[my]) jac[my] -= val; else jac[my] = -(val);
            touched[my] = 1;
        }
    }

    /* write_dcp_dt (cj:1314-1395) */
    for (int b = 0; b < m->n_dcp; ++b) {
        int lo = (T <= m->dcp_tmid[b]);
        double s = 0.0;
        for (int e = m->dcp_off[b]; e < m->dcp_off[b + 1]; ++e) {
            int k = m->dcp_sp[e];
            const double* c = lo ? m->sp_dcp_lo + 4 * k : m->sp_dcp_hi + 4 * k;
            double yk = (k + 1 != n) ? y[k + 1] : y_N;
            double v = (yk * m->sp_ru_mw[k] * (c[0] + T * (c[1] + T * (c[2] + c[3] * T))));
            s = (e == m->dcp_off[b]) ? v : s + v;
        }
        if (b == 0) working_temp = s; else working_temp += s;
    }

    /* write_dt_completion (cj:1870-1905) */
    {
        double s = 0.0;
        for (int k = 0; k < n; ++k) {
            double v = spec_rates[k] * m->sp_mw8[k] * (-working_temp * h[k] / cp_avg + cp[k]);
            s = (k == 0) ? v : s + v;
            if (k + 1 < n) s = s + jac[k + 1] * h[k] * rho;
            else if (J_nplusone_touched) s = s + J_nplusone * h[k] * rho;
        }
        jac[0] = -(s) / (rho * cp_avg);
    }
    free(touched);
    free(w);
}

/* ---------------- batch drivers: row-major states y[n][NSP], pres[n] ---------------- */
void oracle_eval_jacob_batch(const OracleMech* m, int n, const double* pres, const double* y,
                             double* jac, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    size_t nn = (size_t)m->nsp * m->nsp;
    #pragma omp parallel for
    for (int s = 0; s < n; ++s) {
        double* out = jac + s * nn;
        memset(out, 0, sizeof(double) * nn);
        oracle_eval_jacob(m, 0.0, pres[s], y + (size_t)s * m->nsp, out);
    }
}

void oracle_dydt_batch(const OracleMech* m, int n, const double* pres, const double* y,
                       double* dy, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    #pragma omp parallel for
    for (int s = 0; s < n; ++s)
        oracle_dydt(m, 0.0, pres[s], y + (size_t)s * m->nsp, dy + (size_t)s * m->nsp);
}

void oracle_dydt_conv_batch(const OracleMech* m, int n, const double* rho, const double* y,
                            double* dy, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    #pragma omp parallel for
    for (int s = 0; s < n; ++s)
        oracle_dydt_conv(m, 0.0, rho[s], y + (size_t)s * m->nsp, dy + (size_t)s * m->nsp);
}

/* conc[n][NSP], fwd[n][NR], rev[n][NREV], pres_mod[n][NPD], spec_rates[n][NSP] */
void oracle_rates_batch(const OracleMech* m, int n, const double* pres, const double* y,
                        double* conc, double* fwd, double* rev, double* pres_mod,
                        double* spec_rates, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    #pragma omp parallel for
    for (int s = 0; s < n; ++s) {
        const double* ys = y + (size_t)s * m->nsp;
        double y_N, mw_avg, rho;
        double* C = conc + (size_t)s * m->nsp;
        double* f = fwd + (size_t)s * m->nr;
        double* r = rev + (size_t)s * m->nrev;
        double* pm = pres_mod + (size_t)s * m->npd;
        double* sr = spec_rates + (size_t)s * m->nsp;
        oracle_eval_conc(m, ys[0], pres[s], ys + 1, &y_N, &mw_avg, &rho, C);
        oracle_eval_rxn_rates(m, ys[0], pres[s], C, f, r);
        oracle_get_rxn_pres_mod(m, ys[0], pres[s], C, pm);
        oracle_eval_spec_rates(m, f, r, pm, sr, &sr[m->nsp - 1]);
    }
}
