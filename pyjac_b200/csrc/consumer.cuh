// Consumers of the factored Jacobian: the record k_eval<.., M_FACT> writes is
//
//     fac[0 .. nsp)                 energy-equation row          J[0][j]
//     fac[nsp + k]                  temperature column           J[k+1][0]           k = 0 .. nsp-2
//     fac[nsp + (nsp-1) + k]        WA_k = W_k a_k
//     fac[nsp + 2 (nsp-1) + k]      WB_k = W_k b_k
//     fac[nsp + 3 (nsp-1) + s]      sparse block S_s             (row_s, col_s) of a pattern fixed per mechanism
//
//     J[k+1][j] = ca_j WA_k + cb_j WB_k + sum_{s: (row_s, col_s) = (k+1, j)} S_s          j = 1 .. nsp-1
//
// with ca_j = 1 / W_{j-1}, cb_j = 1 / W_N (p5_colfac).  The dense NSP x NSP Jacobian the reference
// writes (create_jacobian.py:3301-3404 gestures at a sparse form with `sparse_multiplier`, broken at :3322)
// is this record expanded; nothing below ever stores it to HBM.
//
//   k_jvp     out = J v per state, one thread per state, straight from the record
//   k_newton  x = (I - gamma J)^-1 r per state: the matrix an implicit integrator factorises
//             (docs/faqs.rst:113-117) is expanded into shared memory by one warp per state, LU with
//             partial pivoting, forward / back substitution
#pragma once
#include "common.cuh"

namespace pjc {

enum : int { NEWTON_MAX_R = 6 };      // rows per lane k_newton is instantiated for: NSP <= 192 (shared memory ends it near 165)

struct Fac {
    int nsp, nnz;
    const double* fac;       // the records
    int sf;                  // 1: state-fastest fac[e * ld + s], 0: one record per state fac[s * nf + e]
    long long ld;
    const double2* colfac;   // (ca_j, cb_j), j = 0 .. nsp-1 (entry 0 unused)
    // sparse block by row (CSR over Jacobian rows 1 .. nsp-1): ptr[nsp], then per entry the slot and the column
    const int* ptr;
    const int* slot;
    const int* col;
    // and in record order (column-major): Jacobian row, column
    const int* rows;
    const int* cols;
};

__device__ __forceinline__ double fac_at(const Fac& f, long long nf, long long s, int e)
{
    return f.sf ? f.fac[(long long)e * f.ld + s] : f.fac[s * nf + e];
}

// out = J v.  v and out are addressed as base[s * ss + i * sv] (i = 0 .. nsp-1).
__global__ void __launch_bounds__(128) k_jvp(Fac f, int n, const double* __restrict__ v, long long v_ss, long long v_sv,
                                             double* __restrict__ out, long long o_ss, long long o_sv)
{
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int nsp = f.nsp, last = nsp - 1;
    const long long nf = (long long)nsp + 3 * last + f.nnz;
    const double* vs = v + s * v_ss;
    double* os = out + s * o_ss;
    const double v0 = vs[0];
    double e = fac_at(f, nf, s, 0) * v0, sa = 0.0, sb = 0.0;
    for (int j = 1; j < nsp; ++j) {
        const double vj = vs[(long long)j * v_sv];
        const double2 c = __ldg(f.colfac + j);
        e = fma(fac_at(f, nf, s, j), vj, e);
        sa = fma(c.x, vj, sa);
        sb = fma(c.y, vj, sb);
    }
    os[0] = e;
    for (int k = 0; k < last; ++k) {
        double a = fac_at(f, nf, s, nsp + k) * v0;
        a = fma(fac_at(f, nf, s, nsp + last + k), sa, a);
        a = fma(fac_at(f, nf, s, nsp + 2 * last + k), sb, a);
        const int p0 = __ldg(f.ptr + k), p1 = __ldg(f.ptr + k + 1);
        for (int p = p0; p < p1; ++p)
            a = fma(fac_at(f, nf, s, __ldg(f.slot + p)), vs[(long long)__ldg(f.col + p) * v_sv], a);
        os[(long long)(k + 1) * o_sv] = a;
    }
}

__device__ __forceinline__ double lds64(unsigned a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts64(unsigned a, double v)
{
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}

// x = (I - gamma J)^-1 r, one warp per state.  Shared memory per warp: the matrix column-major with
// leading dimension ldm (odd), the right-hand side, WA / WB of the state.  A lane owns the rows
// lane, lane + 32, ... (R = ceil(nsp / 32) of them): their multipliers stay in registers through the
// trailing update, which walks the columns with one broadcast load of the pivot row's element and one
// load / fma / store per owned row.  gamma: one value for all states (gamma_s == nullptr) or one per
// state.  info[s] = 0, or c + 1 when column c had no usable pivot (the state's x is then not written).
template <int R>
__global__ void __launch_bounds__(256) k_newton(Fac f, int n, double gamma, const double* __restrict__ gamma_s,
                                                const double* __restrict__ r, long long r_ss, long long r_sv,
                                                double* __restrict__ x, long long x_ss, long long x_sv,
                                                int* __restrict__ info, int ldm)
{
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int nsp = f.nsp, last = nsp - 1;
    const long long nf = (long long)nsp + 3 * last + f.nnz;
    double* M = sm + (size_t)warp * ((size_t)ldm * nsp + 3 * (size_t)nsp);
    double* b = M + (size_t)ldm * nsp;
    double* wa = b + nsp;
    double* wb = wa + nsp;
    const unsigned aM = (unsigned)__cvta_generic_to_shared(M), aB = (unsigned)__cvta_generic_to_shared(b);
    const unsigned cstep = (unsigned)ldm * 8u;                  // bytes between columns
    for (long long s = (long long)blockIdx.x * wpb + warp; s < n; s += (long long)gridDim.x * wpb) {
        const double g = gamma_s ? gamma_s[s] : gamma;
        // ---- expand -gamma J (+ I) into shared memory
        for (int k = lane; k < last; k += 32) {
            wa[k] = fac_at(f, nf, s, nsp + last + k);
            wb[k] = fac_at(f, nf, s, nsp + 2 * last + k);
            M[k + 1] = -g * fac_at(f, nf, s, nsp + k);                       // column 0: temperature column
        }
        for (int j = lane; j < nsp; j += 32) {
            M[(size_t)j * ldm] = -g * fac_at(f, nf, s, j);                   // row 0: energy-equation row
            b[j] = r[s * r_ss + (long long)j * r_sv];
        }
        __syncwarp();
        {
            double was[R], wbs[R];
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int k = lane + 32 * q;
                was[q] = k < last ? -g * wa[k] : 0.0;
                wbs[q] = k < last ? -g * wb[k] : 0.0;
            }
            unsigned ac = aM + cstep + 8u * (unsigned)(lane + 1);
            for (int j = 1; j < nsp; ++j, ac += cstep) {
                const double2 c = __ldg(f.colfac + j);
#pragma unroll
                for (int q = 0; q < R; ++q)
                    if (lane + 32 * q < last) sts64(ac + 256u * q, fma(c.x, was[q], c.y * wbs[q]));
            }
        }
        __syncwarp();
        for (int p = lane; p < f.nnz; p += 32) {
            // an element has one slot: no two lanes touch the same address
            double* q = M + (size_t)__ldg(f.cols + p) * ldm + __ldg(f.rows + p);
            *q = fma(-g, fac_at(f, nf, s, nsp + 3 * last + p), *q);
        }
        __syncwarp();
        for (int j = lane; j < nsp; j += 32) M[(size_t)j * ldm + j] += 1.0;
        __syncwarp();
        // ---- LU with partial pivoting (right-looking), the right-hand side carried along
        int bad = 0;
        for (int c = 0; c < nsp; ++c) {
            const unsigned acol = aM + (unsigned)c * cstep;
            double best = -1.0;
            int arg = c;
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int i = lane + 32 * q;
                if (i >= c && i < nsp) {
                    const double a = fabs(lds64(acol + 8u * i));
                    if (a > best) { best = a; arg = i; }
                }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
                if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
            }
            if (!(best > 0.0) || !isfinite(best)) { bad = c + 1; break; }
            if (arg != c) {
                for (int j = lane; j < nsp; j += 32) {
                    const unsigned q = aM + (unsigned)j * cstep;
                    const double t0 = lds64(q + 8u * c), t1 = lds64(q + 8u * arg);
                    sts64(q + 8u * c, t1); sts64(q + 8u * arg, t0);
                }
                if (lane == 0) { const double t0 = lds64(aB + 8u * c), t1 = lds64(aB + 8u * arg); sts64(aB + 8u * c, t1); sts64(aB + 8u * arg, t0); }
            }
            __syncwarp();
            const double ip = 1.0 / lds64(acol + 8u * c);
            const double bc = lds64(aB + 8u * c);
            __syncwarp();
            // multipliers of the owned rows below the pivot (0 for the others: their updates are no-ops
            // that are skipped by predicate), forward substitution on the fly
            double l[R];
            bool on[R];
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int i = lane + 32 * q;
                on[q] = i > c && i < nsp;
                l[q] = 0.0;
                if (on[q]) {
                    l[q] = lds64(acol + 8u * i) * ip;
                    sts64(acol + 8u * i, l[q]);
                    sts64(aB + 8u * i, fma(-l[q], bc, lds64(aB + 8u * i)));
                }
            }
            // trailing update: column j -= l * M[c][j]
            unsigned aj = acol + cstep;
            int j = c + 1;
            for (; j + 4 <= nsp; j += 4, aj += 4 * cstep) {
                const double u0 = lds64(aj + 8u * c), u1 = lds64(aj + cstep + 8u * c);
                const double u2 = lds64(aj + 2 * cstep + 8u * c), u3 = lds64(aj + 3 * cstep + 8u * c);
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    if (!on[q]) continue;
                    const unsigned ai = aj + 8u * (unsigned)(lane + 32 * q);
                    const double a0 = lds64(ai), a1 = lds64(ai + cstep), a2 = lds64(ai + 2 * cstep), a3 = lds64(ai + 3 * cstep);
                    sts64(ai, fma(-l[q], u0, a0)); sts64(ai + cstep, fma(-l[q], u1, a1));
                    sts64(ai + 2 * cstep, fma(-l[q], u2, a2)); sts64(ai + 3 * cstep, fma(-l[q], u3, a3));
                }
            }
            for (; j < nsp; ++j, aj += cstep) {
                const double u = lds64(aj + 8u * c);
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    if (!on[q]) continue;
                    const unsigned ai = aj + 8u * (unsigned)(lane + 32 * q);
                    sts64(ai, fma(-l[q], u, lds64(ai)));
                }
            }
            __syncwarp();
        }
        if (!bad) {
            // ---- back substitution, column oriented
            for (int c = nsp - 1; c >= 0; --c) {
                const unsigned acol = aM + (unsigned)c * cstep;
                const double xc = lds64(aB + 8u * c) / lds64(acol + 8u * c);
                __syncwarp();
                if (lane == 0) sts64(aB + 8u * c, xc);
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    const int i = lane + 32 * q;
                    if (i < c) sts64(aB + 8u * i, fma(-lds64(acol + 8u * i), xc, lds64(aB + 8u * i)));
                }
                __syncwarp();
            }
            for (int j = lane; j < nsp; j += 32) x[s * x_ss + (long long)j * x_sv] = b[j];
        }
        if (info && lane == 0) info[s] = bad;
        __syncwarp();
    }
}

}  // namespace pjc
