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

// x = (I - gamma J)^-1 r, one warp per state.  Shared memory per warp: the matrix column-major with leading
// dimension ldm (odd), three vectors of NSP doubles (WA, then the pivot rows as ints; WB, then the pivot
// reciprocals; the solution).  A lane owns the rows lane, lane + 32, ... (R = ceil(nsp / 32) of them).
//
// Left-looking LU with implicit partial pivoting: column j is loaded into registers, updated with every earlier
// column c < j (one broadcast of its element in pivot row c by shuffle, one shared-memory load and one fma per
// owned row that was not a pivot row yet), then its pivot is chosen among the rows not used so far, the rows below
// are scaled and the column is stored once.  Compared with a right-looking update this reads each L element once
// per later column but stores each element once instead of once per pivot, and no rows are ever swapped (a row's
// pivot step is kept in a register).  The right-hand side goes through the same update as one more column;
// the back substitution walks the columns from the last with the solution component broadcast by shuffle.
// The arithmetic (order of the fmas of an element, pivot choice) is that of the textbook right-looking form.
// gamma: one value for all states (gamma_s == nullptr) or one per state.  info[s] = 0, or c + 1 when column c had
// no usable pivot (the state's x is then not written).
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
    double* wa = M + (size_t)ldm * nsp;
    double* wb = wa + nsp;
    double* xs = wb + nsp;
    const unsigned aM = (unsigned)__cvta_generic_to_shared(M);
    const unsigned aPV = (unsigned)__cvta_generic_to_shared(wa);      // pivot row of step c (int), reuses WA after the expansion
    const unsigned aIP = (unsigned)__cvta_generic_to_shared(wb);      // 1 / pivot of step c, reuses WB
    const unsigned aX = (unsigned)__cvta_generic_to_shared(xs);
    const unsigned cstep = (unsigned)ldm * 8u;                        // bytes between columns
    const unsigned arow = 8u * (unsigned)lane;                        // byte offset of the lane's first row in a column
    auto pick = [&](const double (&v)[R], int q) {                    // v[q], q warp-uniform
        double o = v[0];
#pragma unroll
        for (int t = 1; t < R; ++t) o = q == t ? v[t] : o;
        return o;
    };
    for (long long s = (long long)blockIdx.x * wpb + warp; s < n; s += (long long)gridDim.x * wpb) {
        const double g = gamma_s ? gamma_s[s] : gamma;
        // ---- expand -gamma J (+ I) into shared memory
        for (int k = lane; k < last; k += 32) {
            wa[k] = fac_at(f, nf, s, nsp + last + k);
            wb[k] = fac_at(f, nf, s, nsp + 2 * last + k);
            M[k + 1] = -g * fac_at(f, nf, s, nsp + k);                       // column 0: temperature column
        }
        for (int j = lane; j < nsp; j += 32) M[(size_t)j * ldm] = -g * fac_at(f, nf, s, j);   // row 0: energy-equation row
        __syncwarp();
        {
            double was[R], wbs[R];
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int k = lane + 32 * q;
                was[q] = k < last ? -g * wa[k] : 0.0;
                wbs[q] = k < last ? -g * wb[k] : 0.0;
            }
            unsigned ac = aM + cstep + arow + 8u;
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
        // ---- factorisation, column by column; ps[q]: the step at which the lane's q-th row became a pivot row
        // (nsp: not yet; -1: the row does not exist)
        int ps[R];
#pragma unroll
        for (int q = 0; q < R; ++q) ps[q] = lane + 32 * q < nsp ? nsp : -1;
        int bad = 0;
        // Panels of NB columns: the updates with the columns of earlier panels (phase 1) run on NB independent
        // shuffle / fma chains that share the loads of the L columns; inside a panel (phase 2) the multipliers are
        // still in registers.  Column index nsp is the right-hand side, taken through the same updates.
        constexpr int NB = R <= 2 ? 8 : 4;
        const int* pvr = reinterpret_cast<const int*>(wa);
        for (int j0 = 0; j0 <= nsp && !bad; j0 += NB) {
            double a[NB][R];
            int pr[NB];                                                      // pivot rows chosen inside this panel
#pragma unroll
            for (int t = 0; t < NB; ++t) {
                const int col = j0 + t;
                pr[t] = 0;
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    a[t][q] = 0.0;
                    if (ps[q] >= 0 && col < nsp) a[t][q] = lds64(aM + (unsigned)col * cstep + arow + 256u * q);
                    if (ps[q] >= 0 && col == nsp) a[t][q] = r[s * r_ss + (long long)(lane + 32 * q) * r_sv];
                }
            }
            // phase 1: plain loads (nothing is stored inside this loop, so the compiler may unroll it and hoist the
            // loads of the pivot rows and of the L columns ahead of the chains)
            const double* lcol = M + lane;
            const int cend = j0 < nsp ? j0 : nsp;
#pragma unroll 2
            for (int c = 0; c < cend; ++c, lcol += ldm) {
                const int pc = pvr[c];
                double u[NB];
#pragma unroll
                for (int t = 0; t < NB; ++t) u[t] = __shfl_sync(0xffffffffu, pick(a[t], pc >> 5), pc & 31);
#pragma unroll
                for (int q = 0; q < R; ++q)
                    if (ps[q] > c) {
                        const double l = lcol[32 * q];
#pragma unroll
                        for (int t = 0; t < NB; ++t) a[t][q] = fma(-l, u[t], a[t][q]);
                    }
            }
            // phase 2: column by column inside the panel
#pragma unroll
            for (int t = 0; t < NB; ++t) {
                const int col = j0 + t;
                if (col <= nsp && !bad) {                                    // (no break: the loop must unroll so that a[][] stays in registers)
#pragma unroll
                for (int t2 = 0; t2 < t; ++t2) {
                    const int c = j0 + t2;
                    const double u = __shfl_sync(0xffffffffu, pick(a[t], pr[t2] >> 5), pr[t2] & 31);
#pragma unroll
                    for (int q = 0; q < R; ++q)
                        if (ps[q] > c) a[t][q] = fma(-a[t2][q], u, a[t][q]);
                }
                if (col == nsp) {
                    // back substitution: y sits in a[t][] by row; component c of the solution is y[pivot row c] / pivot c.
                    // The last columns, their pivot rows and reciprocals may have been stored in this very panel.
                    __syncwarp();
                    for (int c = nsp - 1; c >= 0; --c) {
                        const int pc = pvr[c];
                        const double xc = __shfl_sync(0xffffffffu, pick(a[t], pc >> 5), pc & 31) * lds64(aIP + 8u * c);
                        if (lane == 0) sts64(aX + 8u * c, xc);
                        const unsigned ac = aM + (unsigned)c * cstep + arow;
#pragma unroll
                        for (int q = 0; q < R; ++q)
                            if (ps[q] >= 0 && ps[q] < c) a[t][q] = fma(-lds64(ac + 256u * q), xc, a[t][q]);     // U[ps][c] x_c
                    }
                } else {
                // pivot of this column among the rows not used so far
                double best = -1.0;
                int arg = 0x7fffffff;
#pragma unroll
                for (int q = 0; q < R; ++q)
                    if (ps[q] == nsp) {
                        const double v = fabs(a[t][q]);
                        if (v > best) { best = v; arg = lane + 32 * q; }
                    }
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
                    if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
                }
                if (!(best > 0.0) || !isfinite(best)) bad = col + 1;
                else {
                const double ip = 1.0 / __shfl_sync(0xffffffffu, pick(a[t], arg >> 5), arg & 31);
                pr[t] = arg;
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    if (lane + 32 * q == arg) ps[q] = col;
                    else if (ps[q] == nsp) a[t][q] *= ip;                     // multipliers of the rows still to be eliminated
                    if (ps[q] >= 0) sts64(aM + (unsigned)col * cstep + arow + 256u * q, a[t][q]);
                }
                if (lane == 0) {
                    asm volatile("st.shared.s32 [%0], %1;" ::"r"(aPV + 4u * col), "r"(arg) : "memory");
                    sts64(aIP + 8u * col, ip);
                }
                }
                }
                }
            }
            __syncwarp();
        }
        if (!bad) {
            __syncwarp();
            for (int j = lane; j < nsp; j += 32) x[s * x_ss + (long long)j * x_sv] = xs[j];
        }
        if (info && lane == 0) info[s] = bad;
        __syncwarp();
    }
}

}  // namespace pjc
