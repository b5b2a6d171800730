// psi_aaa.cuh -- warp-cooperative AAA rational approximation on the device (kernel K4's building
// blocks): Loewner-matrix fit by a one-sided complex Jacobi SVD, greedy support-point selection,
// Froissart-doublet cleanup, zeros and poles of the barycentric form by Ehrlich-Aberth iteration.
//
// Replaces RationalApproximation (complex_roots.py:38-368), which leans on LAPACK (numpy.linalg.svd,
// scipy.linalg.eig).  The mathematics is the same -- weights = right singular vector of the smallest
// singular value of the Loewner matrix; poles/zeros = roots of sum_j w_j/(z-z_j) and
// sum_j w_j f_j/(z-z_j) -- the numerical kernels are GPU-friendly ones:
//   * one-sided Jacobi keeps high relative accuracy for the tiny singular values AAA lives on and
//     needs only column operations (lanes split the rows);
//   * Aberth's simultaneous iteration works directly on the barycentric sums (O(m) per root and
//     step, one root per lane) instead of a generalised eigenvalue problem.
// One warp owns one approximation; all arrays live in a per-warp scratch area in global memory (L2).
// Written against the warp policy W like psi_quad.cuh, so the CPU harness runs the same code.
#pragma once
#include "psi_batch.cuh"

namespace psi {

#if defined(PSI_AAA_STATS)
// test-harness counters: fits, Jacobi sweeps, rotations, Aberth calls, Aberth iterations
static long long g_aaa_stats[8];
#define PSI_STAT(i, n) (g_aaa_stats[i] += (n))
#else
#define PSI_STAT(i, n)
#endif

constexpr int kMemoSlots = 12;
constexpr int kMaxSamples = 512;     // largest sample count (cap) the device AAA accepts

// scratch layout of one warp (all sizes in elements, cap = maximum number of samples)
struct AAAScratch {
    cx* Z;        // [cap]   sample points
    cx* F;        // [cap]   function samples
    int* mask;    // [cap]   1 = support point
    cx* A;        // [cap * ncol]  Loewner matrix, column-major (destroyed by the SVD)
    cx* A2;       // [cap * ncol]  Loewner matrix pre-rotated by the previous fit's V (warm start)
    cx* V;        // [ncol * ncol] right singular vectors
    cx* V2;       // [ncol * ncol] copy of the previous V while the warm start is assembled
    cx* w;        // [ncol]  weights
    double* R;    // [cap]   residuals at the non-support points (in sample order of those points)
    int* on;      // [ncol]  sample index of the k-th support point
    int* off;     // [cap]   sample index of the k-th non-support point
    cx* roots;    // [ncol]  Aberth iterates (zeros or poles)
    cx* tmp;      // [ncol]  second root buffer
    // greedy-phase result, replayed by PSIMode's clean_tol-halving loop
    int* g_mask;  // [cap]
    cx* g_w;      // [ncol]
    double* g_R;  // [cap]
    // work lists of the mode machine (psi_mode.cuh)
    cx* w0;       // [ncol]  start values of the polish
    cx* pol;      // [ncol]  polished values
    // memo of Loewner fits keyed by the support mask (kMemoSlots entries, round robin): PSIMode's
    // clean_tol-halving loop revisits the same masks again and again
    cx* memo_w;        // [kMemoSlots * ncol]
    double* memo_R;    // [kMemoSlots * cap]
    double* memo_res;  // [kMemoSlots]
    int* memo_mask;    // [kMemoSlots * cap]
    int* memo_m;       // [kMemoSlots]   number of support points of the entry, -1 = empty
    cx* memo_pole;     // [kMemoSlots * ncol]  per pole inside the domain: (residue ratio, sample to drop)
    cx* memo_zero;     // [kMemoSlots * ncol]  zeros of the approximant
    int* memo_np;      // [kMemoSlots]   number of stored pole decisions, -1 = not computed yet
    int* memo_nz;      // [kMemoSlots]   number of stored zeros, -1 = not computed yet
    double* nrm;       // [ncol]  squared column norms (Jacobi)
    int cap, ncol;
};

PSI_HD size_t aaa_scratch_bytes(int cap) {
    const size_t ncol = (size_t)cap / 2 + 2;
    return sizeof(cx) * (2 * (size_t)cap + 2 * (size_t)cap * ncol + 2 * ncol * ncol + (6 + 3 * kMemoSlots) * ncol) +
           sizeof(double) * ((2 + kMemoSlots) * (size_t)cap + kMemoSlots + ncol) +
           sizeof(int) * ((3 + kMemoSlots) * (size_t)cap + ncol + 3 * kMemoSlots) + 256;
}

PSI_HD AAAScratch aaa_carve(void* base, int cap) {
    AAAScratch s;
    const size_t ncol = (size_t)cap / 2 + 2;
    s.cap = cap; s.ncol = (int)ncol;
    cx* c = reinterpret_cast<cx*>(base);
    s.Z = c; c += cap;
    s.F = c; c += cap;
    s.A = c; c += (size_t)cap * ncol;
    s.A2 = c; c += (size_t)cap * ncol;
    s.V = c; c += ncol * ncol;
    s.V2 = c; c += ncol * ncol;
    s.w = c; c += ncol;
    s.roots = c; c += ncol;
    s.tmp = c; c += ncol;
    s.g_w = c; c += ncol;
    s.w0 = c; c += ncol;
    s.pol = c; c += ncol;
    s.memo_w = c; c += (size_t)kMemoSlots * ncol;
    s.memo_pole = c; c += (size_t)kMemoSlots * ncol;
    s.memo_zero = c; c += (size_t)kMemoSlots * ncol;
    double* d = reinterpret_cast<double*>(c);
    s.R = d; d += cap;
    s.g_R = d; d += cap;
    s.memo_R = d; d += (size_t)kMemoSlots * cap;
    s.memo_res = d; d += kMemoSlots;
    s.nrm = d; d += ncol;
    int* i = reinterpret_cast<int*>(d);
    s.mask = i; i += cap;
    s.off = i; i += cap;
    s.g_mask = i; i += cap;
    s.memo_mask = i; i += (size_t)kMemoSlots * cap;
    s.memo_m = i; i += kMemoSlots;
    s.memo_np = i; i += kMemoSlots;
    s.memo_nz = i; i += kMemoSlots;
    s.on = i;
    return s;
}

struct AAAState {
    int M;          // number of samples
    int m;          // number of support points
    int r;          // number of non-support points
    double fmax;    // max |F|
    double tol, clean_tol;
    RectDom dom;
    int memo_next;  // next memo slot to overwrite
    int slot;       // memo slot that holds the current mask (set by aaa_fit)
    int v_m;        // s.V holds the right singular vectors of the last computed fit with v_m support
                    // points (0 = nothing usable); lets the next greedy step start its SVD warm
};

template <class W>
PSI_HD void aaa_memo_clear(const AAAScratch& s, AAAState& st) {
    for (int k = W::lane(); k < kMemoSlots; k += W::width) s.memo_m[k] = -1;
    st.memo_next = 0;
    W::sync();
}

// rebuild the index lists from the mask (every lane computes the same lists)
template <class W>
PSI_HD void aaa_index(const AAAScratch& s, AAAState& st) {
    W::sync();
    if (W::lane() == 0) {
        int m = 0, r = 0;
        for (int i = 0; i < st.M; ++i) {
            if (s.mask[i]) s.on[m++] = i; else s.off[r++] = i;
        }
    }
    W::sync();
    int m = 0;
    for (int i = 0; i < st.M; ++i) m += s.mask[i] ? 1 : 0;
    st.m = m; st.r = st.M - m;
}

// One-sided Jacobi SVD of the r x m matrix A (column-major, leading dimension r): on return the
// columns of A are orthogonal, V holds the accumulated rotations; returns the index of the column
// with the smallest norm (the right singular vector of the smallest singular value is V[:, j]).
// CH = row chunks of one warp width per lane: the per-rotation work is straight-line predicated code
// (no inner loops: the rows of a column pair live in registers between the dot product and the
// rotation), which matters because this routine is instruction-fetch bound on the GPU.
template <class W, int CH>
PSI_NOINLINE int jacobi_smallest_t(cx* A, cx* V, double* nrm, int r, int m, bool warm) {
    const int lane = W::lane();
    if (!warm)
        for (int j = 0; j < m; ++j)
            for (int i = lane; i < m; i += W::width) V[i + (size_t)j * m] = cx(i == j ? 1.0 : 0.0, 0.0);
    W::sync();
    PSI_STAT(0, 1);
    double fro2 = 0.0;
    for (int j = 0; j < m; ++j) {
        const cx* aj = A + (size_t)j * r;
        for (int i = lane; i < r; i += W::width) fro2 += aj[i].re * aj[i].re + aj[i].im * aj[i].im;
    }
    fro2 = W::sum(fro2);
    const double negligible = 1.0e-30 * fro2;      // squared norm of a column that is pure rounding noise
    for (int sweep = 0; sweep < 40; ++sweep) {
        bool rotated = false;
        PSI_STAT(1, 1);
        // squared column norms, refreshed every sweep and updated analytically after each rotation
        for (int j = 0; j < m; ++j) {
            double n2 = 0.0;
            const cx* aj = A + (size_t)j * r;
            for (int i = lane; i < r; i += W::width) n2 += aj[i].re * aj[i].re + aj[i].im * aj[i].im;
            n2 = W::sum(n2);
            if (lane == 0) nrm[j] = n2;
        }
        W::sync();
        for (int p = 0; p < m - 1; ++p) {
            for (int q = p + 1; q < m; ++q) {
                cx* ap = A + (size_t)p * r;
                cx* aq = A + (size_t)q * r;
                const double al = nrm[p], be = nrm[q];
                if (al <= negligible && be <= negligible) continue;
                constexpr int NX = CH > 0 ? CH : 1;
                cx x[NX], y[NX];
                double gr = 0.0, gi = 0.0;
                if (CH > 0) {
#pragma unroll
                    for (int c = 0; c < NX; ++c) {
                        const int i = lane + c * W::width;
                        if (i < r) { x[c] = ap[i]; y[c] = aq[i]; } else { x[c] = cx(0.0, 0.0); y[c] = cx(0.0, 0.0); }
                        gr += x[c].re * y[c].re + x[c].im * y[c].im;          // conj(x)*y
                        gi += x[c].re * y[c].im - x[c].im * y[c].re;
                    }
                } else {                                  // CH == 0: any number of rows, looped
                    for (int i = lane; i < r; i += W::width) {
                        const cx xx = ap[i], yy = aq[i];
                        gr += xx.re * yy.re + xx.im * yy.im;
                        gi += xx.re * yy.im - xx.im * yy.re;
                    }
                }
                gr = W::sum(gr); gi = W::sum(gi);
                const double g = sqrt(gr * gr + gi * gi);
                if (g <= 1.0e-14 * sqrt(al * be) || g == 0.0) continue;
                rotated = true;
                PSI_STAT(2, 1);
                // Hermitian 2x2 [[al, gamma],[conj(gamma), be]]: rotation that annihilates gamma
                const double zeta = (be - al) / (2.0 * g);
                const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
                const double ig = 1.0 / g;
                const cx ph(gr * ig, gi * ig);                 // gamma/|gamma|
                // a_p' = c*a_p - sn*conj(ph)*a_q ;  a_q' = sn*ph*a_p + c*a_q
                const cx sp(sn * ph.re, -sn * ph.im), sq(sn * ph.re, sn * ph.im);
                cx* vp = V + (size_t)p * m;
                cx* vq = V + (size_t)q * m;
                if (CH > 0) {
#pragma unroll
                    for (int k = 0; k < NX; ++k) {
                        const int i = lane + k * W::width;
                        if (i < r) { ap[i] = c * x[k] - sp * y[k]; aq[i] = sq * x[k] + c * y[k]; }
                    }
#pragma unroll
                    for (int k = 0; k < NX; ++k) {
                        const int i = lane + k * W::width;
                        if (i < m) {
                            const cx vx = vp[i], vy = vq[i];
                            vp[i] = c * vx - sp * vy;
                            vq[i] = sq * vx + c * vy;
                        }
                    }
                } else {
                    for (int i = lane; i < r; i += W::width) {
                        const cx xx = ap[i], yy = aq[i];
                        ap[i] = c * xx - sp * yy; aq[i] = sq * xx + c * yy;
                    }
                    for (int i = lane; i < m; i += W::width) {
                        const cx vx = vp[i], vy = vq[i];
                        vp[i] = c * vx - sp * vy; vq[i] = sq * vx + c * vy;
                    }
                }
                if (lane == 0) { nrm[p] = fmax(al - t * g, 0.0); nrm[q] = be + t * g; }
                W::sync();
            }
        }
        if (!rotated) break;
    }
    int best = 0;
    double bn = -1.0;
    for (int j = 0; j < m; ++j) {
        double n2 = 0.0;
        const cx* aj = A + (size_t)j * r;
        for (int i = lane; i < r; i += W::width) n2 += aj[i].re * aj[i].re + aj[i].im * aj[i].im;
        n2 = W::sum(n2);
        if (bn < 0.0 || n2 < bn) { bn = n2; best = j; }
    }
    return best;
}

template <class W>
PSI_HD int jacobi_smallest(cx* A, cx* V, double* nrm, int r, int m, bool warm) {
    const int rows = r > m ? r : m;
    if constexpr (W::width == 1) {
        return jacobi_smallest_t<W, 0>(A, V, nrm, r, m, warm);                     // CPU harness: one "lane"
    } else {
        // two specialisations only: the routine is instruction-fetch bound, code size matters
        if (rows <= 2 * W::width) return jacobi_smallest_t<W, 2>(A, V, nrm, r, m, warm);
        return jacobi_smallest_t<W, 0>(A, V, nrm, r, m, warm);
    }
}

// Loewner fit for the current mask: weights and residuals (complex_roots.py:56-90).  Returns the
// largest residual divided by max|F|.
template <class W>
PSI_NOINLINE double aaa_fit(const AAAScratch& s, AAAState& st, int added_sample = -1) {
    aaa_index<W>(s, st);
    const int lane = W::lane(), r = st.r, m = st.m;
    // memo lookup
    for (int k = 0; k < kMemoSlots; ++k) {
        if (s.memo_m[k] != m) continue;
        bool same = true;
        const int* mk = s.memo_mask + (size_t)k * s.cap;
        for (int i = lane; i < st.M; i += W::width) if (mk[i] != s.mask[i]) same = false;
        if (W::any(!same)) continue;
        for (int j = lane; j < m; j += W::width) s.w[j] = s.memo_w[(size_t)k * s.ncol + j];
        for (int i = lane; i < r; i += W::width) s.R[i] = s.memo_R[(size_t)k * s.cap + i];
        W::sync();
        st.slot = k;
        st.v_m = 0;                                  // V belongs to some other mask now
        return s.memo_res[k];
    }
    for (int j = 0; j < m; ++j) {
        const cx zj = s.Z[s.on[j]], fj = s.F[s.on[j]];
        for (int i = lane; i < r; i += W::width) {
            const cx c = recip(s.Z[s.off[i]] - zj);
            s.A[i + (size_t)j * r] = (s.F[s.off[i]] - fj) * c;        // F_i C_ij - C_ij f_j
        }
    }
    W::sync();
    if (r == 0) {
        for (int j = lane; j < m; j += W::width) s.w[j] = cx(j == m - 1 ? 1.0 : 0.0, 0.0);
    } else {
        // Warm start of the greedy chain: this mask = previous mask + one support point.  The Loewner
        // entries depend on (sample, support point) only, so the old columns are the previous matrix
        // minus one row; rotated by the previous V they are orthogonal up to a rank-one change, and
        // Jacobi needs ~2-3 sweeps instead of ~9.
        cx* mat = s.A;
        bool warm = false;
        if (added_sample >= 0 && st.v_m == m - 1 && m >= 3) {
            int pnew = 0;
            while (pnew < m && s.on[pnew] != added_sample) ++pnew;
            if (pnew < m) {
                warm = true;
                const int mo = m - 1;
                for (int e = lane; e < mo * mo; e += W::width) s.V2[e] = s.V[e];
                W::sync();
                for (int i = lane; i < r; i += W::width) {
                    for (int j = 0; j < mo; ++j) {
                        cx acc(0.0, 0.0);
                        for (int k = 0; k < mo; ++k)
                            acc = acc + s.A[i + (size_t)(k + (k >= pnew)) * r] * s.V2[k + (size_t)j * mo];
                        s.A2[i + (size_t)(j + (j >= pnew)) * r] = acc;
                    }
                    s.A2[i + (size_t)pnew * r] = s.A[i + (size_t)pnew * r];
                }
                for (int e = lane; e < m * m; e += W::width) {
                    const int i = e % m, j = e / m;
                    cx v(0.0, 0.0);
                    if (i == pnew || j == pnew) v = cx(i == j ? 1.0 : 0.0, 0.0);
                    else v = s.V2[(i - (i > pnew)) + (size_t)(j - (j > pnew)) * mo];
                    s.V[e] = v;
                }
                W::sync();
                mat = s.A2;
            }
        }
        const int jmin = jacobi_smallest<W>(mat, s.V, s.nrm, r, m, warm);
        st.v_m = m;
        for (int j = lane; j < m; j += W::width) s.w[j] = s.V[j + (size_t)jmin * m];
    }
    W::sync();
    double rmax = 0.0;
    for (int i = lane; i < r; i += W::width) {
        const cx Zi = s.Z[s.off[i]];
        cx num(0.0, 0.0), den(0.0, 0.0);
        for (int j = 0; j < m; ++j) {
            const cx t = s.w[j] * recip(Zi - s.Z[s.on[j]]);
            num = num + t * s.F[s.on[j]];
            den = den + t;
        }
        const double res = abs_c(s.F[s.off[i]] - num * recip(den));
        s.R[i] = res;
        rmax = fmax(rmax, res);
    }
    W::sync();
    const double resid = W::max(rmax) / st.fmax;
    {   // memo insert
        const int k = st.memo_next;
        st.memo_next = (k + 1) % kMemoSlots;
        for (int i = lane; i < st.M; i += W::width) s.memo_mask[(size_t)k * s.cap + i] = s.mask[i];
        for (int j = lane; j < m; j += W::width) s.memo_w[(size_t)k * s.ncol + j] = s.w[j];
        for (int i = lane; i < r; i += W::width) s.memo_R[(size_t)k * s.cap + i] = s.R[i];
        if (lane == 0) { s.memo_m[k] = m; s.memo_res[k] = resid; s.memo_np[k] = -1; s.memo_nz[k] = -1; }
        W::sync();
        st.slot = k;
    }
    return resid;
}

// barycentric evaluation r(z) = sum w_j f_j/(z-z_j) / sum w_j/(z-z_j)   (complex_roots.py:331-368)
PSI_HD cx aaa_eval(const AAAScratch& s, const AAAState& st, cx z) {
    cx num(0.0, 0.0), den(0.0, 0.0);
    for (int j = 0; j < st.m; ++j) {
        const cx t = s.w[j] * recip(z - s.Z[s.on[j]]);
        num = num + t * s.F[s.on[j]];
        den = den + t;
    }
    return num * recip(den);
}

// Roots of S(z) = sum_j a_j/(z - z_j) (a_j = w_j for the poles, w_j f_j for the zeros) by the
// Ehrlich-Aberth iteration on p(z) = S(z) prod(z - z_j): p/p' = 1/(S'/S + sum 1/(z - z_j)).
// m - 1 roots are written to out[]; roots that run away (|z| > 1e12 * scale: the leading coefficient
// sum a_j is numerically zero) are returned as +inf.  Returns the number of iterations used.
template <class W>
PSI_NOINLINE int aberth_roots(const AAAScratch& s, const AAAState& st, bool with_f, cx* out) {
    const int lane = W::lane(), m = st.m, nr = m - 1;
    if (nr <= 0) return 0;
    // scale and centre of the support points
    double cre = 0.0, cim = 0.0;
    for (int j = 0; j < m; ++j) { cre += s.Z[s.on[j]].re; cim += s.Z[s.on[j]].im; }
    cre /= m; cim /= m;
    double scale = 0.0;
    for (int j = 0; j < m; ++j) scale = fmax(scale, abs_c(s.Z[s.on[j]] - cx(cre, cim)));
    if (!(scale > 0.0)) scale = 1.0;
    // start values: between consecutive support points, nudged off the connecting line
    for (int k = lane; k < nr; k += W::width) {
        const cx a = s.Z[s.on[k]], b = s.Z[s.on[k + 1]];
        const cx d = b - a;
        out[k] = cx(0.5 * (a.re + b.re) - 0.31 * d.im + 0.013 * scale, 0.5 * (a.im + b.im) + 0.31 * d.re + 0.017 * scale);
    }
    W::sync();
    // Convergence per root: Aberth contracts cubically until the step reaches the rounding noise of
    // the barycentric sums (1e-16..1e-13 relative, depending on how the terms cancel) and then just
    // jitters; a root is frozen once its step is tiny and no longer shrinking fast.  prev[] = previous
    // step length (stored in tmp[k].im after the swap; roots themselves are double-buffered).
    int it = 0;
    PSI_STAT(3, 1);
    double prev_step[8];
    bool frozen[8];
    for (int q = 0; q < 8; ++q) { prev_step[q] = INFINITY; frozen[q] = false; }
    for (; it < 60; ++it) {
        PSI_STAT(4, 1);
        bool moving = false;
        int slot = 0;
        for (int k = lane; k < nr; k += W::width, ++slot) {
            const cx z = out[k];
            const int q = slot < 8 ? slot : 7;
            if (!finite_c(z) || frozen[q]) { s.tmp[k] = z; continue; }
            cx S(0.0, 0.0), S1(0.0, 0.0), T(0.0, 0.0);
            bool hit = false;
            for (int j = 0; j < m; ++j) {
                const cx dz = z - s.Z[s.on[j]];
                if (dz.re == 0.0 && dz.im == 0.0) { hit = true; break; }
                const cx inv = recip(dz);
                cx a = s.w[j];
                if (with_f) a = a * s.F[s.on[j]];
                const cx t = a * inv;
                S = S + t;
                S1 = S1 - t * inv;
                T = T + inv;
            }
            if (hit) {                                   // sitting on a support point: step aside
                s.tmp[k] = cx(z.re + 1.0e-3 * scale, z.im + 1.3e-3 * scale);
                moving = true;
                continue;
            }
            if (S.re == 0.0 && S.im == 0.0) { s.tmp[k] = z; frozen[q] = true; continue; }     // exact root
            const cx newton = recip(S1 * recip(S) + T);      // p/p'
            cx rep(0.0, 0.0);
            for (int l = 0; l < nr; ++l) {
                if (l == k) continue;
                const cx zl = out[l];
                if (!finite_c(zl)) continue;
                const cx dz = z - zl;
                if (dz.re == 0.0 && dz.im == 0.0) continue;
                rep = rep + recip(dz);
            }
            const cx step = np_div(newton, cx(1.0, 0.0) - newton * rep);
            cx zn = z - step;
            const double sl = abs_c(step);
            if (!finite_c(zn) || abs_c(zn - cx(cre, cim)) > 1.0e12 * scale) {
                zn = cx(INFINITY, 0.0);
            } else if (sl <= 1.0e-15 * abs_c(zn) ||
                       (sl <= 1.0e-9 * (abs_c(zn) + scale) && sl > 0.1 * prev_step[q])) {
                if (slot < 8) frozen[q] = true;
                if (sl > prev_step[q]) zn = z;              // keep the better iterate
            } else {
                moving = true;
            }
            prev_step[q] = sl;
            s.tmp[k] = zn;
        }
        W::sync();
        for (int k = lane; k < nr; k += W::width) out[k] = s.tmp[k];
        W::sync();
        if (!W::any(moving)) break;
    }
    return it;
}

// lexicographic order of numpy's sort for complex numbers (real part, then imaginary; NaN/inf last)
PSI_HD bool np_less(cx a, cx b) {
    if (a.re < b.re) return a.im == a.im || b.im != b.im;
    if (a.re > b.re) return b.im != b.im && a.im == a.im;
    if (a.re == b.re || (a.re != a.re && b.re != b.re)) return a.im < b.im || (b.im != b.im && a.im == a.im);
    return b.re != b.re;
}

// sort n values ascending (rank sort: every lane places its own elements); result in `out`
template <class W>
PSI_HD void sort_complex(const cx* in, cx* out, int n) {
    const int lane = W::lane();
    W::sync();
    for (int k = lane; k < n; k += W::width) {
        int rank = 0;
        for (int l = 0; l < n; ++l)
            if (np_less(in[l], in[k]) || (!np_less(in[k], in[l]) && l < k)) ++rank;
        out[rank] = in[k];
    }
    W::sync();
}

// Zeros of the approximant: Aberth roots, each polished by scipy's secant rule on r(z) itself
// (50 iterations, tol 1.48e-8; a start is kept when the iteration degenerates or leaves the finite
// numbers), sorted, entries closer than 1e-12 (numpy's lexicographic ">" on the difference) merged
// (complex_roots.py:178-240).  Result in s.roots[0..n); returns n.
template <class W>
PSI_NOINLINE int aaa_zeros(const AAAScratch& s, const AAAState& st) {
    const int lane = W::lane(), nr = st.m - 1;
    if (nr <= 0) return 0;
    if (s.memo_nz[st.slot] >= 0) {                       // same support mask seen before: replay
        const int n = s.memo_nz[st.slot];
        for (int k = lane; k < n; k += W::width) s.roots[k] = s.memo_zero[(size_t)st.slot * s.ncol + k];
        W::sync();
        return n;
    }
    aberth_roots<W>(s, st, true, s.roots);
    for (int k = lane; k < nr; k += W::width) {
        const cx z0 = s.roots[k];
        if (!finite_c(z0)) continue;
        // secant on the approximant (per lane: every lane polishes its own zero)
        cx p0 = z0, p1 = second_point(z0, 1.0e-4);
        cx q0 = aaa_eval(s, st, p0), q1 = aaa_eval(s, st, p1);
        bool ok = finite_c(q0) && finite_c(q1), flagged = false;
        cx root = z0;
        if (ok) {
            if (abs_c(q1) < abs_c(q0)) { cx t = p0; p0 = p1; p1 = t; t = q0; q0 = q1; q1 = t; }
            cx p = p1;
            bool done = false;
            for (int it = 0; it < 50 && !done; ++it) {
                if (eq_c(q1, q0)) { flagged = !eq_c(p1, p0); root = cx(0.5 * (p1.re + p0.re), 0.5 * (p1.im + p0.im)); done = true; break; }
                if (abs_c(q1) > abs_c(q0)) {
                    const cx rr = np_div(q0, q1);
                    p = np_div(np_mul(-rr, p1) + p0, cx(1.0 - rr.re, -rr.im));
                } else {
                    const cx rr = np_div(q1, q0);
                    p = np_div(np_mul(-rr, p0) + p1, cx(1.0 - rr.re, -rr.im));
                }
                if (!finite_c(p)) { ok = false; break; }
                if (abs_c(p - p1) <= 1.48e-8) { root = p; done = true; break; }
                p0 = p1; q0 = q1; p1 = p;
                q1 = aaa_eval(s, st, p1);
                if (!finite_c(q1)) { ok = false; break; }
                root = p;
            }
        }
        if (ok && !flagged) s.roots[k] = root;
    }
    W::sync();
    sort_complex<W>(s.roots, s.tmp, nr);
    // unique_within_tol (every lane computes the same compaction)
    int n = 0;
    if (lane == 0) {
        for (int k = 0; k < nr; ++k) {
            if (k == 0) { s.roots[n++] = s.tmp[0]; continue; }
            const cx d = s.tmp[k] - s.tmp[k - 1];
            if (np_less(cx(1.0e-12, 0.0), d)) s.roots[n++] = s.tmp[k];
        }
    }
    W::sync();
    n = 0;
    for (int k = 0; k < nr; ++k) {
        if (k == 0) { ++n; continue; }
        const cx d = s.tmp[k] - s.tmp[k - 1];
        if (np_less(cx(1.0e-12, 0.0), d)) ++n;
    }
    for (int k = lane; k < n; k += W::width) s.memo_zero[(size_t)st.slot * s.ncol + k] = s.roots[k];
    if (lane == 0) s.memo_nz[st.slot] = n;
    W::sync();
    return n;
}

// One cleanup pass (complex_roots.py:286-329): every pole inside the domain whose residue, measured by
// a 4-point contour integral of side 1e-5, is below clean_tol*max|F| costs the nearest support point.
// Returns the number of support points removed (the fit is redone either way, as in the reference).
template <class W>
PSI_NOINLINE int aaa_cleanup(const AAAScratch& s, AAAState& st) {
    const int lane = W::lane(), np_ = st.m - 1;
    const int slot = st.slot;
    cx* dec = s.memo_pole + (size_t)slot * s.ncol;       // (residue ratio, sample index) per pole in the domain
    if (s.memo_np[slot] < 0) {
        aberth_roots<W>(s, st, false, s.roots);
        const double side = 1.0e-5;
        const double c45 = 0.7071067811865476;
        const cx ring[4] = {cx(side * c45, -side * c45), cx(side * c45, side * c45), cx(-side * c45, side * c45),
                            cx(-side * c45, -side * c45)};
        for (int k = lane; k < np_; k += W::width) {
            const cx pole = s.roots[k];
            cx d(INFINITY, -1.0);
            if (finite_c(pole) && st.dom.is_in(pole)) {
                cx zc[4], fc[4];
                for (int q = 0; q < 4; ++q) { zc[q] = pole + ring[q]; fc[q] = aaa_eval(s, st, zc[q]); }
                cx integral(0.0, 0.0);
                for (int q = 0; q < 4; ++q) {
                    const cx dz = zc[(q + 1) & 3] - zc[q], fs = fc[(q + 1) & 3] + fc[q];
                    integral = integral + (0.5 * dz) * fs;
                }
                int nearest = 0; double bd = INFINITY;
                for (int j = 0; j < st.m; ++j) {
                    const double dd = abs_c(pole - s.Z[s.on[j]]);
                    if (dd < bd) { bd = dd; nearest = j; }
                }
                d = cx(abs_c(integral) / st.fmax, (double)s.on[nearest]);
            }
            dec[k] = d;
        }
        if (lane == 0) s.memo_np[slot] = np_ > 0 ? np_ : 0;
        W::sync();
    }
    const int nd = s.memo_np[slot];
    int removed = 0;
    for (int k = 0; k < nd; ++k) if (dec[k].im >= 0.0 && dec[k].re < st.clean_tol) ++removed;
    W::sync();
    if (lane == 0)
        for (int k = 0; k < nd; ++k) if (dec[k].im >= 0.0 && dec[k].re < st.clean_tol) s.mask[(int)dec[k].im] = 0;
    W::sync();
    aaa_fit<W>(s, st);
    return removed;
}

// RationalApproximation.calculate (complex_roots.py:126-176).  `reuse_greedy`: replay the greedy phase
// stored by the previous call on the same samples (the clean_tol-halving loop of psi_mode.py:174-217
// calls calculate() on identical samples; the greedy phase does not depend on clean_tol).
// Returns false when the samples are not finite (the reference raises ValueError).
template <class W>
PSI_NOINLINE bool aaa_calculate(const AAAScratch& s, AAAState& st, bool reuse_greedy) {
    const int lane = W::lane(), M = st.M;
    if (!reuse_greedy) {
        aaa_memo_clear<W>(s, st);
        double fm = 0.0; bool fin = true;
        for (int i = lane; i < M; i += W::width) {
            fm = fmax(fm, abs_c(s.F[i]));
            if (!finite_c(s.F[i]) || !finite_c(s.Z[i])) fin = false;
        }
        st.fmax = W::max(fm);
        if (W::any(!fin)) return false;
        for (int i = lane; i < M; i += W::width) s.mask[i] = (i < 2) ? 1 : 0;
        W::sync();
        st.v_m = 0;
        aaa_fit<W>(s, st);
        for (int k = 1; k < M / 2; ++k) {
            // promote the worst-fitted sample (first maximum, like numpy.argmax)
            int best = -1; double bv = -1.0;
            for (int i = lane; i < st.r; i += W::width)
                if (s.R[i] > bv) { bv = s.R[i]; best = i; }
            W::argmax(bv, best);
            const int added = s.off[best];
            W::sync();
            if (lane == 0) s.mask[added] = 1;
            W::sync();
            if (aaa_fit<W>(s, st, added) < st.tol) break;
        }
        for (int i = lane; i < M; i += W::width) s.g_mask[i] = s.mask[i];
        W::sync();
    } else {
        // replay: restore the greedy mask; its fit comes out of the memo (or is recomputed if evicted)
        for (int i = lane; i < M; i += W::width) s.mask[i] = s.g_mask[i];
        W::sync();
        aaa_fit<W>(s, st);
    }
    while (aaa_cleanup<W>(s, st) != 0) {}
    return true;
}

}  // namespace psi
