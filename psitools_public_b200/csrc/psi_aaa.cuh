// psi_aaa.cuh -- warp-cooperative AAA rational approximation on the device (kernel K4's building
// blocks): Loewner-matrix fit by Householder QR + inverse iteration, greedy support-point selection,
// Froissart-doublet cleanup, zeros and poles of the barycentric form by Ehrlich-Aberth iteration.
//
// Replaces RationalApproximation (complex_roots.py:38-368), which leans on LAPACK (numpy.linalg.svd,
// scipy.linalg.eig).  The mathematics is the same -- weights = right singular vector of the smallest
// singular value of the Loewner matrix; poles/zeros = roots of sum_j w_j/(z-z_j) and
// sum_j w_j f_j/(z-z_j) -- the numerical kernels are GPU-friendly ones:
//   * the weights come from a Householder QR of the Loewner matrix (lanes split the rows, four columns
//     per pass) followed by inverse iteration on the small triangular factor -- a fraction of the
//     flops of a full SVD and no serial chain of plane rotations;
//   * Aberth's simultaneous iteration works directly on the barycentric sums (O(m) per root and
//     step, one root per lane) instead of a generalised eigenvalue problem.
// One warp owns one approximation; all arrays live in a per-warp scratch area in global memory (L2).
// Written against the warp policy W like psi_quad.cuh, so the CPU harness runs the same code.
#pragma once
#include "psi_batch.cuh"

namespace psi {

#if defined(PSI_AAA_STATS)
// test-harness counters: fits, inverse iterations, -, Aberth calls, Aberth iterations
static long long g_aaa_stats[16];
#define PSI_STAT(i, n) (g_aaa_stats[i] += (n))
// test-harness switch: run the general (memory) form of the triangular solves also for small systems
static bool g_aaa_general_solves = false;
#define PSI_GENERAL_SOLVES g_aaa_general_solves
#else
#define PSI_STAT(i, n)
#define PSI_GENERAL_SOLVES false
#endif

// Phase profiling of the analysis stage (build with -DPSI_PROF: tools/prof_analysis.py): cycles per phase,
// summed over warps with clock64() on lane 0.
#if defined(PSI_PROF) && defined(__CUDACC__)
static __device__ unsigned long long g_psi_prof[32];
#endif
#if defined(PSI_PROF) && defined(__CUDA_ARCH__)
#define PROF_T0(v) const long long v = clock64()
#define PROF_ADD(id, v) do { if ((threadIdx.x & 31) == 0) atomicAdd(&g_psi_prof[id], (unsigned long long)(clock64() - (v))); } while (0)
#define PROF_MAX(id, val_) do { if ((threadIdx.x & 31) == 0) atomicMax(&g_psi_prof[id], (unsigned long long)(val_)); } while (0)
#else
#define PROF_T0(v)
#define PROF_ADD(id, v)
#define PROF_MAX(id, val_)
#endif

// Out-of-line copies of the small helpers.  The analysis kernel is instruction-fetch bound (ncu: 83 % of
// its stall samples are "no instruction" with 315 KB of SASS and 24 warps per SM in different phases), so
// every helper exists once instead of at each use, and the strided loops are not unrolled.
#if defined(__CUDACC__)
#define PSI_OUTLINE __host__ __device__ __noinline__
#else
#define PSI_OUTLINE inline
#endif
PSI_OUTLINE cx a_recip(cx a) { return recip(a); }
PSI_OUTLINE double a_abs(cx a) { return abs_c(a); }
PSI_OUTLINE cx a_div(cx a, cx b) { return np_div(a, b); }
template <class W> PSI_OUTLINE double a_sum(double v) { return W::sum(v); }
template <class W> PSI_OUTLINE double a_max(double v) { return W::max(v); }

constexpr int kMemoSlots = 12;
constexpr int kMaxSamples = 512;     // largest sample count (cap) the device AAA accepts

// scratch layout of one warp (all sizes in elements, cap = maximum number of samples)
struct AAAScratch {
    cx* Z;        // [cap]   sample points
    cx* F;        // [cap]   function samples
    int* mask;    // [cap]   1 = support point
    cx* A;        // [cap * ncol]  Loewner matrix, column-major (destroyed by the QR)
    cx* V;        // [ncol * ncol] triangular factor R of the Loewner matrix
    cx* V2;       // [4 * ncol]    vectors of the inverse iteration
    cx* w;        // [ncol]  weights
    double* R;    // [cap]   residuals at the non-support points (in sample order of those points)
    int* on;      // [ncol]  sample index of the k-th support point
    int* off;     // [cap]   sample index of the k-th non-support point
    cx* roots;    // [ncol]  Aberth iterates (zeros or poles)
    cx* tmp;      // [ncol]  second root buffer
    // greedy-phase result, replayed by PSIMode's clean_tol-halving loop
    int* g_mask;  // [cap]
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
    int cap, ncol;
};

// Scratch is split in two: the ITEM part (everything that must survive between the phase kernels of the
// analysis stage: index lists, weights, residuals, zeros, the memo) and the WORK part (Loewner matrix, triangular
// factor, vectors of the inverse iteration: only alive inside one fit).  The GPU pipeline keeps one item part
// per search and one work part per group of the fit kernel; the single-warp form (mode_run, CPU harness) carves
// both from one area.
PSI_HD size_t aaa_item_bytes(int cap) {
    const size_t ncol = (size_t)cap / 2 + 2;
    return sizeof(cx) * ((3 + 3 * kMemoSlots) * ncol) +
           sizeof(double) * ((1 + kMemoSlots) * (size_t)cap + kMemoSlots) +
           sizeof(int) * ((3 + kMemoSlots) * (size_t)cap + ncol + 3 * kMemoSlots) + 64;
}
PSI_HD size_t aaa_work_bytes(int cap) {
    const size_t ncol = (size_t)cap / 2 + 2;
    return sizeof(cx) * ((size_t)cap * ncol + ncol * ncol + 4 * ncol) + 64;
}
PSI_HD size_t aaa_scratch_bytes(int cap) {
    const size_t ncol = (size_t)cap / 2 + 2;
    return ((aaa_item_bytes(cap) + 255) & ~(size_t)255) + aaa_work_bytes(cap) + sizeof(cx) * (2 * (size_t)cap + 2 * ncol) + 256;
}

PSI_HD void aaa_carve_item(void* base, int cap, AAAScratch& s) {
    const size_t ncol = (size_t)cap / 2 + 2;
    s.cap = cap; s.ncol = (int)ncol;
    cx* c = reinterpret_cast<cx*>(base);
    s.w = c; c += ncol;
    s.roots = c; c += ncol;
    s.tmp = c; c += ncol;
    s.memo_w = c; c += (size_t)kMemoSlots * ncol;
    s.memo_pole = c; c += (size_t)kMemoSlots * ncol;
    s.memo_zero = c; c += (size_t)kMemoSlots * ncol;
    double* d = reinterpret_cast<double*>(c);
    s.R = d; d += cap;
    s.memo_R = d; d += (size_t)kMemoSlots * cap;
    s.memo_res = d; d += kMemoSlots;
    int* i = reinterpret_cast<int*>(d);
    s.mask = i; i += cap;
    s.off = i; i += cap;
    s.g_mask = i; i += cap;
    s.memo_mask = i; i += (size_t)kMemoSlots * cap;
    s.memo_m = i; i += kMemoSlots;
    s.memo_np = i; i += kMemoSlots;
    s.memo_nz = i; i += kMemoSlots;
    s.on = i;
}
PSI_HD void aaa_carve_work(void* base, int cap, AAAScratch& s) {
    const size_t ncol = (size_t)cap / 2 + 2;
    cx* c = reinterpret_cast<cx*>(base);
    s.A = c; c += (size_t)cap * ncol;
    s.V = c; c += ncol * ncol;
    s.V2 = c;
}

// one area for everything, incl. sample buffers and polish lists (single-warp form of the search)
PSI_HD AAAScratch aaa_carve(void* base, int cap) {
    AAAScratch s;
    const size_t ncol = (size_t)cap / 2 + 2;
    unsigned char* b = reinterpret_cast<unsigned char*>(base);
    aaa_carve_item(b, cap, s);
    b += (aaa_item_bytes(cap) + 255) & ~(size_t)255;
    aaa_carve_work(b, cap, s);
    cx* c = s.V2 + 4 * ncol;
    s.Z = c; c += cap;
    s.F = c; c += cap;
    s.w0 = c; c += ncol;
    s.pol = c;
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
    int v_m;        // s.w holds the weights of the last COMPUTED fit with v_m support points (0 = nothing
                    // usable); lets the next greedy step start its inverse iteration warm
};

template <class W>
PSI_HD void aaa_memo_clear(const AAAScratch& s, AAAState& st) {
    for (int k = W::lane(); k < kMemoSlots; k += W::width) s.memo_m[k] = -1;
    st.memo_next = 0;
    W::sync();
}

// rebuild the index lists from the mask: one ballot per chunk of `width` samples, every lane places its
// own sample behind the support / non-support points counted so far
template <class W>
PSI_OUTLINE void aaa_index(const AAAScratch& s, AAAState& st) {
    W::sync();
    const int lane = W::lane();
    int m = 0, r = 0;
#pragma unroll 1
    for (int base = 0; base < st.M; base += W::width) {
        const int i = base + lane;
        const bool in = i < st.M;
        const bool on = in && s.mask[i] != 0;
        const unsigned set = W::ballot(on), all = W::ballot(in);
        const unsigned below = lane == 0 ? 0u : (0xffffffffu >> (32 - lane));
        if (on) s.on[m + W::popc(set & below)] = i;
        else if (in) s.off[r + W::popc(all & ~set & below)] = i;
        m += W::popc(set);
        r += W::popc(all & ~set);
    }
    W::sync();
    st.m = m; st.r = r;
}

// Right singular vector of the smallest singular value of the r x m matrix A (column-major, leading
// dimension r; destroyed) -- the AAA weights.  Two stages:
//   1. Householder QR, lanes own rows i = lane (mod width), each reflector applied to four columns at a
//      time (eight independent warp reductions in flight): A -> m x m triangle R (zero rows if r < m),
//      written to Rm.  Column-wise backward stable, the same error class as LAPACK's zgesdd that the
//      reference uses.
//   2. inverse iteration on R^H R (two triangular solves per step, column oriented: the lanes update
//      their own entries of the right-hand side, the pivot entry is re-read by everybody), started from
//      `x0` (the previous weights of the greedy chain, or a generic vector) and stopped when the unit
//      vector moves by less than 1e-12 or stops contracting; pivots below eps*max|R_kk| are lifted to
//      that value, which turns an exactly singular R (wide Loewner matrix) into "some null vector".
// vec: 3*m scratch entries.  The result (unit norm) is written to w.  Returns the iteration count.
template <class W>
PSI_NOINLINE int loewner_min_vector(cx* A, cx* Rm, cx* vec, cx* w, int r, int m, bool have_x0) {
    const int lane = W::lane();
    const int steps = r < m ? r : m;
    PSI_STAT(0, 1);
    PSI_STAT(5, (long long)r * m * m);
    PSI_STAT(8 + (m <= 10 ? 0 : m <= 20 ? 1 : m <= 32 ? 2 : 3), 1);
    PROF_T0(t_qr);
    // (A left-looking variant with a four-column panel in registers -- 12x less trailing-matrix traffic -- was
    // measured slower, 0.166 s vs 0.146 s per 256^2 map: its serial chain of load / butterfly / update per reflector
    // is as long, and the extra registers and code cost more than the traffic saved.)
    for (int k = 0; k < steps; ++k) {
        cx* ak = A + (size_t)k * r;
        const int i0 = lane >= k ? lane : lane + ((k - lane + W::width - 1) / W::width) * W::width;   // first own row >= k
        double n2 = 0.0;
#pragma unroll 1
        for (int i = i0; i < r; i += W::width) n2 += ak[i].re * ak[i].re + ak[i].im * ak[i].im;
        n2 = a_sum<W>(n2);
        const cx x0 = ak[k];
        const double nx = sqrt(n2);
        if (!(nx > 0.0)) {                                   // zero column: R_kk = 0, nothing to reflect
            if (lane == 0) Rm[k + (size_t)k * m] = cx(0.0, 0.0);
            W::sync();
            continue;
        }
        const double a0 = sqrt(x0.re * x0.re + x0.im * x0.im);
        const cx ph = a0 > 0.0 ? cx(x0.re / a0, x0.im / a0) : cx(1.0, 0.0);
        // H = I - tau v v^H with v = x + ph*nx*e_0 (no cancellation), H x = -ph*nx*e_0
        const cx v0(ph.re * (a0 + nx), ph.im * (a0 + nx));
        const double tau = 1.0 / (nx * (nx + a0));
        if (lane == 0) Rm[k + (size_t)k * m] = cx(-ph.re * nx, -ph.im * nx);
        for (int j = k + 1; j < m; j += 4) {
            const int nj = m - j < 4 ? m - j : 4;
            double dr[4] = {0.0, 0.0, 0.0, 0.0}, di[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
            for (int i = i0; i < r; i += W::width) {
                const cx v = i == k ? v0 : ak[i];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c < nj) {
                        const cx a = A[i + (size_t)(j + c) * r];
                        dr[c] += v.re * a.re + v.im * a.im;                  // conj(v)*a
                        di[c] += v.re * a.im - v.im * a.re;
                    }
                }
            }
            W::sum4c(dr, di);
#pragma unroll 1
            for (int i = i0; i < r; i += W::width) {
                const cx v = i == k ? v0 : ak[i];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c < nj) {
                        const cx f(tau * dr[c], tau * di[c]);
                        cx* a = A + i + (size_t)(j + c) * r;
                        *a = *a - f * v;
                    }
                }
            }
        }
        W::sync();
    }
    PROF_ADD(3, t_qr);
    if (m > 48) { PROF_ADD(20, t_qr); PROF_MAX(21, m); }
    PROF_T0(t_ii);
    // R: strict upper triangle from A (rows < r), diagonal already in place (zero where no step ran)
#pragma unroll 1
    for (int e = lane; e < m * m; e += W::width) {
        const int i = e % m, j = e / m;
        if (i < j) Rm[e] = i < r ? A[i + (size_t)j * r] : cx(0.0, 0.0);
        else if (i > j || i >= steps) Rm[e] = cx(0.0, 0.0);
    }
    W::sync();
    double dmax = 0.0;
    for (int k = 0; k < m; ++k) dmax = fmax(dmax, fabs(Rm[k + (size_t)k * m].re) + fabs(Rm[k + (size_t)k * m].im));
    if (!(dmax > 0.0) || !(dmax < INFINITY)) {              // A = 0 (or not finite): any unit vector
#pragma unroll 1
        for (int j = lane; j < m; j += W::width) w[j] = cx(j == m - 1 ? 1.0 : 0.0, 0.0);
        W::sync();
        return 0;
    }
    const double piv = 2.220446049250313e-16 * dmax;
    cx* rinv = vec;              // reciprocals of the (guarded) diagonal
    cx* b = vec + m;             // right-hand side, consumed by the solves
    cx* y = vec + 2 * m;         // solution of the current solve
#pragma unroll 1
    for (int k = lane; k < m; k += W::width) {
        const cx d = Rm[k + (size_t)k * m];
        rinv[k] = (fabs(d.re) + fabs(d.im) < piv) ? cx(1.0 / piv, 0.0) : a_recip(d);
        b[k] = have_x0 ? w[k] : cx(1.0 + 0.25 * (k % 3), 0.125 * (k % 5));
    }
    W::sync();
    {   // normalise the start vector; keep a copy in w for the convergence test
        double n2 = 0.0;
#pragma unroll 1
        for (int k = lane; k < m; k += W::width) n2 += b[k].re * b[k].re + b[k].im * b[k].im;
        n2 = a_sum<W>(n2);
        const double sc = n2 > 0.0 ? 1.0 / sqrt(n2) : 0.0;
#pragma unroll 1
        for (int k = lane; k < m; k += W::width) {
            const cx v = n2 > 0.0 ? cx(b[k].re * sc, b[k].im * sc) : cx(k == 0 ? 1.0 : 0.0, 0.0);
            b[k] = v; w[k] = v;
        }
        W::sync();
    }
    int it = 0;
    double prev = INFINITY;
    if (W::width > 1 && m <= W::width && !PSI_GENERAL_SOLVES) {
        // Small systems (the common case: m <= 10 for 20 samples): lane L owns entry L of the right-hand side in
        // registers and the pivot entry is broadcast with shuffles -- the same operations in the same order as
        // the general loop below (bit-identical results) without its store / warp barrier / load per step.
        const bool own = lane < m;
        cx bl = own ? b[lane] : cx(0.0, 0.0);
        for (; it < 80; ++it) {
            PSI_STAT(1, 1);
            PSI_STAT(6, (long long)m * m);
            cx yl(0.0, 0.0);
#pragma unroll 1
            for (int j = 0; j < m; ++j) {                      // R^H y = b
                const cx rj = rinv[j];
                const cx yj = cx(W::bcast(bl.re, j), W::bcast(bl.im, j)) * cx(rj.re, -rj.im);
                if (lane == j) yl = yj;
                if (own && lane > j) {
                    const cx rji = Rm[j + (size_t)lane * m];
                    bl = bl - cx(rji.re, -rji.im) * yj;
                }
            }
            {
                const double n2 = a_max<W>(own ? fabs(yl.re) + fabs(yl.im) : 0.0);
                const double sc = n2 > 0.0 ? 1.0 / n2 : 1.0;
                bl = cx(yl.re * sc, yl.im * sc);
            }
#pragma unroll 1
            for (int j = m - 1; j >= 0; --j) {                 // R z = b
                const cx zj = cx(W::bcast(bl.re, j), W::bcast(bl.im, j)) * rinv[j];
                if (lane == j) yl = zj;
                if (lane < j) bl = bl - Rm[lane + (size_t)j * m] * zj;
            }
            const double n2 = a_max<W>(own ? fabs(yl.re) + fabs(yl.im) : 0.0);
            const double s1 = n2 > 0.0 ? 1.0 / n2 : 1.0;
            double q2 = 0.0;
            if (own) {
                const cx v(yl.re * s1, yl.im * s1);
                q2 += v.re * v.re + v.im * v.im;
            }
            q2 = a_sum<W>(q2);
            const double sc = s1 / sqrt(q2);
            double d2 = 0.0;
            if (own) {
                const cx v(yl.re * sc, yl.im * sc);
                const cx dv = v - w[lane];
                d2 += dv.re * dv.re + dv.im * dv.im;
                w[lane] = v; bl = v;
            }
            d2 = a_sum<W>(d2);
            if (d2 <= 1.0e-24 || (d2 <= 1.0e-12 && d2 > 0.25 * prev)) { ++it; break; }
            prev = d2;
        }
        W::sync();
        PROF_ADD(4, t_ii);
        return it;
    }
    for (; it < 80; ++it) {
        PSI_STAT(1, 1);
        PSI_STAT(6, (long long)m * m);
        // R^H y = b  (forward; L_ij = conj(R_ji), row j of R is strided in memory)
        for (int j = 0; j < m; ++j) {
            const cx rj = rinv[j];
            const cx yj = b[j] * cx(rj.re, -rj.im);
            if (lane == 0) y[j] = yj;
            const int i1 = lane > j ? lane : lane + ((j + 1 - lane + W::width - 1) / W::width) * W::width;
#pragma unroll 1
            for (int i = i1; i < m; i += W::width) {
                const cx rji = Rm[j + (size_t)i * m];
                b[i] = b[i] - cx(rji.re, -rji.im) * yj;
            }
            W::sync();
        }
        {   // rescale (the solve can grow by 1/piv) and move y -> b
            double n2 = 0.0;
#pragma unroll 1
            for (int k = lane; k < m; k += W::width) n2 = fmax(n2, fabs(y[k].re) + fabs(y[k].im));
            n2 = a_max<W>(n2);
            const double sc = n2 > 0.0 ? 1.0 / n2 : 1.0;
#pragma unroll 1
            for (int k = lane; k < m; k += W::width) b[k] = cx(y[k].re * sc, y[k].im * sc);
            W::sync();
        }
        // R z = b  (backward; column j of R is contiguous)
        for (int j = m - 1; j >= 0; --j) {
            const cx zj = b[j] * rinv[j];
            if (lane == 0) y[j] = zj;
#pragma unroll 1
            for (int i = lane; i < j; i += W::width) b[i] = b[i] - Rm[i + (size_t)j * m] * zj;
            W::sync();
        }
        double n2 = 0.0;
#pragma unroll 1
        for (int k = lane; k < m; k += W::width) {
            // scale first: the entries can be ~1/piv^2
            n2 = fmax(n2, fabs(y[k].re) + fabs(y[k].im));
        }
        n2 = a_max<W>(n2);
        const double s1 = n2 > 0.0 ? 1.0 / n2 : 1.0;
        double q2 = 0.0;
#pragma unroll 1
        for (int k = lane; k < m; k += W::width) {
            const cx v(y[k].re * s1, y[k].im * s1);
            q2 += v.re * v.re + v.im * v.im;
        }
        q2 = a_sum<W>(q2);
        const double sc = s1 / sqrt(q2);
        double d2 = 0.0;
#pragma unroll 1
        for (int k = lane; k < m; k += W::width) {
            const cx v(y[k].re * sc, y[k].im * sc);
            const cx dv = v - w[k];
            d2 += dv.re * dv.re + dv.im * dv.im;
            w[k] = v; b[k] = v;
        }
        d2 = a_sum<W>(d2);
        W::sync();
        if (d2 <= 1.0e-24 || (d2 <= 1.0e-12 && d2 > 0.25 * prev)) { ++it; break; }
        prev = d2;
    }
    PROF_ADD(4, t_ii);
    return it;
}

// Loewner fit for the current mask: weights and residuals (complex_roots.py:56-90), in two halves so that the
// pipeline can run the cheap half (index lists + memo lookup) in its control kernel and the expensive half
// (Loewner matrix, QR, inverse iteration, residuals) in the fit kernel.
// aaa_fit_lookup: rebuilds the index lists; true = the mask was fitted before, weights/residuals restored and
// *res = largest residual / max|F|.
template <class W>
PSI_NOINLINE bool aaa_fit_lookup(const AAAScratch& s, AAAState& st, double* res) {
    PROF_T0(t_ix);
    aaa_index<W>(s, st);
    PROF_ADD(0, t_ix);
    const int lane = W::lane(), r = st.r, m = st.m;
    PROF_T0(t_ml);
    for (int k = 0; k < kMemoSlots; ++k) {
        if (s.memo_m[k] != m) continue;
        bool same = true;
        const int* mk = s.memo_mask + (size_t)k * s.cap;
#pragma unroll 1
        for (int i = lane; i < st.M; i += W::width) if (mk[i] != s.mask[i]) same = false;
        if (W::any(!same)) continue;
#pragma unroll 1
        for (int j = lane; j < m; j += W::width) s.w[j] = s.memo_w[(size_t)k * s.ncol + j];
#pragma unroll 1
        for (int i = lane; i < r; i += W::width) s.R[i] = s.memo_R[(size_t)k * s.cap + i];
        W::sync();
        st.slot = k;
        st.v_m = 0;                                  // V belongs to some other mask now
        PSI_STAT(2, 1);
        PROF_ADD(1, t_ml);
        *res = s.memo_res[k];
        return true;
    }
    PROF_ADD(1, t_ml);
    return false;
}

// aaa_fit_compute: the fit itself (index lists of the current mask must be in place: aaa_fit_lookup).  Returns
// the largest residual divided by max|F| and files the fit in the memo.
template <class W>
PSI_NOINLINE double aaa_fit_compute(const AAAScratch& s, AAAState& st, int added_sample) {
    const int lane = W::lane(), r = st.r, m = st.m;
    PROF_T0(t_bd);
    for (int j = 0; j < m; ++j) {
        const cx zj = s.Z[s.on[j]], fj = s.F[s.on[j]];
#pragma unroll 1
        for (int i = lane; i < r; i += W::width) {
            const cx c = a_recip(s.Z[s.off[i]] - zj);
            s.A[i + (size_t)j * r] = (s.F[s.off[i]] - fj) * c;        // F_i C_ij - C_ij f_j
        }
    }
    W::sync();
    PROF_ADD(2, t_bd);
    if (r == 0) {
#pragma unroll 1
        for (int j = lane; j < m; j += W::width) s.w[j] = cx(j == m - 1 ? 1.0 : 0.0, 0.0);
    } else {
        // Start vector of the inverse iteration: along the greedy chain (this mask = previous mask + one
        // support point) the previous weights, with a generic entry for the new column.
        bool have_x0 = false;
        if (added_sample >= 0 && st.v_m == m - 1 && m >= 3) {
            int pnew = 0;
            while (pnew < m && s.on[pnew] != added_sample) ++pnew;
            if (pnew < m) {
                cx* old = s.V2 + 3 * (size_t)s.ncol;
#pragma unroll 1
                for (int j = lane; j < m - 1; j += W::width) old[j] = s.w[j];
                W::sync();
                const double gen = 0.5 / sqrt((double)m);
                for (int j = lane; j < m; j += W::width)
                    s.w[j] = j == pnew ? cx(gen, 0.5 * gen) : old[j - (j > pnew)];
                W::sync();
                have_x0 = true;
            }
        }
        loewner_min_vector<W>(s.A, s.V, s.V2, s.w, r, m, have_x0);
        st.v_m = m;
    }
    W::sync();
    PROF_T0(t_rs);
    double rmax = 0.0;
#pragma unroll 1
    for (int i = lane; i < r; i += W::width) {
        const cx Zi = s.Z[s.off[i]];
        cx num(0.0, 0.0), den(0.0, 0.0);
        for (int j = 0; j < m; ++j) {
            const cx t = s.w[j] * a_recip(Zi - s.Z[s.on[j]]);
            num = num + t * s.F[s.on[j]];
            den = den + t;
        }
        const double res = a_abs(s.F[s.off[i]] - num * a_recip(den));
        s.R[i] = res;
        rmax = fmax(rmax, res);
    }
    W::sync();
    const double resid = a_max<W>(rmax) / st.fmax;
    PROF_ADD(5, t_rs);
    PROF_T0(t_mi);
    {   // memo insert
        const int k = st.memo_next;
        st.memo_next = (k + 1) % kMemoSlots;
#pragma unroll 1
        for (int i = lane; i < st.M; i += W::width) s.memo_mask[(size_t)k * s.cap + i] = s.mask[i];
#pragma unroll 1
        for (int j = lane; j < m; j += W::width) s.memo_w[(size_t)k * s.ncol + j] = s.w[j];
#pragma unroll 1
        for (int i = lane; i < r; i += W::width) s.memo_R[(size_t)k * s.cap + i] = s.R[i];
        if (lane == 0) { s.memo_m[k] = m; s.memo_res[k] = resid; s.memo_np[k] = -1; s.memo_nz[k] = -1; }
        W::sync();
        st.slot = k;
    }
    PROF_ADD(6, t_mi);
    return resid;
}

template <class W>
PSI_HD double aaa_fit(const AAAScratch& s, AAAState& st, int added_sample = -1) {
    double res;
    if (aaa_fit_lookup<W>(s, st, &res)) return res;
    return aaa_fit_compute<W>(s, st, added_sample);
}

// barycentric evaluation r(z) = sum w_j f_j/(z-z_j) / sum w_j/(z-z_j)   (complex_roots.py:331-368)
PSI_OUTLINE cx aaa_eval(const AAAScratch& s, const AAAState& st, cx z) {
    cx num(0.0, 0.0), den(0.0, 0.0);
    for (int j = 0; j < st.m; ++j) {
        const cx t = s.w[j] * a_recip(z - s.Z[s.on[j]]);
        num = num + t * s.F[s.on[j]];
        den = den + t;
    }
    return num * a_recip(den);
}

// Roots of S(z) = sum_j a_j/(z - z_j) (a_j = w_j for the poles, w_j f_j for the zeros) by the
// Ehrlich-Aberth iteration on p(z) = S(z) prod(z - z_j): p/p' = 1/(S'/S + sum 1/(z - z_j)).
// m - 1 roots are written to out[]; roots that run away (|z| > 1e12 * scale: the leading coefficient
// sum a_j is numerically zero) are returned as +inf.  Returns the number of iterations used.
template <class W>
PSI_NOINLINE int aberth_roots(const AAAScratch& s, const AAAState& st, bool with_f, cx* out) {
    const int lane = W::lane(), m = st.m, nr = m - 1;
    if (nr <= 0) return 0;
    PROF_T0(t_ab);
    // scale and centre of the support points
    double cre = 0.0, cim = 0.0;
    for (int j = 0; j < m; ++j) { cre += s.Z[s.on[j]].re; cim += s.Z[s.on[j]].im; }
    cre /= m; cim /= m;
    double scale = 0.0;
    for (int j = 0; j < m; ++j) scale = fmax(scale, a_abs(s.Z[s.on[j]] - cx(cre, cim)));
    if (!(scale > 0.0)) scale = 1.0;
    // start values: between consecutive support points, nudged off the connecting line by a fraction of THEIR
    // distance.  (The samples around guess roots and in zoom boxes form clusters 1e-5 of the domain wide that hold
    // as many roots as support points; a nudge on the scale of the whole set starts all of them far outside their
    // cluster, from where it looks like one root of high multiplicity and the iteration crawls.)
#pragma unroll 1
    for (int k = lane; k < nr; k += W::width) {
        const cx a = s.Z[s.on[k]], b = s.Z[s.on[k + 1]];
        const cx d = b - a;
        const double len = a_abs(d);
        out[k] = cx(0.5 * (a.re + b.re) - 0.31 * d.im + 0.013 * len, 0.5 * (a.im + b.im) + 0.31 * d.re + 0.017 * len);
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
        PSI_STAT(7, (long long)nr * (m + nr));
        bool moving = false;
        int slot = 0;
#pragma unroll 1
        for (int k = lane; k < nr; k += W::width, ++slot) {
            const cx z = out[k];
            const int q = slot < 8 ? slot : 7;
            if (!finite_c(z) || frozen[q]) { s.tmp[k] = z; continue; }
            cx S(0.0, 0.0), S1(0.0, 0.0), T(0.0, 0.0);
            bool hit = false;
            for (int j = 0; j < m; ++j) {
                const cx dz = z - s.Z[s.on[j]];
                if (dz.re == 0.0 && dz.im == 0.0) { hit = true; break; }
                const cx inv = a_recip(dz);
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
            const cx newton = a_recip(S1 * a_recip(S) + T);      // p/p'
            cx rep(0.0, 0.0);
            for (int l = 0; l < nr; ++l) {
                if (l == k) continue;
                const cx zl = out[l];
                if (!finite_c(zl)) continue;
                const cx dz = z - zl;
                if (dz.re == 0.0 && dz.im == 0.0) continue;
                rep = rep + a_recip(dz);
            }
            const cx step = a_div(newton, cx(1.0, 0.0) - newton * rep);
            cx zn = z - step;
            const double sl = a_abs(step);
            if (!finite_c(zn) || a_abs(zn - cx(cre, cim)) > 1.0e12 * scale) {
                zn = cx(INFINITY, 0.0);
            } else if (sl <= 1.0e-15 * a_abs(zn) ||
                       (sl <= 1.0e-9 * (a_abs(zn) + scale) && sl > 0.1 * prev_step[q])) {
                if (slot < 8) frozen[q] = true;
                if (sl > prev_step[q]) zn = z;              // keep the better iterate
            } else {
                moving = true;
            }
            prev_step[q] = sl;
            s.tmp[k] = zn;
        }
        W::sync();
#pragma unroll 1
        for (int k = lane; k < nr; k += W::width) out[k] = s.tmp[k];
        W::sync();
        if (!W::any(moving)) break;
    }
    PROF_ADD(with_f ? 8 : 7, t_ab);
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
#pragma unroll 1
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
// aaa_zeros_load: the zeros of a support mask seen before (true; *n = their number), or false.
template <class W>
PSI_HD bool aaa_zeros_load(const AAAScratch& s, const AAAState& st, int* n_out) {
    const int lane = W::lane(), nr = st.m - 1;
    if (nr <= 0) { *n_out = 0; return true; }
    if (s.memo_nz[st.slot] < 0) return false;
    const int n = s.memo_nz[st.slot];
#pragma unroll 1
    for (int k = lane; k < n; k += W::width) s.roots[k] = s.memo_zero[(size_t)st.slot * s.ncol + k];
    W::sync();
    *n_out = n;
    return true;
}

template <class W>
PSI_NOINLINE int aaa_zeros_compute(const AAAScratch& s, const AAAState& st) {
    const int lane = W::lane(), nr = st.m - 1;
    if (nr <= 0) return 0;
    aberth_roots<W>(s, st, true, s.roots);
    PROF_T0(t_zs);
#pragma unroll 1
    for (int k = lane; k < nr; k += W::width) {
        const cx z0 = s.roots[k];
        if (!finite_c(z0)) continue;
        // secant on the approximant (per lane: every lane polishes its own zero)
        cx p0 = z0, p1 = second_point(z0, 1.0e-4);
        cx q0 = aaa_eval(s, st, p0), q1 = aaa_eval(s, st, p1);
        bool ok = finite_c(q0) && finite_c(q1), flagged = false;
        cx root = z0;
        if (ok) {
            if (a_abs(q1) < a_abs(q0)) { cx t = p0; p0 = p1; p1 = t; t = q0; q0 = q1; q1 = t; }
            cx p = p1;
            bool done = false;
            for (int it = 0; it < 50 && !done; ++it) {
                if (eq_c(q1, q0)) { flagged = !eq_c(p1, p0); root = cx(0.5 * (p1.re + p0.re), 0.5 * (p1.im + p0.im)); done = true; break; }
                if (a_abs(q1) > a_abs(q0)) {
                    const cx rr = a_div(q0, q1);
                    p = a_div(np_mul(-rr, p1) + p0, cx(1.0 - rr.re, -rr.im));
                } else {
                    const cx rr = a_div(q1, q0);
                    p = a_div(np_mul(-rr, p0) + p1, cx(1.0 - rr.re, -rr.im));
                }
                if (!finite_c(p)) { ok = false; break; }
                if (a_abs(p - p1) <= 1.48e-8) { root = p; done = true; break; }
                p0 = p1; q0 = q1; p1 = p;
                q1 = aaa_eval(s, st, p1);
                if (!finite_c(q1)) { ok = false; break; }
                root = p;
            }
        }
        if (ok && !flagged) s.roots[k] = root;
    }
    W::sync();
    PROF_ADD(9, t_zs);
    PROF_T0(t_so);
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
#pragma unroll 1
    for (int k = lane; k < n; k += W::width) s.memo_zero[(size_t)st.slot * s.ncol + k] = s.roots[k];
    if (lane == 0) s.memo_nz[st.slot] = n;
    W::sync();
    PROF_ADD(10, t_so);
    return n;
}

template <class W>
PSI_HD int aaa_zeros(const AAAScratch& s, const AAAState& st) {
    int n;
    if (aaa_zeros_load<W>(s, st, &n)) return n;
    return aaa_zeros_compute<W>(s, st);
}

// One cleanup pass (complex_roots.py:286-329): every pole inside the domain whose residue, measured by
// a 4-point contour integral of side 1e-5, is below clean_tol*max|F| costs the nearest support point.
// aaa_poles_compute: poles of the current fit and, per pole inside the domain, (residue ratio, sample to drop),
// filed under the fit's memo slot (the decisions do not depend on clean_tol).
template <class W>
PSI_NOINLINE void aaa_poles_compute(const AAAScratch& s, AAAState& st) {
    const int lane = W::lane(), np_ = st.m - 1;
    const int slot = st.slot;
    cx* dec = s.memo_pole + (size_t)slot * s.ncol;
    aberth_roots<W>(s, st, false, s.roots);
    PROF_T0(t_cr);
    const double side = 1.0e-5;
    const double c45 = 0.7071067811865476;
    const cx ring[4] = {cx(side * c45, -side * c45), cx(side * c45, side * c45), cx(-side * c45, side * c45),
                        cx(-side * c45, -side * c45)};
#pragma unroll 1
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
                const double dd = a_abs(pole - s.Z[s.on[j]]);
                if (dd < bd) { bd = dd; nearest = j; }
            }
            d = cx(a_abs(integral) / st.fmax, (double)s.on[nearest]);
        }
        dec[k] = d;
    }
    if (lane == 0) s.memo_np[slot] = np_ > 0 ? np_ : 0;
    W::sync();
    PROF_ADD(11, t_cr);
}

PSI_HD bool aaa_poles_ready(const AAAScratch& s, const AAAState& st) { return s.memo_np[st.slot] >= 0; }

// aaa_cleanup_apply: drop the support points the stored decisions condemn at the current clean_tol; returns
// their number (the fit is redone either way, as in the reference).
template <class W>
PSI_HD int aaa_cleanup_apply(const AAAScratch& s, AAAState& st) {
    const int lane = W::lane(), slot = st.slot;
    const cx* dec = s.memo_pole + (size_t)slot * s.ncol;
    const int nd = s.memo_np[slot];
    int removed = 0;
    for (int k = 0; k < nd; ++k) if (dec[k].im >= 0.0 && dec[k].re < st.clean_tol) ++removed;
    W::sync();
    if (lane == 0)
        for (int k = 0; k < nd; ++k) if (dec[k].im >= 0.0 && dec[k].re < st.clean_tol) s.mask[(int)dec[k].im] = 0;
    W::sync();
    return removed;
}

template <class W>
PSI_NOINLINE int aaa_cleanup(const AAAScratch& s, AAAState& st) {
    if (!aaa_poles_ready(s, st)) aaa_poles_compute<W>(s, st);
    const int removed = aaa_cleanup_apply<W>(s, st);
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
#pragma unroll 1
        for (int i = lane; i < M; i += W::width) {
            fm = fmax(fm, a_abs(s.F[i]));
            if (!finite_c(s.F[i]) || !finite_c(s.Z[i])) fin = false;
        }
        st.fmax = a_max<W>(fm);
        if (W::any(!fin)) return false;
#pragma unroll 1
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
#pragma unroll 1
        for (int i = lane; i < M; i += W::width) s.g_mask[i] = s.mask[i];
        W::sync();
    } else {
        // replay: restore the greedy mask; its fit comes out of the memo (or is recomputed if evicted)
#pragma unroll 1
        for (int i = lane; i < M; i += W::width) s.mask[i] = s.g_mask[i];
        W::sync();
        aaa_fit<W>(s, st);
    }
    while (aaa_cleanup<W>(s, st) != 0) {}
    return true;
}

}  // namespace psi
