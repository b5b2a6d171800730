// psi_engine.cu -- native (C++) lock-step engine for batches of PSIMode root searches and contour
// root counts.  HOST code: it owns the per-item state machines and calls kernel K1 once per pass for
// the union of all points the items need (no per-point host round trip, no Python in the loop).
//
// What it restates (file:line under psitools/):
//   Rectangle sampling                complex_roots.py:404-409   (numpy legacy MT19937 stream)
//   RationalApproximation (AAA)       complex_roots.py:38-368    (LAPACK zgesdd / zggev: the SAME
//                                     OpenBLAS build numpy/scipy load, bound at run time with dlopen)
//   ClosedPath.count_roots            complex_roots.py:517-625
//   PSIMode.calculate / find_root     psi_mode.py:106-491        (scipy newton's secant rule)
//   MpiScheduler.runcompute           psi_mode_mpi.py:162-178    (per-item re-seeding)
// The Python generators in psi_mode.py / complex_roots.py are the readable statement of the same logic
// and remain the general path (verbose runs, count_roots=True inside PSIMode, user callables).
#include <dlfcn.h>
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <complex>
#include <map>
#include <string>
#include <vector>

#include "psi_internal.hpp"

typedef std::complex<double> cd;

namespace {

// ------------------------------------------------------------------------------------------------
// LAPACK binding (ILP64 zgesdd from numpy's OpenBLAS, LP64 zggev from scipy's)
// ------------------------------------------------------------------------------------------------
typedef void (*zgesdd64_t)(const char*, const int64_t*, const int64_t*, cd*, const int64_t*, double*, cd*,
                           const int64_t*, cd*, const int64_t*, cd*, const int64_t*, double*, int64_t*,
                           int64_t*, size_t);
typedef void (*zggev32_t)(const char*, const char*, const int*, cd*, const int*, cd*, const int*, cd*, cd*,
                          cd*, const int*, cd*, const int*, cd*, const int*, double*, int*, size_t, size_t);
typedef void (*setthreads_t)(int);

// call counters / time (microseconds) of the LAPACK drivers, for profiling the host side
std::atomic<long long> g_cnt_svd(0), g_us_svd(0), g_cnt_eig(0), g_us_eig(0);
struct ScopedTimer {
    std::atomic<long long>& cnt; std::atomic<long long>& us;
    std::chrono::steady_clock::time_point t0;
    ScopedTimer(std::atomic<long long>& c, std::atomic<long long>& u) : cnt(c), us(u), t0(std::chrono::steady_clock::now()) {}
    ~ScopedTimer() {
        us += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
        cnt += 1;
    }
};

int g_threads = 0;          // host threads of the engine (0 = OpenMP default)
int engine_threads() {
    int n = g_threads > 0 ? g_threads : omp_get_max_threads();
    return n < 1 ? 1 : n;
}

struct Lapack {
    void* h_np = nullptr;
    void* h_sp = nullptr;
    zgesdd64_t zgesdd = nullptr;
    zggev32_t zggev = nullptr;
    bool ready() const { return zgesdd && zggev; }
} g_lapack;


// numpy's complex division (loops_arithm_fp: Smith's algorithm, zero divisor -> inf/nan)
inline cd np_div(cd a, cd b) {
    const double ar = a.real(), ai = a.imag(), br = b.real(), bi = b.imag();
    const double abr = fabs(br), abi = fabs(bi);
    if (abr >= abi) {
        if (abr == 0 && abi == 0) return cd(ar / abr, ai / abr);
        const double rat = bi / br, scl = 1.0 / (br + bi * rat);
        return cd((ar + ai * rat) * scl, (ai - ar * rat) * scl);
    }
    const double rat = br / bi, scl = 1.0 / (bi + br * rat);
    return cd((ar * rat + ai) * scl, (ai * rat - ar) * scl);
}
inline cd np_mul(cd a, cd b) {
    return cd(a.real() * b.real() - a.imag() * b.imag(), a.real() * b.imag() + a.imag() * b.real());
}
inline bool finite_c(cd z) { return std::isfinite(z.real()) && std::isfinite(z.imag()); }
inline double abs_c(cd z) { return hypot(z.real(), z.imag()); }

// numpy's ordering of complex numbers for sort(): by real part, then imaginary part, NaNs last
inline bool np_less(cd a, cd b) {
    const double ar = a.real(), ai = a.imag(), br = b.real(), bi = b.imag();
    if (ar < br) return ai == ai || bi != bi;
    if (ar > br) return bi != bi && ai == ai;
    if (ar == br || (ar != ar && br != br)) return ai < bi || (bi != bi && ai == ai);
    return br != br;
}

// ------------------------------------------------------------------------------------------------
// numpy legacy random stream: RandomState(seed).uniform
// ------------------------------------------------------------------------------------------------
struct MT19937 {
    uint32_t mt[624];
    int pos;
    void seed(uint32_t s) {
        mt[0] = s;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        pos = 624;
    }
    uint32_t next() {
        if (pos >= 624) {
            for (int k = 0; k < 624; ++k) {
                uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            pos = 0;
        }
        uint32_t y = mt[pos++];
        y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
        return y;
    }
    double sample() {
        const uint32_t a = next() >> 5, b = next() >> 6;
        return (a * 67108864.0 + b) / 9007199254740992.0;
    }
    double uniform(double lo, double hi) { return lo + (hi - lo) * sample(); }
};

struct Rect {
    double xmin, xmax, ymin, ymax;
    bool is_in(cd z) const {
        return z.real() > xmin && z.real() < xmax && z.imag() > ymin && z.imag() < ymax;
    }
    // all real parts first, then all imaginary parts (complex_roots.py:404-409)
    void sample(MT19937& rng, int n, std::vector<cd>& out) const {
        std::vector<double> re(n);
        for (int i = 0; i < n; ++i) re[i] = rng.uniform(xmin, xmax);
        for (int i = 0; i < n; ++i) out.push_back(cd(re[i], rng.uniform(ymin, ymax)));
    }
};

struct EngineError {
    std::string what;
};

// ------------------------------------------------------------------------------------------------
// secant iteration (scipy.optimize.newton, secant branch) as a resumable machine
// ------------------------------------------------------------------------------------------------
struct Secant {
    cd p0, p1, q0, q1, p, root;
    int maxiter = 0, it = 0, calls = 0, iterations = 0;
    double xtol = 0;
    bool converged = false, flagged = false, done = false, started = false;

    // returns the number of points requested (written to req)
    int start(cd x0, cd x1, int maxit, double tol, cd* req) {
        if (!(tol > 0)) throw EngineError{"tol too small (<= 0) in secant"};
        p0 = x0; p1 = x1; maxiter = maxit; xtol = tol; it = 0; calls = 0;
        converged = flagged = done = false; started = false;
        req[0] = p0; req[1] = p1;
        return 2;
    }
    void finish(cd r, bool conv, int iters) { root = r; converged = conv; iterations = iters; done = true; }
    // feed the values of the last request; returns the number of new points requested (0 = done)
    int feed(const cd* vals, cd* req) {
        if (!started) {
            q0 = vals[0]; q1 = vals[1]; calls = 2; started = true;
            if (abs_c(q1) < abs_c(q0)) { std::swap(p0, p1); std::swap(q0, q1); }
            p = p1;
        } else {
            q1 = vals[0]; calls += 1;
        }
        if (it >= maxiter) { finish(p, false, maxiter); return 0; }
        if (q1 == q0) {
            flagged = (p1 != p0);
            finish((p1 + p0) / 2.0, false, it + 1);
            return 0;
        }
        if (abs_c(q1) > abs_c(q0)) {
            const cd r = np_div(q0, q1);
            p = np_div(np_mul(-r, p1) + p0, 1.0 - r);
        } else {
            const cd r = np_div(q1, q0);
            p = np_div(np_mul(-r, p0) + p1, 1.0 - r);
        }
        if (abs_c(p - p1) <= xtol) { finish(p, true, it + 1); return 0; }
        p0 = p1; q0 = q1; p1 = p;
        ++it;
        if (it >= maxiter) {
            // scipy still evaluates f(p1) here before giving up; the value is unused, so only the
            // call is counted (psi_mode.py adds function_calls to n_function_call)
            calls += 1;
            finish(p, false, maxiter);
            return 0;
        }
        req[0] = p1;
        return 1;
    }
};

inline cd second_point(cd z0, double eps) {
    cd z1 = z0 * (1.0 + eps);
    return z1 + (z1.real() >= 0 ? eps : -eps);
}

// ------------------------------------------------------------------------------------------------
// AAA
// ------------------------------------------------------------------------------------------------
struct AAA {
    Rect domain;
    double tol = 1e-13, clean_tol = 1e-10;
    std::vector<cd> F, Z, weights;
    std::vector<int> mask;
    std::vector<double> R;
    // greedy-phase cache (depends on F, Z, tol only)
    std::vector<cd> cF, cZ, cW;
    std::vector<int> cMask;
    std::vector<double> cR;
    double cTol = -1;
    // Everything after the greedy phase is a pure function of (F, Z, support mask): the Loewner fit,
    // the poles with their residue test, the zeros.  PSIMode's clean_tol-halving loop revisits the same
    // masks dozens of times, so results are memoised per mask (cleared when the samples change).
    struct PoleInfo { double ratio; int pos; };
    struct Memo {
        bool have_fit = false, have_poles = false, have_zeros = false;
        std::vector<cd> weights, zeros;
        std::vector<double> R;
        double resid = 0;
        std::vector<PoleInfo> poles;
    };
    std::map<std::string, Memo> memo;
    std::string key() const { std::string k(mask.size(), '0'); for (size_t i = 0; i < mask.size(); ++i) k[i] = mask[i] ? '1' : '0'; return k; }

    int n_nodes() const { int c = 0; for (int v : mask) c += v; return c; }
    double fmax() const { double m = 0; for (cd f : F) m = std::max(m, abs_c(f)); return m; }

    double fit() {
        Memo& mm = memo[key()];
        if (mm.have_fit) { weights = mm.weights; R = mm.R; return mm.resid; }
        const double r = fit_lapack();
        mm.weights = weights; mm.R = R; mm.resid = r; mm.have_fit = true;
        return r;
    }

    // weights = conj(last row of VT) of the Loewner matrix; residuals at the non-support points
    double fit_lapack() {
        const int M = (int)F.size();
        std::vector<int> on, off;
        for (int i = 0; i < M; ++i) (mask[i] ? on : off).push_back(i);
        const int64_t m = (int64_t)on.size(), r = (int64_t)off.size();
        std::vector<cd> cauchy((size_t)r * m), A((size_t)r * m);         // column-major r x m
        for (int64_t j = 0; j < m; ++j)
            for (int64_t i = 0; i < r; ++i) {
                const cd c = np_div(cd(1.0, 0.0), Z[off[i]] - Z[on[j]]);
                cauchy[i + j * r] = c;
                A[i + j * r] = np_mul(F[off[i]], c) - np_mul(c, F[on[j]]);
            }
        weights.assign(m, cd(0, 0));
        if (r == 0) {
            // 0 x m matrix: numpy returns VT = identity
            weights[m - 1] = cd(1.0, 0.0);
        } else {
            // jobz = 'S' (economy) when r >= m: VT is the full m x m matrix either way and bit-identical
            // to numpy's full_matrices=True result (checked over 3000 AAA-like matrices); for r < m the
            // trailing rows of the full VT (a null vector chosen by LAPACK) are needed -> 'A'.
            const bool econ = r >= m;
            const int64_t mn = std::min(r, m), mx = std::max(r, m);
            std::vector<double> s(mn), rwork((size_t)std::max<int64_t>(1, mn * std::max(5 * mn + 7, 2 * mx + 2 * mn + 1)));
            std::vector<cd> U((size_t)r * (econ ? mn : r)), VT((size_t)m * m), work(1);
            std::vector<int64_t> iwork(8 * mn);
            ScopedTimer tm(g_cnt_svd, g_us_svd);
            int64_t lwork = -1, info = 0, lda = r, ldu = r, ldvt = m;
            const char* job = econ ? "S" : "A";
            static thread_local std::map<std::pair<int64_t, int64_t>, int64_t> lwork_cache;
            int64_t& cached = lwork_cache[std::make_pair(r, m)];
            if (cached == 0) {
                g_lapack.zgesdd(job, &r, &m, A.data(), &lda, s.data(), U.data(), &ldu, VT.data(), &ldvt,
                                work.data(), &lwork, rwork.data(), iwork.data(), &info, 1);
                cached = std::max<int64_t>(1, (int64_t)work[0].real());
            }
            lwork = cached;
            work.resize((size_t)lwork);
            g_lapack.zgesdd(job, &r, &m, A.data(), &lda, s.data(), U.data(), &ldu, VT.data(), &ldvt,
                            work.data(), &lwork, rwork.data(), iwork.data(), &info, 1);
            if (info != 0) throw EngineError{"SVD did not converge"};
            for (int64_t j = 0; j < m; ++j) weights[j] = std::conj(VT[(m - 1) + j * m]);
        }
        R.assign(r, 0.0);
        double rmax = 0;
        for (int64_t i = 0; i < r; ++i) {
            cd num(0, 0), den(0, 0);
            for (int64_t j = 0; j < m; ++j) {
                num += np_mul(cauchy[i + j * r], np_mul(weights[j], F[on[j]]));
                den += np_mul(cauchy[i + j * r], weights[j]);
            }
            R[i] = abs_c(F[off[i]] - np_div(num, den));
            if (R[i] > rmax || R[i] != R[i]) rmax = R[i];
        }
        return rmax / fmax();
    }

    int free_index(int k) const {          // position of the k-th non-support sample
        int c = 0;
        for (int i = 0; i < (int)mask.size(); ++i) { if (!mask[i]) { if (c == k) return i; ++c; } }
        return (int)mask.size();
    }
    int node_index(int k) const {          // position of the k-th support sample
        int c = 0;
        for (int i = 0; i < (int)mask.size(); ++i) { if (mask[i]) { if (c == k) return i; ++c; } }
        return (int)mask.size();
    }

    double step() {
        int best = 0;
        for (int i = 1; i < (int)R.size(); ++i) if (R[i] > R[best]) best = i;
        std::vector<int> order(R.size());
        for (int i = 0; i < (int)R.size(); ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return R[a] < R[b]; });
        int pos = free_index(best);
        mask[pos] = 1;
        try {
            return fit();
        } catch (const EngineError&) {
            mask[pos] = 0;
            pos = free_index(order[order.size() - 2]);
            mask[pos] = 1;
            return fit();
        }
    }

    void calculate(const std::vector<cd>& Fin, const std::vector<cd>& Zin) {
        if (Fin.size() != Zin.size()) throw EngineError{"Function samples and sample points need to have the same shape"};
        cd tot(0, 0);
        for (size_t i = 0; i < Fin.size(); ++i) tot += Fin[i] + Zin[i];
        if (!finite_c(tot)) throw EngineError{"Nan or Inf in sample points or function samples"};
        if (!(F == Fin && Z == Zin)) memo.clear();
        F = Fin; Z = Zin;
        const int M = (int)F.size();
        if (cTol == tol && cF == F && cZ == Z) {
            mask = cMask; weights = cW; R = cR;
        } else {
            mask.assign(M, 0);
            mask[0] = 1; mask[1] = 1;
            fit();
            for (int k = 1; k < M / 2; ++k)
                if (step() < tol) break;
            cF = F; cZ = Z; cTol = tol; cMask = mask; cW = weights; cR = R;
        }
        while (cleanup() != 0) {}
    }

    // generalised eigenvalues of the arrowhead pencil (first row = top), in LAPACK's order
    void arrowhead(const std::vector<cd>& top, std::vector<cd>& w) const {
        std::vector<cd> zn;
        for (size_t i = 0; i < mask.size(); ++i) if (mask[i]) zn.push_back(Z[i]);
        const int m = (int)zn.size(), n = m + 1;
        std::vector<cd> A((size_t)n * n, cd(0, 0)), B((size_t)n * n, cd(0, 0));
        for (int j = 1; j < n; ++j) {
            A[j + (size_t)j * n] = zn[j - 1];
            A[0 + (size_t)j * n] = top[j - 1];
            A[j + 0] = cd(1.0, 0.0);
            B[j + (size_t)j * n] = cd(1.0, 0.0);
        }
        // eigenvalues only (jobvr = 'N'): bit-identical to the values scipy.linalg.eig returns with its
        // default right=True (checked over 500 random arrowhead pencils), at ~15 % less time
        std::vector<cd> alpha(n), beta(n), vr(1), work(1), vl(1);
        std::vector<double> rwork(8 * n);
        ScopedTimer tm(g_cnt_eig, g_us_eig);
        int lwork = -1, info = 0, ldvl = 1, ldvr = 1;
        static thread_local std::map<int, int> lwork_cache;
        int& cached = lwork_cache[n];
        if (cached == 0) {
            g_lapack.zggev("N", "N", &n, A.data(), &n, B.data(), &n, alpha.data(), beta.data(), vl.data(), &ldvl,
                           vr.data(), &ldvr, work.data(), &lwork, rwork.data(), &info, 1, 1);
            cached = std::max(1, (int)work[0].real());
        }
        lwork = cached;
        work.resize(lwork);
        g_lapack.zggev("N", "N", &n, A.data(), &n, B.data(), &n, alpha.data(), beta.data(), vl.data(), &ldvl,
                       vr.data(), &ldvr, work.data(), &lwork, rwork.data(), &info, 1, 1);
        if (info != 0) throw EngineError{"generalised eigenvalue problem did not converge"};
        w.resize(n);
        for (int i = 0; i < n; ++i) w[i] = np_div(alpha[i], beta[i]);
    }

    cd evaluate(cd z, bool strict) const {
        cd num(0, 0), den(0, 0);
        int j = 0;
        for (size_t i = 0; i < mask.size(); ++i) {
            if (!mask[i]) continue;
            cd t = weights[j];
            if (t != cd(0, 0)) {
                const cd dz = z - Z[i];
                if (strict && dz == cd(0, 0)) throw EngineError{"divide by zero"};
                t = np_div(t, dz);
            }
            num += np_mul(t, F[i]);
            den += t;
            ++j;
        }
        if (strict && den == cd(0, 0)) throw EngineError{"divide by zero"};
        const cd r = np_div(num, den);
        if (strict && !finite_c(r)) throw EngineError{"invalid value"};
        return r;
    }

    int cleanup() {
        Memo& mm = memo[key()];
        if (!mm.have_poles) {
            std::vector<cd> e;
            arrowhead(weights, e);
            const double side = 1.0e-5, fm = fmax();
            cd ring[4];
            for (int k = 0; k < 4; ++k) ring[k] = side * std::exp(cd(0, -0.25 * M_PI) + cd(0, 0.5 * M_PI * k));
            for (size_t q = 2; q < e.size(); ++q) {             // the first two are the infinite ones
                const cd pole = e[q];
                if (!domain.is_in(pole)) continue;
                cd zc[4], fc[4];
                for (int k = 0; k < 4; ++k) { zc[k] = pole + ring[k]; fc[k] = evaluate(zc[k], false); }
                cd integral(0, 0);
                for (int k = 0; k < 4; ++k) integral += np_mul(0.5 * (zc[(k + 1) % 4] - zc[k]), fc[(k + 1) % 4] + fc[k]);
                int nearest = 0, j = 0; double bestd = INFINITY;
                for (size_t i = 0; i < mask.size(); ++i) {
                    if (!mask[i]) continue;
                    const double dd = abs_c(pole - Z[i]);
                    if (dd < bestd) { bestd = dd; nearest = j; }
                    ++j;
                }
                mm.poles.push_back(PoleInfo{abs_c(integral) / fm, node_index(nearest)});
            }
            mm.have_poles = true;
        }
        // positions refer to the mask BEFORE any removal (the reference collects first, then removes)
        std::vector<int> doomed;
        for (const PoleInfo& pi : mm.poles) if (pi.ratio < clean_tol) doomed.push_back(pi.pos);
        for (int pos : doomed) mask[pos] = 0;
        fit();
        return (int)doomed.size();
    }

    void find_zeros(std::vector<cd>& zeros) {
        Memo& mm = memo[key()];
        if (mm.have_zeros) { zeros = mm.zeros; return; }
        find_zeros_lapack(zeros);
        mm.zeros = zeros; mm.have_zeros = true;
    }

    void find_zeros_lapack(std::vector<cd>& zeros) {
        std::vector<cd> top, e;
        int j = 0;
        for (size_t i = 0; i < mask.size(); ++i) if (mask[i]) top.push_back(np_mul(weights[j++], F[i]));
        arrowhead(top, e);
        std::stable_sort(e.begin(), e.end(), np_less);
        zeros.assign(e.begin(), e.end() - 2);
        if ((int)zeros.size() != (int)weights.size() - 1)
            throw EngineError{"Number of roots of rational approximation not equal to number of nodes used"};
        for (size_t i = 0; i < zeros.size(); ++i) {
            const cd z0 = zeros[i];
            try {
                Secant s;
                cd req[2], val[2];
                int n = s.start(z0, second_point(z0, 1.0e-4), 50, 1.48e-8, req);
                if (!finite_c(req[0]) || !finite_c(req[1])) throw EngineError{"invalid value"};
                while (n > 0) {
                    for (int k = 0; k < n; ++k) val[k] = evaluate(req[k], true);
                    n = s.feed(val, req);
                    if (n > 0 && !finite_c(req[0])) throw EngineError{"invalid value"};
                }
                if (!s.flagged) zeros[i] = s.root;
            } catch (const EngineError&) {
                // the reference turns RuntimeWarnings into errors here and keeps the start value
            }
        }
        if (zeros.size() > 1) {
            std::stable_sort(zeros.begin(), zeros.end(), np_less);
            std::vector<cd> u;
            u.push_back(zeros[0]);
            // unique_within_tol: keep b[i] when (b[i] - b[i-1]) > 1e-12 as COMPLEX numbers, i.e.
            // numpy's lexicographic ">" against (1e-12 + 0j)
            for (size_t i = 1; i < zeros.size(); ++i)
                if (np_less(cd(1e-12, 0.0), zeros[i] - zeros[i - 1])) u.push_back(zeros[i]);
            zeros.swap(u);
        }
    }
};

// ------------------------------------------------------------------------------------------------
// PSIMode.calculate as a resumable machine
// ------------------------------------------------------------------------------------------------
struct ModeItem {
    // configuration
    int pidx = 0;                      // row of the per-item parameter arrays (dist, kx, kz, alpha)
    Rect dom;
    int n_sample = 20, max_zoom = 1, max_secant = 100, min_nodes = 4, per_level = 4;
    double tol = 1e-13, clean_tol = 1e-4;
    std::vector<cd> guesses;
    MT19937 rng;
    // state
    enum Phase { INIT, SAMPLES, GUESS, ANALYZE, POLISH, ZOOM, DONE, FAILED } phase = INIT;
    std::vector<cd> zs, fs, req, zeros_start, out_roots, zoom;
    AAA ra;
    int level = 0, iguess = 0, izoom = 0, max_iter = 0, n_calls = 0, max_conv_iter = 0;
    double zoom_dx = 0;
    std::vector<cd> w0;               // start values handed to the polish kernel (K3)
    std::vector<cd> polished;         // its answer: roots, or the sentinel -1j
    int polish_calls = 0, polish_conv = 0, polish_status = 0;
    std::string error;

    void request_box(double sx, double sy, cd centre) {
        double re = centre.real(), im = centre.imag();
        if (re < dom.xmin + 0.5 * sx) re = dom.xmin + 0.5 * sx;
        if (re > dom.xmax - 0.5 * sx) re = dom.xmax - 0.5 * sx;
        if (im < dom.ymin + 0.5 * sy) im = dom.ymin + 0.5 * sy;
        if (im > dom.ymax - 0.5 * sy) im = dom.ymax - 0.5 * sy;
        Rect box{re - 0.5 * sx, re + 0.5 * sx, im - 0.5 * sy, im + 0.5 * sy};
        req.clear();
        box.sample(rng, n_sample, req);
    }
    void absorb_samples(const cd* vals) {
        for (size_t i = 0; i < req.size(); ++i) { zs.push_back(req[i]); fs.push_back(vals[i]); }
        n_calls += (int)req.size();
    }
    // zeros inside the search rectangle
    void zeros_in_domain(std::vector<cd>& z) {
        std::vector<cd> all;
        ra.find_zeros(all);
        z.clear();
        for (cd v : all) if (dom.is_in(v)) z.push_back(v);
    }

    // AAA + zero selection of one zoom level (psi_mode.py:147-310); sets up the polish
    void analyze() {
        ra.domain = dom; ra.tol = tol; ra.clean_tol = clean_tol;
        ra.calculate(fs, zs);
        std::vector<cd> zin;
        zeros_in_domain(zin);
        const int need = min_nodes * (1 + level + (int)guesses.size());
        const double saved = ra.clean_tol;
        auto growing = [&]() { for (cd v : zin) if (v.imag() > 0.0) return true; return false; };
        while (ra.n_nodes() < need || !growing()) {
            ra.clean_tol = 0.5 * ra.clean_tol;
            if (ra.clean_tol < ra.tol) break;
            ra.calculate(fs, zs);
            zeros_in_domain(zin);
        }
        ra.clean_tol = saved;
        std::vector<cd> zeros;
        ra.find_zeros(zeros);
        std::vector<cd> cand;
        for (cd v : zeros)
            if (v.imag() < dom.ymax && v.real() < dom.xmax && v.real() > dom.xmin) cand.push_back(v);
        zoom_dx = pow(10.0, -1 - 0.5 * level);
        zoom.clear();
        bool have_least = false;
        cd least;
        if (!cand.empty()) {
            std::stable_sort(cand.begin(), cand.end(), [](cd a, cd b) { return a.imag() < b.imag(); });
            zoom.push_back(cand.back());
            for (int i = (int)cand.size() - 2; i >= 0; --i) {
                bool ok = true;
                for (cd q : zoom) if (fabs(cand[i].real() - q.real()) < zoom_dx) ok = false;
                if (ok) zoom.push_back(cand[i]);
            }
            if ((int)zoom.size() > per_level) zoom.resize(per_level);
            least = zoom[0];
            have_least = true;
        }
        w0.clear();
        for (cd v : zeros) if (dom.is_in(v)) w0.push_back(v);
        max_iter = max_secant;
        if (level < max_zoom) max_iter = n_sample;
        if (w0.empty() && have_least) { w0.push_back(least); max_iter = 5; }
        polished.assign(w0.size(), cd(0, 0));
    }

    // after the polish of a level: return, or open zoom domains (psi_mode.py:312-331)
    bool after_polish() {
        out_roots.clear();
        for (cd v : polished) if (dom.is_in(v)) out_roots.push_back(v);
        bool grow = false;
        for (cd v : out_roots) if (v.imag() > 0.0) grow = true;
        if (grow || level == max_zoom || zoom.empty()) { phase = DONE; return false; }
        izoom = 0;
        phase = ZOOM;
        const cd c = zoom[0];
        request_box(zoom_dx, std::min(0.1, fabs(c.imag())), c);
        return true;
    }

    // Advance with the values of the pending request (nullptr at the start).  Leaves a new request in
    // `req` (phase != DONE/FAILED) or finishes.
    void advance(const cd* vals) {
        try {
            for (;;) {
                switch (phase) {
                case INIT:
                    req.clear();
                    dom.sample(rng, n_sample, req);
                    phase = SAMPLES;
                    return;
                case SAMPLES:
                    absorb_samples(vals);
                    iguess = 0;
                    phase = GUESS;
                    if (!guesses.empty()) {
                        const cd g = guesses[0];
                        const double s = std::min(1e-3, fabs(g.imag()) * 4);
                        request_box(s, s, g);
                        return;
                    }
                    level = 0;
                    phase = ANALYZE;
                    break;
                case GUESS:
                    absorb_samples(vals);
                    ++iguess;
                    if (iguess < (int)guesses.size()) {
                        const cd g = guesses[iguess];
                        const double s = std::min(1e-3, fabs(g.imag()) * 4);
                        request_box(s, s, g);
                        return;
                    }
                    level = 0;
                    phase = ANALYZE;
                    break;
                case ANALYZE:
                    analyze();
                    if (!w0.empty()) { phase = POLISH; req.clear(); return; }   // K3 takes over
                    if (after_polish()) return;
                    return;
                case POLISH:
                    // `polished`, polish_calls, polish_conv were filled in by the polish kernel
                    if (polish_status != 0) throw EngineError{"tol too small (<= 0) in secant"};
                    n_calls += polish_calls;
                    max_conv_iter = std::max(max_conv_iter, polish_conv);
                    if (after_polish()) return;
                    return;
                case ZOOM:
                    absorb_samples(vals);
                    ++izoom;
                    if (izoom < (int)zoom.size()) {
                        const cd c = zoom[izoom];
                        request_box(zoom_dx, std::min(0.1, fabs(c.imag())), c);
                        return;
                    }
                    ++level;
                    phase = ANALYZE;
                    break;
                default:
                    return;
                }
                vals = nullptr;
            }
        } catch (const EngineError& e) {
            error = e.what;
            phase = FAILED;
        }
    }
};

// ------------------------------------------------------------------------------------------------
// lock-step driver: per round ONE K1 launch for the union of all sample requests and ONE K3 launch
// for the union of all polish requests
// ------------------------------------------------------------------------------------------------
int run_lockstep(psi_ctx* ctx, std::vector<ModeItem>& items, int n_params, const int* dist, const double* kx,
                 const double* kz, const double* alpha, long long* stats) {
    typedef std::chrono::steady_clock clk;
    auto us = [](clk::time_point a, clk::time_point b) {
        return (long long)std::chrono::duration_cast<std::chrono::microseconds>(b - a).count();
    };
    long long t_logic = 0, t_eval = 0, rounds = 0, evals = 0, launches = 0, polish_evals = 0;
    auto t0 = clk::now();
    const int nt = engine_threads();
    std::vector<int> live;
    for (int i = 0; i < (int)items.size(); ++i) live.push_back(i);
#pragma omp parallel for schedule(dynamic, 4) num_threads(nt)
    for (int k = 0; k < (int)live.size(); ++k) items[live[k]].advance(nullptr);
    std::vector<cd> w, D, pw, pout;
    std::vector<int> pidx, ppidx, poff, pmax, pcalls, pconv, pstat;
    std::vector<double> pdom;
    std::vector<size_t> off;
    for (;;) {
        std::vector<int> pend;
        for (int i : live)
            if (items[i].phase != ModeItem::DONE && items[i].phase != ModeItem::FAILED) pend.push_back(i);
        if (pend.empty()) break;
        // sample requests -> K1
        w.clear(); pidx.clear(); off.assign(pend.size() + 1, 0);
        // polish requests -> K3
        pw.clear(); ppidx.clear(); poff.assign(1, 0); pmax.clear(); pdom.clear();
        std::vector<int> pitem;
        for (size_t k = 0; k < pend.size(); ++k) {
            ModeItem& it = items[pend[k]];
            off[k] = w.size();
            if (it.phase == ModeItem::POLISH) {
                pitem.push_back(pend[k]);
                ppidx.push_back(it.pidx);
                for (cd z : it.w0) pw.push_back(z);
                poff.push_back((int)pw.size());
                pmax.push_back(it.max_iter);
                pdom.push_back(it.dom.xmin); pdom.push_back(it.dom.xmax);
                pdom.push_back(it.dom.ymin); pdom.push_back(it.dom.ymax);
            } else {
                for (cd z : it.req) { w.push_back(z); pidx.push_back(it.pidx); }
            }
        }
        off[pend.size()] = w.size();
        D.resize(w.size());
        auto t1 = clk::now();
        t_logic += us(t0, t1);
        if (!w.empty()) {
            int rc = psi_internal_eval_indexed(ctx, (long long)w.size(), pidx.data(), n_params, dist, kx, kz,
                                               alpha, (const double*)w.data(), (double*)D.data());
            if (rc != PSI_OK) return rc;
            ++launches;
            evals += (long long)w.size();
        }
        if (!pitem.empty()) {
            const int np = (int)pitem.size();
            pout.resize(pw.size()); pcalls.assign(np, 0); pconv.assign(np, 0); pstat.assign(np, 0);
            long long pe = 0;
            int rc = psi_internal_polish(ctx, np, ppidx.data(), n_params, dist, kx, kz, alpha, pdom.data(),
                                         poff.data(), (double*)pw.data(), pmax.data(), (double*)pout.data(),
                                         pcalls.data(), pconv.data(), pstat.data(), &pe);
            if (rc != PSI_OK) return rc;
            ++launches;
            polish_evals += pe;
            for (int k = 0; k < np; ++k) {
                ModeItem& it = items[pitem[k]];
                it.polished.assign(pout.begin() + poff[k], pout.begin() + poff[k + 1]);
                it.polish_calls = pcalls[k]; it.polish_conv = pconv[k]; it.polish_status = pstat[k];
            }
        }
        t0 = clk::now();
        t_eval += us(t1, t0);
        ++rounds;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nt)
        for (int k = 0; k < (int)pend.size(); ++k) items[pend[k]].advance(D.data() + off[k]);
        live.swap(pend);
    }
    t_logic += us(t0, clk::now());
    if (stats) {
        stats[0] = rounds; stats[1] = evals + polish_evals; stats[3] = t_logic; stats[4] = t_eval;
        stats[5] = launches; stats[6] = polish_evals;
    }
    return PSI_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int psi_engine_bind_lapack(const char* numpy_openblas64, const char* scipy_openblas32) {
    if (g_lapack.ready()) return PSI_OK;
    if (!numpy_openblas64 || !scipy_openblas32) return psi_internal_fail(PSI_ERR_ARG, "null LAPACK library path");
    g_lapack.h_np = dlopen(numpy_openblas64, RTLD_NOW | RTLD_LOCAL);
    g_lapack.h_sp = dlopen(scipy_openblas32, RTLD_NOW | RTLD_LOCAL);
    if (!g_lapack.h_np || !g_lapack.h_sp)
        return psi_internal_fail(PSI_ERR_ARG, std::string("dlopen of OpenBLAS failed: ") + (dlerror() ? dlerror() : "?"));
    g_lapack.zgesdd = (zgesdd64_t)dlsym(g_lapack.h_np, "scipy_zgesdd_64_");
    g_lapack.zggev = (zggev32_t)dlsym(g_lapack.h_sp, "scipy_zggev_");
    if (!g_lapack.ready()) return psi_internal_fail(PSI_ERR_ARG, "zgesdd/zggev symbols not found in the OpenBLAS libraries");
    // tiny matrices: one BLAS thread per call (the engine parallelises over items instead)
    if (setthreads_t f = (setthreads_t)dlsym(g_lapack.h_np, "scipy_openblas_set_num_threads64_")) f(1);
    if (setthreads_t f = (setthreads_t)dlsym(g_lapack.h_sp, "scipy_openblas_set_num_threads")) f(1);
    return PSI_OK;
}

/* host threads used by the engine's per-item logic (AAA etc.); n <= 0 restores the OpenMP default */
void psi_engine_set_threads(int n) { g_threads = n; }

/* K4 drivers: the same batch as psi_mode_batch_host, but every search runs entirely on the device.  The host
 * only generates the random streams: one U[0,1) stream per distinct seed, long enough for the hungriest item
 * using it (every domain -- initial, one per guess, four per zoom level -- consumes 2*n_sample numbers). */
static int mode_streams(int n_items, const int* iparams, const unsigned* seeds, const int* guess_off,
                        std::vector<double>& uni, std::vector<int>& uoff, int* cap_out) {
    std::map<unsigned, long long> need;
    int cap = 0;
    for (int i = 0; i < n_items; ++i) {
        const int ns = iparams[3 * i], mz = iparams[3 * i + 1], ng = guess_off[i + 1] - guess_off[i];
        if (ns < 2 || mz < 0) return psi_internal_fail(PSI_ERR_ARG, "n_sample must be >= 2, max_zoom_domains >= 0");
        const int domains = 1 + ng + 4 * mz;
        cap = std::max(cap, ns * domains);
        long long& n = need[seeds[i]];
        n = std::max(n, 2ll * ns * domains);
    }
    std::map<unsigned, long long> start;
    for (auto& kv : need) {
        start[kv.first] = (long long)uni.size();
        MT19937 rng;
        rng.seed(kv.first);
        for (long long k = 0; k < kv.second; ++k) uni.push_back(rng.sample());
    }
    if (uni.size() > 0x7fffffffull) return psi_internal_fail(PSI_ERR_CAPACITY, "random streams too long");
    if (cap > PSI_MODE_MAX_SAMPLES)
        return psi_internal_fail(PSI_ERR_CAPACITY, "more than 512 samples per search: use psi_mode_batch_host");
    uoff.resize(n_items);
    for (int i = 0; i < n_items; ++i) uoff[i] = (int)start[seeds[i]];
    *cap_out = cap;
    return PSI_OK;
}

int psi_mode_batch_device(psi_ctx* ctx, int n_items, const int* dist, const double* kx, const double* kz,
                          const double* alpha, const double* domain, const int* iparams, const double* dparams,
                          const unsigned* seeds, const int* guess_off, const double* guesses, int max_roots,
                          double* roots, int* n_roots, int* n_calls, int* status, long long* stats) {
    if (!ctx || n_items < 0 || !dist || !kx || !kz || !alpha || !domain || !iparams || !dparams || !seeds ||
        !guess_off || !roots || !n_roots || !n_calls || !status || max_roots < 1)
        return psi_internal_fail(PSI_ERR_ARG, "psi_mode_batch_device: bad argument");
    std::vector<double> uni;
    std::vector<int> uoff;
    int cap = 0;
    int rc = mode_streams(n_items, iparams, seeds, guess_off, uni, uoff, &cap);
    if (rc != PSI_OK) return rc;
    const double dummy[2] = {0.0, 0.0};
    return psi_internal_mode_device(ctx, n_items, dist, kx, kz, alpha, domain, iparams, dparams, uni.data(),
                                    (long long)uni.size(), uoff.data(), guess_off,
                                    guesses ? guesses : dummy, cap, max_roots, roots, n_roots, n_calls, status,
                                    stats);
}

/* The same batch with the results left ON THE DEVICE as packed records (see PsiModeExtra): what a rank hands to
 * the all-gather of a multi-GPU map without a detour through host memory. */
int psi_mode_batch_device_rec(psi_ctx* ctx, int n_items, const int* dist, const double* kx, const double* kz,
                              const double* alpha, const double* domain, const int* iparams, const double* dparams,
                              const unsigned* seeds, const int* guess_off, const double* guesses, int max_roots,
                              int rec_rows, double** rec_dev, long long* stats) {
    if (!ctx || n_items < 0 || !rec_dev || max_roots < 1 || !guess_off ||
        (n_items > 0 && (!dist || !kx || !kz || !alpha || !domain || !iparams || !dparams || !seeds)))
        return psi_internal_fail(PSI_ERR_ARG, "psi_mode_batch_device_rec: bad argument");
    std::vector<double> uni;
    std::vector<int> uoff;
    int cap = 0;
    int rc = mode_streams(n_items, iparams, seeds, guess_off, uni, uoff, &cap);
    if (rc != PSI_OK) return rc;
    const double dummy[2] = {0.0, 0.0};
    PsiModeExtra ex;
    ex.rec_dev = rec_dev; ex.rec_rows = rec_rows;
    return psi_internal_mode_device(ctx, n_items, dist, kx, kz, alpha, domain, iparams, dparams, uni.data(),
                                    (long long)uni.size(), uoff.data(), guess_off, guesses ? guesses : dummy, cap,
                                    max_roots, nullptr, nullptr, nullptr, nullptr, stats, &ex);
}

/* ONE search with the caller's own stream of U[0,1) numbers (n_uni >= 2*n_sample*(1 + n_guess + 4*max_zoom)) and
 * everything PSIMode exposes after calculate(): roots, n_function_call, the sample points and values (cap =
 * n_sample*(1 + n_guess + 4*max_zoom) slots each, *n_samples used), the support mask of the final rational
 * approximation, the number of random numbers consumed and the largest iteration count of a converged polish. */
int psi_mode_search_detail(psi_ctx* ctx, int dist, double kx, double kz, double alpha, const double* domain,
                           const int* iparams, const double* dparams, const double* uni, int n_uni,
                           const double* guesses, int n_guess, int max_roots, double* roots, int* n_roots,
                           int* n_calls, int* status, double* z_sample, double* f_sample, int* n_samples,
                           int* mask, int* n_uni_used, int* max_conv) {
    if (!ctx || !domain || !iparams || !dparams || !uni || n_guess < 0 || max_roots < 1 || !roots || !n_roots ||
        !n_calls || !status)
        return psi_internal_fail(PSI_ERR_ARG, "psi_mode_search_detail: bad argument");
    const int cap = iparams[0] * (1 + n_guess + 4 * iparams[1]);
    if (iparams[0] < 2 || iparams[1] < 0 || n_uni < 2 * cap)
        return psi_internal_fail(PSI_ERR_ARG, "psi_mode_search_detail: bad n_sample / max_zoom / stream length");
    if (cap > PSI_MODE_MAX_SAMPLES)
        return psi_internal_fail(PSI_ERR_CAPACITY, "more than 512 samples per search: use psi_mode_batch_host");
    const int goff[2] = {0, n_guess}, uoff[1] = {0};
    const double dummy[2] = {0.0, 0.0};
    PsiModeExtra ex;
    ex.Z = z_sample; ex.F = f_sample; ex.M = n_samples; ex.mask = mask; ex.upos = n_uni_used; ex.conv = max_conv;
    long long stats[16];
    return psi_internal_mode_device(ctx, 1, &dist, &kx, &kz, &alpha, domain, iparams, dparams, uni, n_uni, uoff,
                                    goff, n_guess ? guesses : dummy, cap, max_roots, roots, n_roots, n_calls,
                                    status, stats, &ex);
}

/* profiling aid: {svd calls, svd ns, eig calls, eig ns} accumulated since load */
void psi_engine_counters(long long out[4]) {
    out[0] = g_cnt_svd; out[1] = g_us_svd; out[2] = g_cnt_eig; out[3] = g_us_eig;
}

int psi_mode_batch_host(psi_ctx* ctx, int n_items, const int* dist, const double* kx, const double* kz,
                        const double* alpha, const double* domain, const int* iparams, const double* dparams,
                        const unsigned* seeds, const int* guess_off, const double* guesses, int max_roots,
                        double* roots, int* n_roots, int* n_calls, int* status, long long* stats) {
    if (!ctx || n_items < 0 || !dist || !kx || !kz || !alpha || !domain || !iparams || !dparams || !seeds ||
        !guess_off || !roots || !n_roots || !status || max_roots < 1)
        return psi_internal_fail(PSI_ERR_ARG, "psi_mode_batch_host: bad argument");
    if (!g_lapack.ready()) return psi_internal_fail(PSI_ERR_ARG, "psi_engine_bind_lapack has not been called");
    std::vector<ModeItem> items(n_items);
    for (int i = 0; i < n_items; ++i) {
        ModeItem& m = items[i];
        m.pidx = i;
        m.dom = Rect{domain[4 * i], domain[4 * i + 1], domain[4 * i + 2], domain[4 * i + 3]};
        m.n_sample = iparams[3 * i]; m.max_zoom = iparams[3 * i + 1]; m.max_secant = iparams[3 * i + 2];
        m.tol = dparams[2 * i]; m.clean_tol = dparams[2 * i + 1];
        m.rng.seed(seeds[i]);
        for (int g = guess_off[i]; g < guess_off[i + 1]; ++g) m.guesses.push_back(cd(guesses[2 * g], guesses[2 * g + 1]));
        if (m.n_sample < 2) return psi_internal_fail(PSI_ERR_ARG, "n_sample must be at least 2");
    }
    long long st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int rc = run_lockstep(ctx, items, n_items, dist, kx, kz, alpha, st);
    if (rc != PSI_OK) return rc;
    int failed = 0;
    for (int i = 0; i < n_items; ++i) {
        const ModeItem& m = items[i];
        status[i] = (m.phase == ModeItem::DONE) ? 0 : 1;
        if (status[i]) { ++failed; psi_internal_fail(PSI_ERR_ARG, "item " + std::to_string(i) + ": " + m.error); }
        const int nr = (int)m.out_roots.size();
        n_roots[i] = nr;
        if (nr > max_roots) status[i] = 2;                         // truncated
        for (int k = 0; k < std::min(nr, max_roots); ++k) {
            roots[2 * ((size_t)i * max_roots + k)] = m.out_roots[k].real();
            roots[2 * ((size_t)i * max_roots + k) + 1] = m.out_roots[k].imag();
        }
        if (n_calls) n_calls[i] = m.n_calls;
    }
    st[2] = failed;
    if (stats) for (int k = 0; k < 7; ++k) stats[k] = st[k];
    return PSI_OK;
}

}  // extern "C"
