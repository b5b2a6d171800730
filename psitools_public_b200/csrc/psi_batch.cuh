// psi_batch.cuh -- device-resident state machines that sit on top of one D(omega) evaluation:
//
//   K2 contour_count_warp : ClosedPath.count_roots (complex_roots.py:517-625) -- adaptive mid-point
//                           refinement of a closed contour and the winding number of D around it;
//   K3 polish_warp        : PSIMode.find_dispersion_roots / find_root (psi_mode.py:340-352, 411-491) --
//                           two-stage secant polish (scipy.optimize.newton's secant rule) of a list of
//                           start values on the exact dispersion relation.
//
// One warp owns one contour / one root search and evaluates every D it needs inline with the
// warp-cooperative quadrature of psi_quad.cuh, so a whole batch is ONE kernel launch with no host
// round trip at all; items are pulled from a device-side atomic queue.  PSI_HD like the rest, so the
// test-only lane-emulation harness runs the same code on the CPU.
#pragma once
#include "psi_quad.cuh"

namespace psi {

// ---- complex helpers with numpy's semantics (the host engine uses the same rules) --------------------
// numpy's complex division: Smith's algorithm; a zero divisor gives (re/0, im/0)
PSI_HD cx np_div(cx a, cx b) {
    const double abr = fabs(b.re), abi = fabs(b.im);
    if (abr >= abi) {
        if (abr == 0.0 && abi == 0.0) return cx(a.re / abr, a.im / abr);
        const double rat = b.im / b.re, scl = 1.0 / (b.re + b.im * rat);
        return cx((a.re + a.im * rat) * scl, (a.im - a.re * rat) * scl);
    }
    const double rat = b.re / b.im, scl = 1.0 / (b.im + b.re * rat);
    return cx((a.re * rat + a.im) * scl, (a.im * rat - a.re) * scl);
}
PSI_HD cx np_mul(cx a, cx b) {
    return cx(mul_rn(a.re, b.re) - mul_rn(a.im, b.im), mul_rn(a.re, b.im) + mul_rn(a.im, b.re));
}
PSI_HD double abs_c(cx a) { return hypot(a.re, a.im); }
PSI_HD bool finite_c(cx a) { return isfinite(a.re) && isfinite(a.im); }
PSI_HD bool eq_c(cx a, cx b) { return a.re == b.re && a.im == b.im; }

struct RectDom {
    double xmin, xmax, ymin, ymax;
    PSI_HD bool is_in(cx z) const { return z.re > xmin && z.re < xmax && z.im > ymin && z.im < ymax; }
};

// companion start value z0*(1+eps) +- eps (psi_mode.py:431-433)
PSI_HD cx second_point(cx z0, double eps) {
    cx z1((1.0 + eps) * z0.re, (1.0 + eps) * z0.im);
    z1.re += (z1.re >= 0.0 ? eps : -eps);
    return z1;
}

// ---- K3: secant polish --------------------------------------------------------------------------------
struct SecantOut {
    cx root;
    int converged, iterations, calls;
};

// scipy.optimize.newton, secant branch (fprime=None, rtol=0), driven by an evaluator f(z) -> D.
template <class EVAL>
PSI_NOINLINE SecantOut secant_run(EVAL& f, cx x0, cx x1, int maxiter, double xtol) {
    SecantOut o;
    cx p0 = x0, p1 = x1;
    cx q0 = f(p0), q1 = f(p1);
    o.calls = 2;
    if (abs_c(q1) < abs_c(q0)) { cx t = p0; p0 = p1; p1 = t; t = q0; q0 = q1; q1 = t; }
    cx p = p1;
    for (int it = 0; it < maxiter; ++it) {
        if (eq_c(q1, q0)) {
            o.root = cx(0.5 * (p1.re + p0.re), 0.5 * (p1.im + p0.im));
            o.converged = 0; o.iterations = it + 1;
            return o;
        }
        if (abs_c(q1) > abs_c(q0)) {
            const cx r = np_div(q0, q1);
            p = np_div(np_mul(-r, p1) + p0, cx(1.0 - r.re, -r.im));
        } else {
            const cx r = np_div(q1, q0);
            p = np_div(np_mul(-r, p0) + p1, cx(1.0 - r.re, -r.im));
        }
        if (abs_c(p - p1) <= xtol) { o.root = p; o.converged = 1; o.iterations = it + 1; return o; }
        p0 = p1; q0 = q1; p1 = p;
        if (it + 1 < maxiter) q1 = f(p1);     // scipy evaluates once more before giving up; the value
        o.calls += 1;                         // would be unused, so only the call is counted
    }
    o.root = p; o.converged = 0; o.iterations = maxiter;
    return o;
}

// Polish `n` start values one after the other (psi_mode.py:340-352 + 411-491): five iterations first;
// if that has not converged but stayed within |start| of the start, continue from there for up to
// max_iter.  The step tolerance is min(1.48e-8, 1e-5*|Im start_i|) of THIS start value alone
// (find_dispersion_roots hands find_root one scalar at a time, psi_mode.py:340-352, so the w0 array of
// :438-444 has a single element) and of its restart point after a restart (:455-464).  w0[] is updated in
// place with the restart points; out[i] receives the root, or the sentinel -1j when the iteration failed or
// left the domain.  Returns the number of dispersion evaluations (PSIMode.n_function_call increment);
// status 1 = the reference would raise ValueError (this start value lies on the real axis, which makes
// its tolerance zero).
template <class EVAL>
PSI_NOINLINE int polish_run(EVAL& f, const RectDom& dom, cx* w0, int n, int max_iter, cx* out, int* max_conv_iter,
                      int* status) {
    int calls = 0;
    *status = 0;
    for (int i = 0; i < n; ++i) {
        const cx start = w0[i];
        double xtol = fmin(1.48e-8, fabs(start.im * 1e-5));
        if (!(xtol > 0.0)) { *status = 1; return calls; }
        SecantOut r = secant_run(f, start, second_point(start, 1.0e-6), 5, xtol);
        calls += r.calls;
        if (!r.converged && abs_c(r.root - start) / abs_c(start) < 1.0 && max_iter > 5) {
            w0[i] = r.root;
            xtol = fmin(1.48e-8, fabs(r.root.im * 1e-5));
            if (!(xtol > 0.0)) { *status = 1; return calls; }
            r = secant_run(f, w0[i], second_point(w0[i], 1.0e-6), max_iter, xtol);
            calls += r.calls;
        }
        if (r.converged && dom.is_in(r.root)) {
            if (r.iterations > *max_conv_iter) *max_conv_iter = r.iterations;
            out[i] = r.root;
        } else {
            out[i] = cx(0.0, -1.0);
        }
    }
    return calls;
}

// ---- K2: contour root count ---------------------------------------------------------------------------
enum ContourStatus { kContourOk = 0, kContourNoInteger = 1, kContourMaxIter = 2, kContourNonFinite = 3,
                     kContourCapacity = 4 };

// pts/fv and pts2/fv2: two scratch buffers of `cap` points each (ping-pong); the corners are expected
// in pts[0..n0).  Returns the status; *count gets the winding number, *n_out the final point count.
template <class EVAL>
PSI_HD int contour_run(EVAL& f, cx* pts, cx* fv, cx* pts2, cx* fv2, int n0, int cap, double tol,
                       double max_step, int max_iter, int* count, int* n_out) {
    int n = n0;
    for (int i = 0; i < n; ++i) fv[i] = f(pts[i]);                 // ClosedPath.__init__ (:526-531)
    *count = -666;
    for (int pass = 0; pass < max_iter; ++pass) {
        // refine_select (:556-595): decisions use the arrays of the START of the pass
        int k = 0, added = 0;
        bool bad = false;
        for (int i = 0; i < n; ++i) {
            const int im = (i + n - 1) % n, ip = (i + 1) % n, imm = (i + n - 2) % n;
            const double len_i = abs_c(pts[i] - pts[im]);
            const double len_m = abs_c(pts[im] - pts[imm]);
            const double len_p = abs_c(pts[ip] - pts[i]);
            const double jump = atan2(fv[i].im, fv[i].re) - atan2(fv[im].im, fv[im].re);
            const bool uneven = (len_i > 2.0 * len_m) || (len_i > 2.0 * len_p);
            if (((fabs(jump) > tol) || uneven || (len_i > max_step)) && len_i > 1.0e-8) {
                if (k + 2 > cap) { *n_out = n; return kContourCapacity; }
                const cx mid(0.5 * (pts[i].re + pts[im].re), 0.5 * (pts[i].im + pts[im].im));
                const cx fm = f(mid);
                if (!finite_c(fm)) bad = true;
                pts2[k] = mid; fv2[k] = fm; ++k; ++added;
            }
            if (k + 1 > cap) { *n_out = n; return kContourCapacity; }
            pts2[k] = pts[i]; fv2[k] = fv[i]; ++k;
        }
        if (bad) { *n_out = n; return kContourNonFinite; }
        if (added == 0) {
            // sum_angles (:533-538)
            double sum = 0.0;
            for (int i = 0; i < n; ++i) {
                const cx a(fv[i].re, fv[i].im + 1.0e-30), b(fv[(i + 1) % n].re, fv[(i + 1) % n].im + 1.0e-30);
                const cx q = np_div(b, a);
                sum += atan2(q.im, q.re);
            }
            const double winding = 0.5 * sum / 3.141592653589793, nearest = rint(winding);
            *n_out = n;
            if (fabs(winding - nearest) < 1.0e-6) { *count = (int)nearest; return kContourOk; }
            return kContourNoInteger;
        }
        n = k;
        cx* t = pts; pts = pts2; pts2 = t;
        t = fv; fv = fv2; fv2 = t;
    }
    *n_out = n;
    return kContourMaxIter;
}

// Evaluator handed to the state machines: one warp-cooperative D(omega).  The work class of the point
// (size-density family x viscous) is dispatched at run time inside ONE non-inlined function, so the
// kernels K2-K4 share a single compiled copy of the eight quadrature specialisations.
template <class W>
struct WarpEval {
    const NodeTableView* t;
    const DistParams* d;
    double kx, kz, alpha;
    unsigned long long nodes;
    long long evals;
    PSI_NOINLINE cx operator()(cx w) {
        ++evals;
        return disp_eval_any<W>(*t, *d, w, kx, kz, alpha, nullptr, nodes);
    }
};

// Same, for ONE work class: the GPU kernels K2-K4 are compiled per class so that each carries a single
// quadrature specialisation (the eight-way version is ~0.7 MB of machine code and thrashes the
// instruction cache when every warp of an SM sits in a different phase of its state machine).
template <class W, int NK, bool VISC>
struct WarpEvalClass {
    const NodeTableView* t;
    const DistParams* d;
    double kx, kz, alpha;
    unsigned long long nodes;
    long long evals;
    PSI_NOINLINE cx operator()(cx w) {
        ++evals;
        return disp_eval_warp<W, NK, VISC>(*t, *d, w, kx, kz, alpha, nullptr, nodes);
    }
};

}  // namespace psi
