// psi_mode.cuh -- kernel K4's body: one whole PSIMode.calculate (psi_mode.py:106-333) by one warp, with
// no host round trip: random samples -> D (inline, warp-cooperative) -> AAA (psi_aaa.cuh) -> zeros ->
// clean_tol halving -> zoom-domain selection -> secant polish (psi_batch.cuh) -> more samples ...
//
// The random sample points come from a pre-generated stream of U[0,1) numbers per work item (the numpy
// legacy MT19937 stream of the item's seed, generated once per distinct seed on the host); the device
// consumes it in the reference's order: for every domain first all real parts, then all imaginary parts
// (complex_roots.py:404-409), lo + (hi-lo)*u with separate multiply and add.
#pragma once
#include "psi_aaa.cuh"

namespace psi {

struct ModeCfg {
    RectDom dom;
    int n_sample, max_zoom, max_secant;
    double tol, clean_tol;
};

enum ModeStatus { kModeOk = 0, kModeBadSamples = 1, kModeCapacity = 2, kModeZeroTol = 3, kModeTruncated = 4 };

// append n_sample random points of the box [x0,x1]x[y0,y1] and their D values
template <class W, class EVAL>
PSI_HD void mode_sample(EVAL& f, const AAAScratch& s, AAAState& st, const double* uni, int& upos, int n,
                        double x0, double x1, double y0, double y1) {
    const int base = st.M;
    for (int i = W::lane(); i < n; i += W::width) {
        const double re = add_rn(x0, mul_rn(x1 - x0, uni[upos + i]));
        const double im = add_rn(y0, mul_rn(y1 - y0, uni[upos + n + i]));
        s.Z[base + i] = cx(re, im);
    }
    W::sync();
    for (int i = 0; i < n; ++i) {
        const cx v = f(s.Z[base + i]);
        if (W::lane() == 0) s.F[base + i] = v;
    }
    W::sync();
    upos += 2 * n;
    st.M = base + n;
}

// zoom box around `centre`, shifted to lie inside the search domain (psi_mode.py:354-397)
template <class W, class EVAL>
PSI_HD void mode_extra_domain(EVAL& f, const AAAScratch& s, AAAState& st, const ModeCfg& cfg, const double* uni,
                              int& upos, double sx, double sy, cx centre) {
    double re = centre.re, im = centre.im;
    if (re < cfg.dom.xmin + 0.5 * sx) re = cfg.dom.xmin + 0.5 * sx;
    if (re > cfg.dom.xmax - 0.5 * sx) re = cfg.dom.xmax - 0.5 * sx;
    if (im < cfg.dom.ymin + 0.5 * sy) im = cfg.dom.ymin + 0.5 * sy;
    if (im > cfg.dom.ymax - 0.5 * sy) im = cfg.dom.ymax - 0.5 * sy;
    mode_sample<W>(f, s, st, uni, upos, cfg.n_sample, re - 0.5 * sx, re + 0.5 * sx, im - 0.5 * sy, im + 0.5 * sy);
}

// What the analysis of one zoom level hands to the polish and to the next sampling step
struct ModePlan {
    int n0;          // number of polish start values (written to s.w0)
    int max_iter;    // secant iteration budget of this level
    int n_zoom;      // zoom boxes to open if no growing root is found
    cx zoom[4];
    double box_dx;   // width of the zoom boxes in Re
};

// AAA + zero selection of one zoom level (psi_mode.py:147-310): rational approximation of the current
// samples, clean_tol halving until enough support points and one growing zero, zoom candidates, start
// values of the polish -- written as a RESUMABLE state machine.  mode_ctl() advances a search through all its
// cheap transitions (index lists, memo lookups, greedy promotion, cleanup decisions, clean_tol halving, the zoom
// plan) and stops when it needs one of three expensive operations:
//     kReqFit   a Loewner fit of the current mask            (aaa_fit_compute:   Loewner matrix, QR, inverse iteration)
//     kReqPoles poles + residue decisions of the current fit (aaa_poles_compute: Ehrlich-Aberth + 4-point contours)
//     kReqZeros zeros of the current fit                     (aaa_zeros_compute: Ehrlich-Aberth + secant on r(z))
// mode_analyze() below drives the state machine in a loop; the split keeps the three expensive operations apart
// from the control flow (they can be timed, replaced and scheduled separately -- psi_b200.cu records the
// experiment that ran each of the four pieces as its own kernel).
enum AnRequest { kReqNone = 0, kReqFit = 1, kReqPoles = 2, kReqZeros = 3, kReqFailed = 4 };
enum AnPc { kPcStart = 0, kPcGreedyNext, kPcGreedyFit, kPcGreedyEnd, kPcClean, kPcCleanApply, kPcCleanFit,
            kPcZeros, kPcZerosDone, kPcHalve, kPcReplayFit, kPcPlan, kPcDone };

struct AnDriver {
    int pc;          // where mode_ctl resumes
    int req;         // expensive operation it is waiting for
    int k;           // greedy step (number of promotions so far + 1)
    int added;       // sample promoted before the pending fit (-1: none; lets the inverse iteration start warm)
    int removed;     // support points dropped by the last cleanup pass
    int nz;          // zeros of the current fit (in s.roots)
    double res;      // largest residual / max|F| of the last fit
};

// the three expensive operations
template <class W>
PSI_HD void mode_heavy(int req, const AAAScratch& s, AAAState& st, AnDriver& d) {
    if (req == kReqFit) d.res = aaa_fit_compute<W>(s, st, d.added);
    else if (req == kReqPoles) aaa_poles_compute<W>(s, st);
    else if (req == kReqZeros) aaa_zeros_compute<W>(s, st);
}

// Runs until the search needs an expensive operation (returned, also in d.req) or is finished (d.pc == kPcDone:
// kReqNone = plan filled in, kReqFailed = non-finite samples, the reference's ValueError).
template <class W>
PSI_HD int mode_ctl(const ModeCfg& cfg, int level, int n_guess, const AAAScratch& s, AAAState& st, AnDriver& d,
                    ModePlan& plan) {
    const int lane = W::lane(), M = st.M;
    d.req = kReqNone;
    for (;;) {
        switch (d.pc) {
        case kPcStart: {                                   // RationalApproximation.calculate (complex_roots.py:126-176)
            st.clean_tol = cfg.clean_tol;
            aaa_memo_clear<W>(s, st);
            double fm = 0.0; bool fin = true;
#pragma unroll 1
            for (int i = lane; i < M; i += W::width) {
                fm = fmax(fm, a_abs(s.F[i]));
                if (!finite_c(s.F[i]) || !finite_c(s.Z[i])) fin = false;
            }
            st.fmax = a_max<W>(fm);
            if (W::any(!fin)) { d.pc = kPcDone; d.req = kReqFailed; return d.req; }
#pragma unroll 1
            for (int i = lane; i < M; i += W::width) s.mask[i] = (i < 2) ? 1 : 0;
            W::sync();
            st.v_m = 0;
            d.k = 1; d.added = -1;
            d.pc = kPcGreedyNext;                          // the two-point fit never ends the greedy loop
            if (!aaa_fit_lookup<W>(s, st, &d.res)) { d.req = kReqFit; return d.req; }
            continue;
        }
        case kPcGreedyNext: {
            if (d.k >= M / 2) { d.pc = kPcGreedyEnd; continue; }
            // promote the worst-fitted sample (first maximum, like numpy.argmax)
            int best = -1; double bv = -1.0;
            for (int i = lane; i < st.r; i += W::width)
                if (s.R[i] > bv) { bv = s.R[i]; best = i; }
            W::argmax(bv, best);
            const int added = s.off[best];
            W::sync();
            if (lane == 0) s.mask[added] = 1;
            W::sync();
            d.added = added; d.k += 1;
            d.pc = kPcGreedyFit;
            if (!aaa_fit_lookup<W>(s, st, &d.res)) { d.req = kReqFit; return d.req; }
            continue;
        }
        case kPcGreedyFit:
            d.pc = d.res < st.tol ? kPcGreedyEnd : kPcGreedyNext;
            continue;
        case kPcGreedyEnd:
#pragma unroll 1
            for (int i = lane; i < M; i += W::width) s.g_mask[i] = s.mask[i];
            W::sync();
            d.pc = kPcClean;
            continue;
        case kPcClean:                                     // cleanup passes until nothing is removed (:286-329)
            d.pc = kPcCleanApply;
            if (!aaa_poles_ready(s, st)) { d.req = kReqPoles; return d.req; }
            continue;
        case kPcCleanApply:
            d.removed = aaa_cleanup_apply<W>(s, st);
            d.added = -1;
            d.pc = kPcCleanFit;
            if (!aaa_fit_lookup<W>(s, st, &d.res)) { d.req = kReqFit; return d.req; }
            continue;
        case kPcCleanFit:
            d.pc = d.removed != 0 ? kPcClean : kPcZeros;
            continue;
        case kPcZeros:
            d.pc = kPcHalve;
            if (!aaa_zeros_load<W>(s, st, &d.nz)) { d.pc = kPcZerosDone; d.req = kReqZeros; return d.req; }
            continue;
        case kPcZerosDone:
            d.nz = st.m - 1 <= 0 ? 0 : s.memo_nz[st.slot];
            d.pc = kPcHalve;
            continue;
        case kPcHalve: {
            // halve clean_tol until enough support points and one growing zero inside the domain (:174-217)
            const int need = 4 * (1 + level + n_guess);
            bool growing = false;
            for (int k = 0; k < d.nz; ++k) {
                const cx z = s.roots[k];
                if (cfg.dom.is_in(z) && z.im > 0.0) growing = true;
            }
            if (st.m >= need && growing) { d.pc = kPcPlan; continue; }
            st.clean_tol = 0.5 * st.clean_tol;
            if (st.clean_tol < st.tol) { d.pc = kPcPlan; continue; }
            // calculate() again on the same samples: the greedy phase does not depend on clean_tol -> replay it
            W::sync();
#pragma unroll 1
            for (int i = lane; i < M; i += W::width) s.mask[i] = s.g_mask[i];
            W::sync();
            d.added = -1;
            d.pc = kPcReplayFit;
            if (!aaa_fit_lookup<W>(s, st, &d.res)) { d.req = kReqFit; return d.req; }
            continue;
        }
        case kPcReplayFit:
            d.pc = kPcClean;
            continue;
        case kPcPlan: {
            const int nz = d.nz;
            // zoom candidates: below the top of the box, inside in Re; by decreasing Im, at least box_dx
            // apart in Re, at most four (:236-261)
            plan.box_dx = pow(10.0, -1.0 - 0.5 * level);
            plan.n_zoom = 0;
            {
                int nc = 0;
                for (int k = 0; k < nz; ++k) {
                    const cx z = s.roots[k];
                    if (z.im < cfg.dom.ymax && z.re < cfg.dom.xmax && z.re > cfg.dom.xmin) ++nc;
                }
                W::sync();
                for (int k = lane; k < nz; k += W::width) {       // rank by Im (ascending) into tmp
                    const cx z = s.roots[k];
                    if (!(z.im < cfg.dom.ymax && z.re < cfg.dom.xmax && z.re > cfg.dom.xmin)) continue;
                    int rank = 0;
                    for (int l = 0; l < nz; ++l) {
                        const cx y = s.roots[l];
                        if (!(y.im < cfg.dom.ymax && y.re < cfg.dom.xmax && y.re > cfg.dom.xmin)) continue;
                        if (y.im < z.im || (y.im == z.im && l < k)) ++rank;
                    }
                    s.tmp[rank] = z;
                }
                W::sync();
                for (int i = nc - 1; i >= 0 && plan.n_zoom < 4; --i) {
                    const cx z = s.tmp[i];
                    bool ok = true;
                    for (int q = 0; q < plan.n_zoom; ++q) if (fabs(z.re - plan.zoom[q].re) < plan.box_dx) ok = false;
                    if (ok) plan.zoom[plan.n_zoom++] = z;
                }
            }
            // start values of the polish: zeros inside the domain, else the least damped candidate
            int n0 = 0;
            W::sync();
            if (lane == 0) {
                int c = 0;
                for (int k = 0; k < nz; ++k) if (cfg.dom.is_in(s.roots[k])) s.w0[c++] = s.roots[k];
            }
            for (int k = 0; k < nz; ++k) if (cfg.dom.is_in(s.roots[k])) ++n0;
            plan.max_iter = cfg.max_secant;
            if (level < cfg.max_zoom) plan.max_iter = cfg.n_sample;
            if (n0 == 0 && plan.n_zoom > 0) {
                if (lane == 0) s.w0[0] = plan.zoom[0];
                n0 = 1;
                plan.max_iter = 5;
            }
            W::sync();
            plan.n0 = n0;
            d.pc = kPcDone;
            return kReqNone;
        }
        default:
            return kReqNone;
        }
    }
}

// One zoom level in one go (single-warp form).  Returns false when the samples are not finite (ValueError in
// the reference).
template <class W>
PSI_HD bool mode_analyze(const ModeCfg& cfg, int level, int n_guess, const AAAScratch& s, AAAState& st,
                         ModePlan& plan) {
    AnDriver d;
    d.pc = kPcStart; d.req = kReqNone; d.k = 1; d.added = -1; d.removed = 0; d.nz = 0; d.res = 0.0;
    for (;;) {
        const int req = mode_ctl<W>(cfg, level, n_guess, s, st, d, plan);
        if (req == kReqFailed) return false;
        if (d.pc == kPcDone) return true;
        mode_heavy<W>(req, s, st, d);
    }
}

// Returns a ModeStatus; out_roots[0..min(*n_roots, max_roots)) receive the roots inside the domain,
// *n_calls the reference's n_function_call.  (Single-warp form of the whole search; the GPU library runs
// the same pieces as a pipeline of homogeneous kernels, see mode_pipeline in psi_b200.cu.)
struct ModeDetail { int M, upos, conv; };    // samples used, random numbers consumed, max iterations of a converged polish

template <class W, class EVAL>
PSI_HD int mode_run(EVAL& f, const ModeCfg& cfg, const cx* guesses, int n_guess, const double* uni,
                    const AAAScratch& s, cx* out_roots, int max_roots, int* n_roots, int* n_calls,
                    ModeDetail* detail = nullptr) {
    AAAState st;
    st.M = 0; st.m = 0; st.r = 0; st.fmax = 0.0; st.tol = cfg.tol; st.clean_tol = cfg.clean_tol; st.dom = cfg.dom;
    st.memo_next = 0; st.slot = 0; st.v_m = 0;
    int upos = 0, calls = 0, conv_max = 0;
    *n_roots = 0; *n_calls = 0;
    if (detail) { detail->M = 0; detail->upos = 0; detail->conv = 0; }
    const int n_domains_max = 1 + n_guess + 4 * cfg.max_zoom;
    if (cfg.n_sample * n_domains_max > s.cap) return kModeCapacity;

    mode_sample<W>(f, s, st, uni, upos, cfg.n_sample, cfg.dom.xmin, cfg.dom.xmax, cfg.dom.ymin, cfg.dom.ymax);
    calls += cfg.n_sample;
    for (int g = 0; g < n_guess; ++g) {
        const double side = fmin(1.0e-3, fabs(guesses[g].im) * 4.0);      // guess_domain_size (:34-37)
        mode_extra_domain<W>(f, s, st, cfg, uni, upos, side, side, guesses[g]);
        calls += cfg.n_sample;
    }

    int n_out = 0;
    for (int level = 0; level <= cfg.max_zoom; ++level) {
        ModePlan plan;
        if (!mode_analyze<W>(cfg, level, n_guess, s, st, plan)) { *n_calls = calls; return kModeBadSamples; }
        int conv = 0, pst = 0;
        calls += polish_run(f, cfg.dom, s.w0, plan.n0, plan.max_iter, s.pol, &conv, &pst);
        W::sync();
        if (conv > conv_max) conv_max = conv;
        if (detail) { detail->M = st.M; detail->upos = upos; detail->conv = conv_max; }
        if (pst != 0) { *n_calls = calls; return kModeZeroTol; }
        n_out = 0;
        bool grow = false;
        for (int k = 0; k < plan.n0; ++k) {
            const cx z = s.pol[k];
            if (!cfg.dom.is_in(z)) continue;
            if (n_out < max_roots && W::lane() == 0) out_roots[n_out] = z;
            ++n_out;
            if (z.im > 0.0) grow = true;
        }
        if (grow || level == cfg.max_zoom || plan.n_zoom == 0) break;
        for (int q = 0; q < plan.n_zoom; ++q) {
            mode_extra_domain<W>(f, s, st, cfg, uni, upos, plan.box_dx, fmin(0.1, fabs(plan.zoom[q].im)), plan.zoom[q]);
            calls += cfg.n_sample;
        }
    }
    *n_roots = n_out;
    *n_calls = calls;
    if (detail) { detail->M = st.M; detail->upos = upos; detail->conv = conv_max; }
    return n_out > max_roots ? kModeTruncated : kModeOk;
}

}  // namespace psi
