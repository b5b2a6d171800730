// psi_quad.cuh -- warp-cooperative tanh-sinh quadrature and one full D(omega) evaluation.
//
// One warp owns one (Kx, Kz, omega) point.  Lanes split the nodes of a quadrature level, keep the
// partial sums of all seven integrands in registers and combine them with xor-butterfly shuffles.
// The code is written against a small "warp policy" W (lane id, width, reductions) so that the
// test-only harness in tests/hostsim can instantiate it with a one-lane policy on the CPU.
//
// Replaces tanhsinh.py:72-128 (TanhSinh.integrate), psi_dispersion.py:124-157 (average) and
// psi_dispersion.py:375-456 (calculate).
#pragma once
#include "psi_math.cuh"

namespace psi {

#ifndef PSI_PAIR_UNROLL
#define PSI_PAIR_UNROLL 1
#endif
constexpr int kPairUnroll = PSI_PAIR_UNROLL;    // node pairs per lane in flight in the level loop

// Node table in LEVEL-MAJOR order: level 0 holds the nodes j = k*2^max_m (k = 0, 1, ...), level m
// (1 <= m <= max_m) the odd multiples j = (2i+1)*2^(max_m-m); each entry is (x_j, w_j).  Within a
// level a warp therefore reads consecutive 16-byte entries (no shared-memory bank conflicts).
struct alignas(16) node_t { double x, y; };     // (abscissa x_j, weight w_j)

struct NodeTableView {
    const node_t* nodes;      // level-major (x, w) pairs
    const int* level_off;     // max_m + 2 offsets into nodes[]
    int n_nodes;              // total number of nodes (original index j in [0, n_nodes))
    int max_m;
    double eps;               // 10^-precision_digit: end-point guard AND default tolerance
    int predict_exit;         // 1: skip confirming levels whose increment is predicted far below tol (below);
                              // 2: the same without the noise-floor check (thread-per-point fast path of K1)
    int level_cap;            // > 0: give up after this level if a component is still running (results = NaN)
};

PSI_HD int ctz_int(int v) {
#if defined(__CUDA_ARCH__)
    return __ffs(v) - 1;
#else
    return __builtin_ctz(v);
#endif
}

// (x_j, w_j) by ORIGINAL index j
PSI_HD node_t node_by_index(const NodeTableView& t, int j) {
    if (j == 0) return t.nodes[0];
    int tz = ctz_int(j);
    if (tz >= t.max_m) return t.nodes[j >> t.max_m];            // level 0
    int lvl = t.max_m - tz;
    return t.nodes[t.level_off[lvl] + ((j >> tz) >> 1)];
}

// tanhsinh.py:68-70 with the reference's operation order and no FMA contraction
PSI_HD double abscissa(double x, double bma, double a, double b) {
    return 0.5 * add_rn(add_rn(mul_rn(x, bma), b), a);
}

// Number of leading nodes (original order) whose image stays below b - eps (tanhsinh.py:95-98).
// The image is monotone in x, so the kept set is a prefix; for b <= a it is empty.
PSI_HD int kept_nodes(const NodeTableView& t, double a, double b) {
    const double bma = b - a;
    const double lim = b - t.eps;
    if (!(bma > 0.0)) return 0;     // image never drops below b: nothing survives the guard
    int lo = 0, hi = t.n_nodes;               // invariant: nodes [0, lo) kept, [hi, n) dropped
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (abscissa(node_by_index(t, mid).x, bma, a, b) < lim) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Generic joint tanh-sinh over [a, b] of NC complex components.
//   gen  : NodeGen<..> already begin()'ed on [a, b]: turns an abscissa pair +-x into (tau, 1/tau, weight)
//   fn   : fn(node, wq, acc) does acc[k] += wq * node.wt * integrand_k(node.tau)
//   cscale[k] : constant factor the caller applies to component k afterwards (the exit test sees it)
// Each component keeps the reference's own early exit: it stops, WITHOUT adding the last
// increment, once |increment| < tol (tanhsinh.py:120-126).  `res` receives the NC results and
// `nodes_done` is increased by the number of integrand evaluations this warp performed.
template <class W, int NC, class GEN, class FN>
PSI_HD void tanhsinh_warp(const NodeTableView& t, double a, double b, double tol, const GEN& gen,
                          const FN& fn, const double* cscale, cx* res,
                          unsigned long long& nodes_done) {
    const int lane = W::lane();
    const double half = 0.5 * (b - a);
    const int nsel = kept_nodes(t, a, b);
    unsigned active = (1u << NC) - 1u;
    unsigned long long evals = 0;

    cx acc[NC];
    // ---- level 0: stride 2^max_m, x = 0 counted once (tanhsinh.py:101-104)
#pragma unroll
    for (int k = 0; k < NC; ++k) acc[k] = cx(0.0, 0.0);
    {
        const int cnt = nsel > 0 ? ((nsel - 1) >> t.max_m) + 1 : 0;     // j = k << max_m < nsel
        for (int k = lane; k < cnt; k += W::width) {
            const node_t nd = t.nodes[k];
            if (k == 0) {
                fn(gen.centre(), nd.y, acc);
            } else {
                NodeVals np, nn;
                gen.pair(nd.x, np, nn);
                fn(np, nd.y, acc);
                fn(nn, nd.y, acc);
            }
        }
        evals += cnt > 0 ? 2 * cnt - 1 : 0;
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        res[k].re = half * W::sum(acc[k].re);
        res[k].im = half * W::sum(acc[k].im);
    }
    // ---- levels 1..max_m: odd multiples of stride 2^(max_m-m) (tanhsinh.py:107-126)
    double prev[NC];                       // previous increment of every component
    unsigned falling = 0u;                 // bit k: that increment was already < 0.1 of the one before it
#pragma unroll
    for (int k = 0; k < NC; ++k) prev[k] = INFINITY;
    const bool predict = t.predict_exit != 0;
    const bool loose = t.predict_exit == 2;
    const int last_level = (t.level_cap > 0 && t.level_cap < t.max_m) ? t.level_cap : t.max_m;
    double h = 1.0;
    for (int m = 1; m <= last_level; ++m) {
        h *= 0.5;
        const int sh = t.max_m - m;
        // (2i+1) << sh < nsel  <=>  i < ((nsel-1 >> sh) + 1) >> 1
        const int cnt = nsel > 0 ? ((((nsel - 1) >> sh) + 1) >> 1) : 0;
        const node_t* lv = t.nodes + t.level_off[m];
#pragma unroll
        for (int k = 0; k < NC; ++k) acc[k] = cx(0.0, 0.0);
#pragma unroll kPairUnroll
        for (int i = lane; i < cnt; i += W::width) {
            const node_t nd = lv[i];
            NodeVals np, nn;
            gen.pair(nd.x, np, nn);
            fn(np, nd.y, acc);
            fn(nn, nd.y, acc);
        }
        evals += 2ull * cnt;
        const double hh = h * half;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            if (active & (1u << k)) {
                cx dres(-0.5 * res[k].re + hh * W::sum(acc[k].re),
                        -0.5 * res[k].im + hh * W::sum(acc[k].im));
                const double a = cabs(dres) * cscale[k];
                if (a < tol) active &= ~(1u << k);
                else {
                    res[k].re += dres.re; res[k].im += dres.im;
                    // Convergence predictor (predict_exit).  The reference's rule evaluates levels until an
                    // increment falls below tol = 1e-15 ABSOLUTE and then discards it; for |I| >~ 1 that is
                    // below the rounding noise of the sums, so the last levels -- each as expensive as all
                    // earlier ones together -- only resolve noise.  Tanh-sinh increments collapse
                    // quadratically (d_{m+1} ~ d_m^2/|I|).  When the last two increments SHOW that collapse
                    // (d_m < 1e-3 d_{m-1}, d_{m-1} < 0.1 d_{m-2}, d_m |I| <= 100 d_{m-1}^2) and the
                    // extrapolated next one, 1e4 d_m^2/|I|, is below tol, the remaining levels are skipped.
                    // The value returned is the one the reference's rule returns when its next increment
                    // is indeed below tol (the sum after THIS level).
                    if (predict) {
                        const double s = cabs(res[k]) * cscale[k];
                        if (a < 1e-3 * prev[k] && (falling & (1u << k)) && 1e4 * a * a < tol * s &&
                            (loose || a * s <= 100.0 * prev[k] * prev[k])) active &= ~(1u << k);
                    }
                    if (a < 0.1 * prev[k]) falling |= 1u << k; else falling &= ~(1u << k);
                    prev[k] = a;
                }
            }
        }
        if (active == 0u) break;
    }
    if (last_level < t.max_m && active != 0u) {          // capped run that did not converge: say so
#pragma unroll
        for (int k = 0; k < NC; ++k) res[k] = cx(NAN, NAN);
    }
    nodes_done += evals;
}

// mu * sum over sub-intervals [ln smin, ln p1], [ln p1, ln p2], ..., [ln pn, 0] of the joint
// integral of g(taumax*e^u) * e^u * F(e^u) du                      (psi_dispersion.py:135-157)
template <class W, int NC, int NK, class FN>
PSI_HD void average_kind(const NodeTableView& t, const DistParams& d, const double* poles, int npoles,
                         const FN& fn, const double* cscale, cx* out, unsigned long long& nodes_done) {
#pragma unroll
    for (int k = 0; k < NC; ++k) out[k] = cx(0.0, 0.0);
    double lo = log(d.taumin / d.taumax);
    NodeGen<NK> gen;
    for (int s = 0; s <= npoles; ++s) {
        const double hi = (s < npoles) ? log(poles[s]) : 0.0;
        cx r[NC];
        gen.begin(d, lo, hi);
        tanhsinh_warp<W, NC>(t, lo, hi, t.eps, gen, fn, cscale, r, nodes_done);
#pragma unroll
        for (int k = 0; k < NC; ++k) { out[k].re += r[k].re; out[k].im += r[k].im; }
        lo = hi;
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) { out[k].re *= d.mu; out[k].im *= d.mu; }
}

// K0: background integrals J0, J1 and the derived equilibrium velocities (psi_dispersion.py:74-84).
// Writes j0, j1, denom, vgx, vgy into d (all lanes hold identical values).
template <class W>
PSI_HD void background_warp(const NodeTableView& t, DistParams& d, unsigned long long& nodes_done) {
    cx j[2];
    if (d.single_size) {
        const double r1 = 1.0 / (1.0 + d.taumax * d.taumax);
        j[0] = cx(d.mu * r1, 0.0);
        j[1] = cx(d.mu * d.taumax * r1, 0.0);
    } else {
        const double one[2] = {1.0, 1.0};
        switch (node_kind_of(d)) {            // warp-uniform
        case kNodePowerSqrt: average_kind<W, 2, kNodePowerSqrt>(t, d, d.init_poles, d.n_init_poles, JNode(), one, j, nodes_done); break;
        case kNodePowerGen:  average_kind<W, 2, kNodePowerGen>(t, d, d.init_poles, d.n_init_poles, JNode(), one, j, nodes_done); break;
        case kNodeBump:      average_kind<W, 2, kNodeBump>(t, d, d.init_poles, d.n_init_poles, JNode(), one, j, nodes_done); break;
        case kNodeBumpTail:  average_kind<W, 2, kNodeBumpTail>(t, d, d.init_poles, d.n_init_poles, JNode(), one, j, nodes_done); break;
        default:             average_kind<W, 2, kNodeTable>(t, d, d.init_poles, d.n_init_poles, JNode(), one, j, nodes_done); break;
        }
    }
    d.j0 = j[0].re;
    d.j1 = j[1].re;
    d.denom = 1.0 / ((1.0 + d.j0) * (1.0 + d.j0) + d.j1 * d.j1);
    d.vgx = 2.0 * d.j1 * d.denom;
    d.vgy = -(1.0 + d.j0) * d.denom;
}

// Work class of a point: which specialisation of K1 evaluates it (kernels are compiled per class so
// that each keeps its own, smaller register footprint).  class = node kind * 2 + (viscous ? 1 : 0).
constexpr int kNumClasses = 10;
constexpr double kNearPoleBand = 2.5;      // |Re d(tau*)| below which the damped-frequency split point is added
constexpr double kResonantBand = 1.0e-5;   // |Im(w - kx ux + i(...))| at a resonance pole below which the exit predictor is off
PSI_HD int point_class(const DistParams& d, double alpha) {
    return node_kind_of(d) * 2 + ((alpha * d.c * d.c) != 0.0 ? 1 : 0);
}

// K1 body: one D(omega; Kx, Kz, alpha) evaluation by one warp, for a point of class (NK, VISC).
// Returns D; m_out (7 entries, may be null) gets M.
template <class W, int NK, bool VISC>
PSI_HD cx disp_eval_warp(const NodeTableView& t, const DistParams& d, cx w, double kx, double kz,
                         double alpha, cx* m_out, unsigned long long& nodes_done) {
    const PointConst p = make_point(d, w, kx, kz, alpha);
    cx m[7];
    if (d.single_size) {
        // average() shortcut: mu * g(taumax)                       (psi_dispersion.py:130-131)
#pragma unroll
        for (int k = 0; k < 7; ++k) m[k] = cx(0.0, 0.0);
        NodeVals nv; nv.tau = d.taumax; nv.it = 1.0 / d.taumax; nv.wt = d.mu;
        m_node<VISC>(p, nv, 1.0, m);
    } else {
        double poles[2 * kMaxPoles], res_tau[2];
        int n_res = 0;
        int np = resonance_poles(d, w.re, kx, poles, res_tau, &n_res);
        // Resonant band: next to a resonance pole tau_p the integrands carry 1/(w - kx ux(tau) + i(k^2 D + 1e-10)),
        // whose imaginary part y_p = Im w + k^2 D(tau_p) + 1e-10 is all that keeps it finite there.  The nodes that
        // cluster at the pole cannot be told apart to better than an ulp of u = ln(tau), so every level adds
        // rounding noise of relative size ~3e-17/|y_p| (the reference's own reproducibility floor): the
        // increments reach that plateau instead of collapsing, and a predicted exit would freeze one early sample
        // of the noise (measured: 8e-10 off the literal rule at |Im w| = 1e-7).  Below |y_p| = 1e-5 the point is
        // therefore evaluated under the literal rule of tanhsinh.py:120-126; the thread-per-point fast path hands
        // it to the warp kernel without spending its capped levels on it.
        bool band = false;
        for (int i = 0; i < n_res; ++i) {
            double y = w.im + 1.0e-10;
            if (VISC) {
                const double tp = res_tau[i], t1 = 1.0 + tp * tp;
                y += p.k2 * (1.0 + tp + 4.0 * tp * tp) * p.nu / (t1 * t1);
            }
            if (fabs(y) < kResonantBand) band = true;
        }
        // Damped frequencies with Im w < -1/taumax: at tau* = -1/Im w the imaginary part of d = kx ux - w - i/tau
        // vanishes, and where |Re d(tau*)| is of order one the integrands' factors 1/d and 1/(1 - d^2) have poles
        // within ~1e-2 (in ln tau) of the path -- in the MIDDLE of a sub-interval, where the rule needs its last
        // levels to resolve them (these points, 1-3 % of BASELINE config 2's grid, cost half of its run time).
        // The reference's rule gets them right all the same (1e-14 against 40-digit arithmetic), so under the
        // predicted rule the interval is split at tau* as it is at the resonance poles: the near-poles then sit
        // next to an end point, which tanh-sinh resolves by level 5-6.  exit_rule = "reference" keeps the
        // reference's own sub-intervals.
        if (t.predict_exit != 0 && w.im < 0.0 && np < 2 * kMaxPoles) {
            const double ts = -1.0 / w.im;
            if (ts > d.taumin && ts < d.taumax) {
                const double kxux = p.kxA2 * fma(-ts, p.j0p, p.j1) / fma(ts, ts, 1.0);
                if (fabs(kxux - w.re) <= kNearPoleBand) {
                    const double v = ts / d.taumax;
                    int j = np - 1;
                    while (j >= 0 && poles[j] > v) { poles[j + 1] = poles[j]; --j; }
                    poles[j + 1] = v;
                    ++np;
                }
            }
        }
        double sc[7];
        m_scales(p, sc);
        MNode<VISC> fn; fn.p = &p;
        NodeTableView tl = t;
        if (band && t.predict_exit != 0) {
            if (t.level_cap > 0) return cx(NAN, NAN);
            tl.predict_exit = 0;
        }
        average_kind<W, 7, NK>(tl, d, poles, np, fn, sc, m, nodes_done);
    }
    m_apply_scales(p, m);
    if (m_out) {
#pragma unroll
        for (int k = 0; k < 7; ++k) m_out[k] = m[k];
    }
    return det_P_plus_M(d, p, m);
}

// run-time dispatch over the classes (host harness, small kernels)
template <class W>
PSI_HD cx disp_eval_any(const NodeTableView& t, const DistParams& d, cx w, double kx, double kz,
                        double alpha, cx* m_out, unsigned long long& nodes_done) {
    switch (point_class(d, alpha)) {
    case 0: return disp_eval_warp<W, 0, false>(t, d, w, kx, kz, alpha, m_out, nodes_done);
    case 1: return disp_eval_warp<W, 0, true>(t, d, w, kx, kz, alpha, m_out, nodes_done);
    case 2: return disp_eval_warp<W, 1, false>(t, d, w, kx, kz, alpha, m_out, nodes_done);
    case 3: return disp_eval_warp<W, 1, true>(t, d, w, kx, kz, alpha, m_out, nodes_done);
    case 4: return disp_eval_warp<W, 2, false>(t, d, w, kx, kz, alpha, m_out, nodes_done);
    case 5: return disp_eval_warp<W, 2, true>(t, d, w, kx, kz, alpha, m_out, nodes_done);
    case 6: return disp_eval_warp<W, 3, false>(t, d, w, kx, kz, alpha, m_out, nodes_done);
    case 7: return disp_eval_warp<W, 3, true>(t, d, w, kx, kz, alpha, m_out, nodes_done);
    case 8: return disp_eval_warp<W, 4, false>(t, d, w, kx, kz, alpha, m_out, nodes_done);
    default: return disp_eval_warp<W, 4, true>(t, d, w, kx, kz, alpha, m_out, nodes_done);
    }
}

}  // namespace psi
