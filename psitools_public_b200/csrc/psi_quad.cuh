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
//   f(u, wt, acc) must do acc[k] += wt * integrand_k(u).
// Each component keeps the reference's own early exit: it stops, WITHOUT adding the last
// increment, once |increment| < tol (tanhsinh.py:120-126).  `res` receives the NC results and
// `nodes_done` is increased by the number of integrand evaluations this warp performed.
template <class W, int NC, class F>
PSI_HD void tanhsinh_warp(const NodeTableView& t, double a, double b, double tol, F f, cx* res,
                          unsigned long long& nodes_done) {
    const int lane = W::lane();
    const double bma = b - a;
    const double half = 0.5 * bma;
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
            f(abscissa(nd.x, bma, a, b), nd.y, acc);
            if (k > 0) f(abscissa(-nd.x, bma, a, b), nd.y, acc);
        }
        evals += cnt > 0 ? 2 * cnt - 1 : 0;
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        res[k].re = half * W::sum(acc[k].re);
        res[k].im = half * W::sum(acc[k].im);
    }
    // ---- levels 1..max_m: odd multiples of stride 2^(max_m-m) (tanhsinh.py:107-126)
    double h = 1.0;
    for (int m = 1; m <= t.max_m; ++m) {
        h *= 0.5;
        const int sh = t.max_m - m;
        // (2i+1) << sh < nsel  <=>  i < ((nsel-1 >> sh) + 1) >> 1
        const int cnt = nsel > 0 ? ((((nsel - 1) >> sh) + 1) >> 1) : 0;
        const node_t* lv = t.nodes + t.level_off[m];
#pragma unroll
        for (int k = 0; k < NC; ++k) acc[k] = cx(0.0, 0.0);
        for (int i = lane; i < cnt; i += W::width) {
            const node_t nd = lv[i];
            f(abscissa(nd.x, bma, a, b), nd.y, acc);
            f(abscissa(-nd.x, bma, a, b), nd.y, acc);
        }
        evals += 2ull * cnt;
        const double hh = h * half;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            if (active & (1u << k)) {
                cx dres(-0.5 * res[k].re + hh * W::sum(acc[k].re),
                        -0.5 * res[k].im + hh * W::sum(acc[k].im));
                if (cabs(dres) < tol) active &= ~(1u << k);
                else { res[k].re += dres.re; res[k].im += dres.im; }
            }
        }
        if (active == 0u) break;
    }
    nodes_done += evals;
}

// mu * sum over sub-intervals [ln smin, ln p1], [ln p1, ln p2], ..., [ln pn, 0] of the joint
// integral of g(taumax*e^u) * e^u * F(e^u) du                      (psi_dispersion.py:135-157)
template <class W, int NC, class G>
PSI_HD void average_warp(const NodeTableView& t, const DistParams& d, const double* poles, int npoles,
                         G g, cx* out, unsigned long long& nodes_done) {
#pragma unroll
    for (int k = 0; k < NC; ++k) out[k] = cx(0.0, 0.0);
    const double a0 = log(d.taumin / d.taumax);
    auto f = [&](double u, double wq, cx* acc) {
        const double e = exp(u);
        g(d.taumax * e, wq * (e * F_of(d, e)), acc);
    };
    double lo = a0;
    for (int s = 0; s <= npoles; ++s) {
        const double hi = (s < npoles) ? log(poles[s]) : 0.0;
        cx r[NC];
        tanhsinh_warp<W, NC>(t, lo, hi, t.eps, f, r, nodes_done);
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
        average_warp<W, 2>(t, d, d.init_poles, d.n_init_poles,
                           [](double tau, double wt, cx* acc) { j_integrands_acc(tau, wt, acc); }, j,
                           nodes_done);
    }
    d.j0 = j[0].re;
    d.j1 = j[1].re;
    d.denom = 1.0 / ((1.0 + d.j0) * (1.0 + d.j0) + d.j1 * d.j1);
    d.vgx = 2.0 * d.j1 * d.denom;
    d.vgy = -(1.0 + d.j0) * d.denom;
}

// K1 body: one D(omega; Kx, Kz, alpha) evaluation by one warp.  Returns D; m_out (7 entries) gets M.
template <class W>
PSI_HD cx disp_eval_warp(const NodeTableView& t, const DistParams& d, cx w, double kx, double kz,
                         double alpha, cx* m_out, unsigned long long& nodes_done) {
    const PointConst p = make_point(d, w, kx, kz, alpha);
    cx m[7];
    if (d.single_size) {
        // average() shortcut: mu * g(taumax)                       (psi_dispersion.py:130-131)
#pragma unroll
        for (int k = 0; k < 7; ++k) m[k] = cx(0.0, 0.0);
        if (p.nu != 0.0) m_integrands_acc<true>(p, d.taumax, d.mu, m);
        else m_integrands_acc<false>(p, d.taumax, d.mu, m);
    } else {
        double poles[2 * kMaxPoles];
        const int np = resonance_poles(d, w.re, kx, poles);
        if (p.nu != 0.0)
            average_warp<W, 7>(t, d, poles, np,
                               [&](double tau, double wt, cx* acc) { m_integrands_acc<true>(p, tau, wt, acc); },
                               m, nodes_done);
        else
            average_warp<W, 7>(t, d, poles, np,
                               [&](double tau, double wt, cx* acc) { m_integrands_acc<false>(p, tau, wt, acc); },
                               m, nodes_done);
    }
    if (m_out) {
#pragma unroll
        for (int k = 0; k < 7; ++k) m_out[k] = m[k];
    }
    return det_P_plus_M(d, p, m);
}

}  // namespace psi
