// psi_math.cuh -- scalar FP64 building blocks of the D(omega) hot path.
//
// Everything here is PSI_HD (host+device) so that the same source can be compiled by g++ into the
// test-only lane-emulation harness (tests/hostsim); the product library only ever runs it on the GPU.
//
// Reference being replaced (file:line relative to the psitools tree):
//   size densities   sizedensity.py:41-48, psi_dispersion.py:57-66, power_bump.py:75-130,193-212
//   background       psi_dispersion.py:74-84
//   M integrands     psi_dispersion.py:212-373
//   P matrix, det    psi_dispersion.py:189-210, :451
//   resonance poles  psi_dispersion.py:408-446
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define PSI_HD __host__ __device__ __forceinline__
#define PSI_NOINLINE __host__ __device__ __noinline__
#else
#define PSI_HD inline
#define PSI_NOINLINE
#endif

namespace psi {

// ---- rounding-exact helpers (no FMA contraction where the reference's operation order matters) ----
PSI_HD double mul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b; return r;
#endif
}
PSI_HD double add_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b; return r;
#endif
}

// ---- complex128 ------------------------------------------------------------------------------------
struct cx {
    double re, im;
    PSI_HD cx() {}
    PSI_HD cx(double r, double i) : re(r), im(i) {}
};
PSI_HD cx operator+(cx a, cx b) { return cx(a.re + b.re, a.im + b.im); }
PSI_HD cx operator-(cx a, cx b) { return cx(a.re - b.re, a.im - b.im); }
PSI_HD cx operator-(cx a) { return cx(-a.re, -a.im); }
PSI_HD cx operator*(cx a, cx b) { return cx(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
PSI_HD cx operator*(double s, cx a) { return cx(s * a.re, s * a.im); }
PSI_HD cx operator*(cx a, double s) { return cx(s * a.re, s * a.im); }
PSI_HD cx operator+(cx a, double s) { return cx(a.re + s, a.im); }
PSI_HD cx operator-(cx a, double s) { return cx(a.re - s, a.im); }
PSI_HD cx mul_i(cx a) { return cx(-a.im, a.re); }                       // i*a
PSI_HD cx conj(cx a) { return cx(a.re, -a.im); }
PSI_HD double norm2(cx a) { return a.re * a.re + a.im * a.im; }
PSI_HD double cabs(cx a) { return sqrt(norm2(a)); }
PSI_HD double cabs1(cx a) { return fabs(a.re) + fabs(a.im); }             // LAPACK izamax measure
PSI_HD cx recip(cx a) { double s = 1.0 / norm2(a); return cx(a.re * s, -a.im * s); }
PSI_HD cx operator/(cx a, cx b) { return a * recip(b); }
// overflow-safe division (Smith), used where magnitudes are not under control (secant, AAA)
PSI_HD cx div_safe(cx a, cx b) {
    if (fabs(b.re) >= fabs(b.im)) {
        double r = b.im / b.re, d = b.re + b.im * r;
        return cx((a.re + a.im * r) / d, (a.im - a.re * r) / d);
    }
    double r = b.re / b.im, d = b.re * r + b.im;
    return cx((a.re * r + a.im) / d, (a.im * r - a.re) / d);
}

// ---- dust distributions ----------------------------------------------------------------------------
enum DistKind { kPowerLaw = 0, kPowerBump = 1, kPowerBumpTail = 2, kTable = 3 };
constexpr int kMaxTableSegs = 5;     // smooth pieces of a tabulated size density (at most kMaxPoles kinks)
constexpr int kMaxPoles = 4;

struct DistParams {
    int kind;
    int single_size;          // PSIDispersion(single_size_flag=True): M = mu*g(taumax), J likewise
    double mu, taumin, taumax, c;
    // size density sigma(a); F(s) = taumax*sigma(taumax*s)/rhod      (sizedensity.py:41-48)
    double rhod;
    double q;                 // power law: sigma = a^q, q = 3 - size_distribution_power
    double beta, norm_pl, amp, curv, aL, aP, aBT, scale;   // power bump (+tail): power_bump.py
    // background state                                      (psi_dispersion.py:74-84)
    double j0, j1, denom, vgx, vgy;
    // split points of the quadrature: SizeDensity.poles as given (used by __init__, :67,142) and
    // divided by taumax (used by calculate, :410)
    int n_init_poles, n_calc_poles;
    double init_poles[kMaxPoles];
    double calc_poles[kMaxPoles];
    // tabulated size density (kind == kTable): w(u) = e^u F(e^u), u = ln s in [ln smin, 0], as piecewise cubics on
    // a uniform grid per smooth piece [brk[k], brk[k+1]) -- cell i of piece k: v(t) = c0 + t (c1 + t (c2 + t c3)),
    // t in [0, 1), four doubles at tab[4 * (tab_off[k] + i)] (32 bytes: two 16-byte loads); tab_log[k]: v = ln w
    const double* tab;
    int tab_nseg;
    int tab_off[kMaxTableSegs], tab_n[kMaxTableSegs], tab_log[kMaxTableSegs];
    double tab_brk[kMaxTableSegs + 1], tab_invh[kMaxTableSegs];
};

// weight e^u F(e^u) of a tabulated size density
PSI_HD double table_weight(const DistParams& d, double u) {
    int k = 0;
    while (k + 1 < d.tab_nseg && u >= d.tab_brk[k + 1]) ++k;
    double t = (u - d.tab_brk[k]) * d.tab_invh[k];
    int i = (int)t;
    if (i < 0) i = 0;
    if (i > d.tab_n[k] - 1) i = d.tab_n[k] - 1;
    t -= (double)i;
    const double* c = d.tab + 4 * ((size_t)d.tab_off[k] + i);
#if defined(__CUDA_ARCH__)
    const double2 c01 = __ldg(reinterpret_cast<const double2*>(c));
    const double2 c23 = __ldg(reinterpret_cast<const double2*>(c) + 1);
    const double v = fma(fma(fma(c23.y, t, c23.x), t, c01.y), t, c01.x);
#else
    const double v = ((c[3] * t + c[2]) * t + c[1]) * t + c[0];
#endif
    return d.tab_log[k] ? exp(v) : v;
}

// ---- fast reciprocal / reciprocal square root (MUFU seed + Newton, ~1 ulp) ---------------------------
PSI_HD double fast_rcp(double a) {
#if defined(__CUDA_ARCH__)
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));     // ~20 bits (MUFU.RCP64H)
    double e = fma(-a, x, 1.0);
    x = fma(x, e, x);
    e = fma(-a, x, 1.0);
    return fma(x, e, x);
#else
    return 1.0 / a;
#endif
}
PSI_HD double fast_rsqrt(double a) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));   // ~20 bits (MUFU.RSQ64H)
    const double ha = 0.5 * a;
    y = y * fma(-ha * y, y, 1.5);
    y = y * fma(-ha * y, y, 1.5);
    return y;
#else
    return 1.0 / sqrt(a);
#endif
}
PSI_HD cx fast_recip(cx a) { const double s = fast_rcp(norm2(a)); return cx(a.re * s, -a.im * s); }

// sigma(a) of the analytic families (straightforward form; the quadrature uses NodeGen below)
PSI_HD double sigma_of(const DistParams& d, double a) {
    if (d.kind == kPowerLaw) return pow(a, d.q);
    if (d.kind == kTable) {          // w(u) = e^u F(e^u) = s taumax sigma(taumax s)/rhod  ->  sigma = w rhod/a
        const double u = log(a / d.taumax);
        return table_weight(d, u) * d.rhod / a;
    }
    // PowerBump.mnn (power_bump.py:97-123) / PowerBumpTail.fnn (:193-212)
    double pl = d.norm_pl * pow(a, d.beta);
    if (d.kind == kPowerBumpTail && !(a > d.aBT))
        pl = d.norm_pl * pow(d.aBT, d.beta) * sqrt(a / d.aBT);
    double t = a - d.aP;
    double bb = d.amp * exp(d.curv * (t * t));
    double v = pl;
    if (a > d.aL) v = fmax(pl, bb);
    if (a > d.aP) v = bb;
    return v * (a * a * a) * d.scale;
}
PSI_HD double F_of(const DistParams& d, double s) {
    return d.taumax * sigma_of(d, d.taumax * s) / d.rhod;
}

// ---- quadrature nodes of one sub-interval ------------------------------------------------------------
// For the abscissa pair +-x of the tanh-sinh rule on [a, b] (in u = ln s, psi_dispersion.py:135-145)
// the quadrature needs, per node:  tau = taumax*e^u,  1/tau,  and the weight e^u * F(e^u).
// With mid = (a+b)/2, half = (b-a)/2:  e^{u(+-x)} = e^{mid} * E^{+-1},  E = exp(half*x), so ONE exp and
// one reciprocal serve both nodes of the pair, and 1/tau comes for free.  For a power law
// e^u F(e^u) = taumax^(q+1)/rhod * e^{(q+1)u}  (psi_dispersion.py:57-66), i.e. a second exp pair --
// or a square root when q + 1 = 1/2 (MRN).  For the bump families e^u F(e^u) = tau^4 v(tau) scale/rhod
// with v the piecewise max(power law, Gaussian) of power_bump.py:97-123.
enum NodeKind { kNodePowerSqrt = 0, kNodePowerGen = 1, kNodeBump = 2, kNodeBumpTail = 3, kNodeTable = 4 };

PSI_HD int node_kind_of(const DistParams& d) {
    if (d.kind == kPowerLaw) return (d.q == -0.5) ? kNodePowerSqrt : kNodePowerGen;
    if (d.kind == kTable) return kNodeTable;
    return d.kind == kPowerBump ? kNodeBump : kNodeBumpTail;
}

struct NodeVals { double tau, it, wt; };      // stopping time, 1/tau, e^u F(e^u)

template <int NK>
struct NodeGen {
    double half, tm, itm, cw, qe, plm, tailc, umid;
    const DistParams* d;
    PSI_HD void begin(const DistParams& dd, double a, double b) {
        d = &dd;
        half = 0.5 * (b - a);
        const double mid = 0.5 * (b + a);
        umid = mid;
        const double em = exp(mid);
        tm = dd.taumax * em;                    // tau at x = 0
        itm = 1.0 / tm;
        if (NK == kNodePowerSqrt || NK == kNodePowerGen) {
            qe = dd.q + 1.0;
            cw = pow(dd.taumax, qe) / dd.rhod * exp(qe * mid);
        } else if (NK == kNodeTable) {
            cw = 1.0;
        } else {
            cw = dd.scale / dd.rhod;
            plm = dd.norm_pl * pow(tm, dd.beta);                         // power law at x = 0
            tailc = dd.norm_pl * pow(dd.aBT > 0 ? dd.aBT : 1.0, dd.beta);  // value at aBT
        }
    }
    PSI_HD double bump_weight(double tau, double pl) const {
        const DistParams& dd = *d;
        if (NK == kNodeBumpTail && !(tau > dd.aBT)) pl = tailc * sqrt(tau * fast_rcp(dd.aBT));
        const double t = tau - dd.aP;
        const double bb = dd.amp * exp(dd.curv * (t * t));
        double v = pl;
        if (tau > dd.aL) v = fmax(pl, bb);
        if (tau > dd.aP) v = bb;
        const double t2 = tau * tau;
        return cw * v * (t2 * t2);
    }
    // nodes at +x (p) and -x (n)
    PSI_HD void pair(double x, NodeVals& p, NodeVals& n) const {
        const double h = half * x;
        if (NK == kNodePowerSqrt) {
            // e^{u/2} weights: exponentiate h/2 once; E = R^2, 1/E = (1/R)^2
            const double R = exp(0.5 * h), Ri = fast_rcp(R);
            const double E = R * R, Ei = Ri * Ri;
            p.tau = tm * E;  p.it = itm * Ei;  p.wt = cw * R;
            n.tau = tm * Ei; n.it = itm * E;   n.wt = cw * Ri;
            return;
        }
        const double E = exp(h), Ei = fast_rcp(E);
        p.tau = tm * E;  p.it = itm * Ei;
        n.tau = tm * Ei; n.it = itm * E;
        if (NK == kNodePowerGen) {
            const double G = exp(qe * h);
            p.wt = cw * G; n.wt = cw * fast_rcp(G);
        } else if (NK == kNodeTable) {
            p.wt = table_weight(*d, umid + h);
            n.wt = table_weight(*d, umid - h);
        } else {
            const double B = exp(d->beta * h);
            p.wt = bump_weight(p.tau, plm * B);
            n.wt = bump_weight(n.tau, plm * fast_rcp(B));
        }
    }
    PSI_HD NodeVals centre() const {            // x = 0
        NodeVals c; c.tau = tm; c.it = itm;
        if (NK == kNodePowerSqrt || NK == kNodePowerGen) c.wt = cw;
        else if (NK == kNodeTable) c.wt = table_weight(*d, umid);
        else c.wt = bump_weight(tm, plm);
        return c;
    }
};

// ---- per-point constants ----------------------------------------------------------------------------
struct PointConst {
    double kx, kz, k2, nu, kappa;   // kappa = kz/kx
    cx w;            // omega
    cx c0;           // 1/(omega - kx*vgx)
    cx ic0;          // i*c0
    double c1;       // Im(w)^2 + 1
    double A2;       // 2*denom
    double j0p;      // 1 + J0
    double j1, dn, vgx, vgy;
    double kxA2, mkxdn, kxvgx, kxvgy;   // kx*2*denom, -kx*denom, kx*vgx, kx*vgy
};

PSI_HD PointConst make_point(const DistParams& d, cx w, double kx, double kz, double alpha) {
    PointConst p;
    p.kx = kx; p.kz = kz; p.k2 = kx * kx + kz * kz; p.kappa = kz / kx;
    p.nu = alpha * d.c * d.c;                       // psi_dispersion.py:389
    p.w = w;
    p.c0 = recip(cx(w.re - kx * d.vgx, w.im));
    p.ic0 = mul_i(p.c0);
    p.c1 = w.im * w.im + 1.0;
    p.A2 = 2 * d.denom; p.j0p = 1 + d.j0; p.j1 = d.j1; p.dn = d.denom; p.vgx = d.vgx; p.vgy = d.vgy;
    p.kxA2 = kx * p.A2; p.mkxdn = -kx * d.denom; p.kxvgx = kx * d.vgx; p.kxvgy = kx * d.vgy;
    return p;
}

// Scale factors deferred out of the node loop: M12 = 2*int(...), M13 = kappa*int(...), M23 likewise.
PSI_HD void m_scales(const PointConst& p, double* sc) {
    sc[0] = 1.0; sc[1] = 2.0; sc[2] = fabs(p.kappa); sc[3] = 1.0; sc[4] = 1.0; sc[5] = fabs(p.kappa); sc[6] = 1.0;
}
PSI_HD void m_apply_scales(const PointConst& p, cx* m) {
    m[1] = 2.0 * m[1]; m[2] = p.kappa * m[2]; m[5] = p.kappa * m[5];
}

// The seven integrands (V A^-1 D + W)_ij * i/tau, ij in {11,12,13,21,22,23,33}, at one node
// (psi_dispersion.py:212-373).  Notation: delta = kx*ux - w, d = delta - i/tau, den = 1/(1-d^2),
// rd = 1/d, dv = 1/(w - kx*ux + i(k^2 Dtau + 1e-10)), s1 = dux*kx, s3 = duy*kx, kappa = kz/kx.
// Algebra used (all exact identities of the reference's expressions):
//   A^-1 D columns 1 and 3 are E*(p, r) - (i/tau)(a11, a21) and kappa*E*(p, r) with
//     E = (i/tau) c0,  p = den*(-d s1 + 2i s3),  r = den*(-d s3 - i s1/2);
//   the "-1" of W11, W22, W33 is absorbed without cancellation:
//     -(i/tau) a11 - 1 = (delta^2 - (i/tau) delta - 1) den =: N den,   A33D33 - 1 = -delta*rd;
//   v11 = 1 + s1 dv, v21 = s3 dv, v13 = kappa s1 dv, v23 = kappa s3 dv.
// acc[k] += wt * (i/tau) * integrand_k, with the constant factors of m_scales() left out.
template <bool VISC>
PSI_HD void m_node(const PointConst& p, const NodeVals& nv, double wq, cx* acc) {
    const double tau = nv.tau, it = nv.it;
    const double t1 = fma(tau, tau, 1.0);                          // 1 + tau^2
    // kx*ux, kx*uy from :80-81 with the common factor kx/(1+tau^2) applied once
    const double nx = fma(-tau, p.j0p, p.j1);                      // j1 - tau*(1+j0)
    const double ny = fma(tau, p.j1, p.j0p);                       // (1+j0) + tau*j1
    const double dr_num = p.kxA2 * nx;                             // kx*ux*(1+tau^2)
    // d = kx*ux - w - i/tau                                         (:214);  dr = Re d, di = Im d
    const double di = -p.w.im - it;
    // the four real denominators are inverted together (one MUFU seed + Newton, products back out):
    //   n1 = 1+tau^2,  n2 = |1-d^2|^2,  n3 = |d|^2,  n4 = |w - kx*ux + i(k^2 Dtau + 1e-10)|^2
    const double r1 = fast_rcp(t1);
    const double kxux = dr_num * r1;
    const double s1 = kxux - p.kxvgx;                              // (ux - vgx)*kx
    const double s3 = fma(p.mkxdn * ny, r1, -p.kxvgy);             // (uy - vgy)*kx
    const double dr = kxux - p.w.re;
    const double dr2 = dr * dr;
    const cx one_md2(fma(di, di, 1.0 - dr2), -2.0 * dr * di);      // 1 - d^2
    double dif = 0.0;
    if (VISC) dif = (1.0 + tau + 4.0 * tau * tau) * p.nu * (r1 * r1);   // :390
    const double yi = p.w.im + (p.k2 * dif + 1.0e-10);             // Im(w - kx*ux + i(...)); Re = -dr
    const double n2 = norm2(one_md2), n3 = fma(di, di, dr2), n4 = fma(yi, yi, dr2);
    const double p23 = n2 * n3;
    const double inv = fast_rcp(p23 * n4);
    const double i4 = inv * p23, i23 = inv * n4;
    const double i2 = i23 * n3, i3 = i23 * n2;
    const cx den(one_md2.re * i2, -one_md2.im * i2);               // 1/(1-d^2)   (:215)
    const cx rd(dr * i3, -di * i3);                                // 1/d
    const cx dv(-dr * i4, -yi * i4);                               // (:288-289)
    // N = delta^2 - (i/tau) delta - 1, delta = (dr, -w.im)
    const cx N(dr2 - p.c1 - it * p.w.im, -dr * (2.0 * p.w.im + it));
    const cx Nden = N * den;
    const cx P0(-dr * s1, 2.0 * s3 - di * s1);
    const cx R0(-dr * s3, -di * s3 - 0.5 * s1);
    const cx cden = p.ic0 * den;
    const cx Ep = cden * P0, Er = cden * R0;           // E p / (1/tau), E r / (1/tau)
    const cx X(fma(it, Ep.re, Nden.re), fma(it, Ep.im, Nden.im));         // A11D11... - 1
    const cx T1 = dv * cx(X.re + 1.0, X.im);
    const cx Y = dv * den;
    const cx S = dv * cx(Ep.re + rd.im, Ep.im - rd.re);                    // dv*(Ep - i rd)
    cx o0(fma(s1, T1.re, X.re), fma(s1, T1.im, X.im));
    const cx o1(fma(s1, Y.re, den.re), fma(s1, Y.im, den.im));            // * it (folded into g1)
    cx o2(fma(s1, S.re, Ep.re), fma(s1, S.im, Ep.im));                    // * it
    cx o3(fma(it, fma(-0.5, den.re, Er.re), s3 * T1.re), fma(it, fma(-0.5, den.im, Er.im), s3 * T1.im));
    const double its3 = 2.0 * it * s3;
    const cx o4(fma(its3, Y.re, Nden.re), fma(its3, Y.im, Nden.im));
    cx o5(fma(s3, S.re, Er.re), fma(s3, S.im, Er.im));                    // * it
    const cx o6(-(dr * rd.re + p.w.im * rd.im), -(dr * rd.im - p.w.im * rd.re));   // -delta*rd
    const double g = wq * nv.wt * it;
    const double g1 = g * it;
    if (VISC) {
        // W + I: q = i*Dtau*k2*dv*c0                                       (:338-345)
        const cx q = (dif * p.k2) * (dv * p.ic0);
        o0 = o0 + s1 * q; o3 = o3 + s3 * q;
        // columns 3 carry kz = kappa*kx; with kappa deferred the W13/W23 terms are s1 q, s3 q and are
        // NOT multiplied by 1/tau: accumulate them with g instead of g1
        acc[2].re = fma(-g, s1 * q.im, acc[2].re); acc[2].im = fma(g, s1 * q.re, acc[2].im);
        acc[5].re = fma(-g, s3 * q.im, acc[5].re); acc[5].im = fma(g, s3 * q.re, acc[5].im);
    }
    acc[0].re = fma(-g, o0.im, acc[0].re); acc[0].im = fma(g, o0.re, acc[0].im);
    acc[1].re = fma(-g1, o1.im, acc[1].re); acc[1].im = fma(g1, o1.re, acc[1].im);
    acc[2].re = fma(-g1, o2.im, acc[2].re); acc[2].im = fma(g1, o2.re, acc[2].im);
    acc[3].re = fma(-g, o3.im, acc[3].re); acc[3].im = fma(g, o3.re, acc[3].im);
    acc[4].re = fma(-g, o4.im, acc[4].re); acc[4].im = fma(g, o4.re, acc[4].im);
    acc[5].re = fma(-g1, o5.im, acc[5].re); acc[5].im = fma(g1, o5.re, acc[5].im);
    acc[6].re = fma(-g, o6.im, acc[6].re); acc[6].im = fma(g, o6.re, acc[6].im);
}

template <bool VISC>
struct MNode {
    const PointConst* p;
    PSI_HD void operator()(const NodeVals& nv, double wq, cx* acc) const { m_node<VISC>(*p, nv, wq, acc); }
};

// J_alpha integrands tau^alpha/(1+tau^2), alpha = 0, 1              (psi_dispersion.py:183-187)
struct JNode {
    PSI_HD void operator()(const NodeVals& nv, double wq, cx* acc) const {
        const double r1 = wq * nv.wt / fma(nv.tau, nv.tau, 1.0);
        acc[0].re += r1;
        acc[1].re = fma(nv.tau, r1, acc[1].re);
    }
};

// ---- resonance poles ---------------------------------------------------------------------------------
// Real roots of c0 + c1*tau + c2*tau^2 = 0 strictly inside (taumin, taumax), divided by taumax, merged
// with the size-density poles and sorted (psi_dispersion.py:408-446).  The reference takes the
// eigenvalues of the companion matrix; the cancellation-free closed form below agrees with it to an
// ulp or two.  Returns the number of poles written to out[] (ascending).
// res_tau (may be null) receives the stopping times of the resonance poles that were added, *n_res their number.
PSI_HD int resonance_poles(const DistParams& d, double wre, double kx, double* out, double* res_tau = nullptr,
                           int* n_res = nullptr) {
    int n = 0;
    for (int i = 0; i < d.n_calc_poles; ++i) out[n++] = d.calc_poles[i];
    const double c2 = wre / kx;
    const double c0 = -2.0 * d.denom * d.j1 + c2;       // base_poly[0] + Re(w)/kx
    const double c1 = 2.0 * d.denom * (1.0 + d.j0);     // base_poly[1]
    double r[2]; int nr = 0;
    if (c2 == 0.0) {                                    // trailing zero trimmed -> linear
        if (c1 != 0.0) r[nr++] = -c0 / c1;
    } else {
        const double disc = c1 * c1 - 4.0 * c2 * c0;
        if (disc >= 0.0) {
            const double sq = sqrt(disc);
            const double qq = -0.5 * (c1 + (c1 >= 0.0 ? sq : -sq));
            r[nr++] = qq / c2;
            if (qq != 0.0) r[nr++] = c0 / qq; else r[nr++] = 0.0;
        }
    }
    int nres = 0;
    for (int i = 0; i < nr; ++i)
        if (r[i] > d.taumin && r[i] < d.taumax && n < 2 * kMaxPoles) {
            out[n++] = r[i] / d.taumax;
            if (res_tau) res_tau[nres] = r[i];
            ++nres;
        }
    if (n_res) *n_res = nres;
    for (int i = 1; i < n; ++i) {                       // insertion sort, n <= 6
        double v = out[i]; int j = i - 1;
        while (j >= 0 && out[j] > v) { out[j + 1] = out[j]; --j; }
        out[j + 1] = v;
    }
    return n;
}

// ---- P matrix and the determinant --------------------------------------------------------------------
// D(omega) = det(P + M), psi_dispersion.py:189-210 and :451; m[] = {M11,M12,M13,M21,M22,M23,M33},
// M31 = M32 = 0.  With g = c^2/(w - kx*vgx) the entries P_xx, P_xz, P_zx, P_zz carry kx^2 g, kx kz g,
// kx kz g, kz^2 g, whose products cancel EXACTLY in the xz-minor:
//     (d+e00+kx^2 g)(d+e22+kz^2 g) - (kx kz g+e02)(kx kz g+e20)
//       = (d+e00)(d+e22) - e02 e20 + g*(kx^2 (d+e22) + kz^2 (d+e00) - kx kz (e02+e20)).
// LAPACK's LU (numpy.linalg.det in the reference) forms the (kx kz g)^2 ~ 1e15 products and loses
// up to 7 digits to that cancellation at large K; expanding it analytically keeps D at full FP64
// accuracy.  The difference to the reference's own value is the reference's LU rounding noise
// (quantified in tests/test_disp_parity.py against an extended-precision determinant).
PSI_HD cx det_P_plus_M(const DistParams& dd, const PointConst& p, const cx* m) {
    const double kx = p.kx, kz = p.kz, nu = p.nu, c = dd.c;
    const cx g = (c * c) * recip(cx(p.w.re - kx * p.vgx, p.w.im));   // c^2/(w - kx*vgx)
    const cx d(kx * p.vgx - p.w.re, -p.w.im - nu * p.k2);            // :192
    const double third = nu / 3.0;
    const cx e00(m[0].re, m[0].im - kx * kx * third);
    const cx e02(m[2].re, m[2].im - kx * kz * third);
    const cx e20(0.0, -kx * kz * third);
    const cx e22(m[6].re, m[6].im - kz * kz * third);
    const cx A = d + e00, B = d + e22;
    const cx minor_xz = A * B - e02 * e20 + g * ((kx * kx) * B + (kz * kz) * A - (kx * kz) * (e02 + e20));
    const cx a01(m[1].re, m[1].im + 2.0);
    const cx a10(m[3].re, m[3].im - 0.5);
    const cx a11 = d + m[4];
    const cx a12 = m[5];
    const cx a20 = (kx * kz) * g + e20;
    const cx a22 = (kz * kz) * g + B;
    return a11 * minor_xz - a01 * (a10 * a22 - a12 * a20);
}

}  // namespace psi
