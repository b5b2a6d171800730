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
#else
#define PSI_HD inline
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
enum DistKind { kPowerLaw = 0, kPowerBump = 1, kPowerBumpTail = 2 };
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
};

// sigma(a) of the analytic families
PSI_HD double sigma_of(const DistParams& d, double a) {
    if (d.kind == kPowerLaw) return pow(a, d.q);
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

// ---- per-point constants ----------------------------------------------------------------------------
struct PointConst {
    double kx, kz, k2, nu;
    cx w;            // omega
    cx c0;           // 1/(omega - kx*vgx)
    double A2;       // 2*denom
    double j0p;      // 1 + J0
    double j1, dn, vgx, vgy;
};

PSI_HD PointConst make_point(const DistParams& d, cx w, double kx, double kz, double alpha) {
    PointConst p;
    p.kx = kx; p.kz = kz; p.k2 = kx * kx + kz * kz;
    p.nu = alpha * d.c * d.c;                       // psi_dispersion.py:389
    p.w = w;
    p.c0 = recip(cx(w.re - kx * d.vgx, w.im));
    p.A2 = 2 * d.denom; p.j0p = 1 + d.j0; p.j1 = d.j1; p.dn = d.denom; p.vgx = d.vgx; p.vgy = d.vgy;
    return p;
}

// The seven integrands (V A^-1 D + W)_ij * i/tau, ij in {11,12,13,21,22,23,33}, at stopping time tau,
// from shared sub-expressions (psi_dispersion.py:212-373; expansion in SURVEY.md section 8 a6).
// out[k] += wt * integrand_k.
template <bool VISC>
PSI_HD void m_integrands_acc(const PointConst& p, double tau, double wt, cx* acc) {
    const double r1 = 1.0 / (1.0 + tau * tau);
    const double ux = p.A2 * (p.j1 - tau * p.j0p) * r1;            // :80
    const double uy = -p.dn * (p.j0p + tau * p.j1) * r1;           // :81
    const double dux = ux - p.vgx, duy = uy - p.vgy;
    const double it = 1.0 / tau;                                    // K = i/tau
    const double s1 = dux * p.kx, s2 = dux * p.kz, s3 = duy * p.kx, s4 = duy * p.kz;
    // d = kx*ux - w - i/tau                                           (:214)
    const cx d(p.kx * ux - p.w.re, -p.w.im - it);
    // den = 1/(1 - d^2)                                               (:215)
    const cx den = recip(cx(1.0 - (d.re * d.re - d.im * d.im), -2.0 * d.re * d.im));
    const cx dden = d * den;
    const cx a11 = -dden;                       // = a22
    const cx a12(-2.0 * den.im, 2.0 * den.re);  // 2i*den
    const cx a21(0.5 * den.im, -0.5 * den.re);  // -0.5i*den
    const cx rd = recip(d);                     // a33
    // E = (i/tau)*c0 ; D-matrix entries                               (:238-246)
    const cx E(-it * p.c0.im, it * p.c0.re);
    const cx d11(s1 * E.re, s1 * E.im - it);
    const cx d13 = s2 * E, d21 = s3 * E, d23 = s4 * E;
    // A^-1 D                                                          (:249-275)
    const cx ad11 = a11 * d11 + a12 * d21;
    const cx ad12 = (2.0 * it) * den;                    // a12*(-i/tau)
    const cx ad13 = a11 * d13 + a12 * d23;
    const cx ad21 = a21 * d11 + a11 * d21;
    const cx ad22(-it * dden.im, it * dden.re);          // a11*(-i/tau) = (i/tau)*d*den
    const cx ad23 = a21 * d13 + a11 * d23;
    const cx ad33(it * rd.im, -it * rd.re);              // (1/d)*(-i/tau)
    // dv = 1/(w - kx*ux + i*(k2*Dtau + 1e-10))                        (:288-289)
    double dif = 0.0;
    if (VISC) {
        const double r2 = r1 * r1;
        dif = (1.0 + tau + 4.0 * tau * tau) * p.nu * r2;              // :390
    }
    const cx dv = recip(cx(-d.re, p.w.im + (p.k2 * dif + 1.0e-10)));
    const cx v11(1.0 + s1 * dv.re, s1 * dv.im);
    const cx v13 = s2 * dv, v21 = s3 * dv, v23 = s4 * dv;
    cx o0 = v11 * ad11, o1 = v11 * ad12, o2 = v11 * ad13 + v13 * ad33;
    cx o3 = v21 * ad11 + ad21, o4 = v21 * ad12 + ad22, o5 = v21 * ad13 + ad23 + v23 * ad33;
    cx o6 = ad33;
    if (VISC) {
        // W - (-I) part: q = i*Dtau*k2*dv*c0                           (:338-345)
        const cx q = mul_i((dif * p.k2) * (dv * p.c0));
        o0 = o0 + s1 * q; o2 = o2 + s2 * q; o3 = o3 + s3 * q; o5 = o5 + s4 * q;
    }
    o0.re -= 1.0; o4.re -= 1.0; o6.re -= 1.0;           // W11 = W22 = W33 = -1
    // times K = i/tau, times the quadrature weight
    const double g = wt * it;
    acc[0].re -= g * o0.im; acc[0].im += g * o0.re;
    acc[1].re -= g * o1.im; acc[1].im += g * o1.re;
    acc[2].re -= g * o2.im; acc[2].im += g * o2.re;
    acc[3].re -= g * o3.im; acc[3].im += g * o3.re;
    acc[4].re -= g * o4.im; acc[4].im += g * o4.re;
    acc[5].re -= g * o5.im; acc[5].im += g * o5.re;
    acc[6].re -= g * o6.im; acc[6].im += g * o6.re;
}

// J_alpha integrands tau^alpha/(1+tau^2), alpha = 0, 1              (psi_dispersion.py:183-187)
PSI_HD void j_integrands_acc(double tau, double wt, cx* acc) {
    const double r1 = wt / (1.0 + tau * tau);
    acc[0].re += r1;
    acc[1].re += tau * r1;
}

// ---- resonance poles ---------------------------------------------------------------------------------
// Real roots of c0 + c1*tau + c2*tau^2 = 0 strictly inside (taumin, taumax), divided by taumax, merged
// with the size-density poles and sorted (psi_dispersion.py:408-446).  The reference takes the
// eigenvalues of the companion matrix; the cancellation-free closed form below agrees with it to an
// ulp or two.  Returns the number of poles written to out[] (ascending).
PSI_HD int resonance_poles(const DistParams& d, double wre, double kx, double* out) {
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
    for (int i = 0; i < nr; ++i)
        if (r[i] > d.taumin && r[i] < d.taumax && n < 2 * kMaxPoles) out[n++] = r[i] / d.taumax;
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
