// psi_tv.cuh -- terminal-velocity PSI mode of a power-law size distribution for ONE wave-number pair: the
// reference's PSIModeTV.find_roots (terminalvelocitysolver.py:195-418) as per-thread device logic, used as a guess
// generator for root searches on (Kx, Kz) maps.
//
// The relation is closed form (complex log / sqrt, :238-263, :300-315); its growing root is followed from the
// single-size limit s = tau_min/tau_max = 1 (a cubic, :272-294) down to the wanted s in steps that grow by sqrt(2)
// after an accepted root and shrink tenfold after a rejected one (:317-366); every root is found by scipy's secant
// rule (scipy.optimize.newton without a derivative: companion start value x0 (1 + 1e-4) +- 1e-4, 50 iterations,
// |p - p1| <= 1.48e-8, "tolerance reached" / "failed to converge" = rejected step).
// PSI_HD like the other headers: the CPU harness runs the same code against the Python class.
#pragma once
#include "psi_math.cuh"

namespace psi {

struct TvParams {
    double fd, tau_max, b;     // dust fraction mu/(1+mu), largest Stokes number, b = -(3 + beta)
    double s_want;             // tau_min / tau_max
    double diff;               // dust diffusion D = alpha (c/eta)^2
    int max_it;                // continuation steps (accepted or rejected) before giving up
};

PSI_HD cx tv_sqrt(cx z) {      // principal branch
    const double r = hypot(z.re, z.im);
    if (r == 0.0) return cx(0.0, z.im);
    const double t = sqrt(0.5 * (r + fabs(z.re)));
    return z.re >= 0.0 ? cx(t, z.im / (2.0 * t)) : cx(fabs(z.im) / (2.0 * t), copysign(t, z.im));
}
PSI_HD cx tv_log(cx z) { return cx(log(hypot(z.re, z.im)), atan2(z.im, z.re)); }
// arctan through logarithms, the reference's branch choice (:234-236)
PSI_HD cx tv_atanc(cx z) {
    const cx iz = mul_i(z);
    return 0.5 * mul_i(tv_log(cx(1.0 - iz.re, -iz.im)) - tv_log(cx(1.0 + iz.re, iz.im)));
}

// F(z)/(1 + z) for b in {0, 1/2, 3/2, 5/2}   (:238-263)
PSI_HD cx tv_F(const TvParams& P, cx z, double s) {
    if (1.0 - s < 1.0e-12) return div_safe(cx(1.0, 0.0), z + 1.0);
    const double b = P.b, p1 = pow(s, 1.0 - b), p2 = pow(s, 2.0 - b), rs = sqrt(s);
    const cx w = ((b - 1.0) * (1.0 - p2) / ((b - 2.0) * (1.0 - p1 + 1.0e-30))) * z;
    const cx sw = tv_sqrt(w);
    const cx arg = div_safe(cx(1.0 - rs, 0.0), sw + tv_sqrt(div_safe(cx(s, 0.0), w)));
    if (b == 0.5) return cx(1.0, 0.0) - (1.0 / (1.0 - rs)) * (sw * tv_atanc(arg));
    if (b == 1.5) return -div_safe(tv_atanc(arg), (1.0 - p1) * sw);
    if (b == 2.5) return div_safe(cx(1.0, 0.0) + div_safe(tv_atanc(arg), (1.0 - p2) * sw), z);
    return cx(1.0, 0.0) - (1.0 / (1.0 - s)) * (w * tv_log(div_safe(w + 1.0, w + s)));       // b == 0
}

PSI_HD bool tv_supported(double b) { return b == 0.0 || b == 0.5 || b == 1.5 || b == 2.5; }

PSI_HD double tv_average_tau(const TvParams& P, double s) {      // (:296-304)
    s = s * s / (s + 1.0e-10);
    const double num = (P.b - 1.0) * (1.0 - pow(s, P.b - 2.0)), den = (P.b - 2.0) * (1.0 - pow(s, P.b - 1.0));
    return s * P.tau_max * num / den;
}

PSI_HD cx tv_z_to_omega(const TvParams& P, cx z, double tau, double kx, double kz) {
    const double fac = 2.0 * (1.0 - P.fd) * tau * kx;
    const cx w = fac * (z + P.fd);
    return cx(w.re, w.im - P.diff * (kx * kx + kz * kz));
}

PSI_HD cx tv_disp(const TvParams& P, cx z, double s, double kx, double kz) {      // (:300-315)
    const double tau = tv_average_tau(P, s), K2 = kx * kx + kz * kz, kz2 = kz * kz;
    const cx omega = tv_z_to_omega(P, z, tau, kx, kz);
    const cx lhs = cx(1.0, 0.0) - (K2 / kz2) * (omega * omega);
    const cx fac = P.fd * (mul_i(((tau * K2) / kz2) * (omega * (z + P.fd))) + 1.0);
    return fac * tv_F(P, z, s) - lhs;
}

// scipy.optimize.newton(f, x0) without derivative.  false = scipy raises RuntimeError (step rejected).
PSI_HD bool tv_secant(const TvParams& P, double s, double kx, double kz, cx x0, cx* root) {
    cx p0 = x0, p1 = (1.0 + 1.0e-4) * x0;
    p1.re += (p1.re > 0.0 || (p1.re == 0.0 && p1.im >= 0.0)) ? 1.0e-4 : -1.0e-4;      // numpy orders complex by (re, im)
    cx q0 = tv_disp(P, p0, s, kx, kz), q1 = tv_disp(P, p1, s, kx, kz);
    if (hypot(q1.re, q1.im) < hypot(q0.re, q0.im)) { cx t = p0; p0 = p1; p1 = t; t = q0; q0 = q1; q1 = t; }
    for (int it = 0; it < 50; ++it) {
        if (q1.re == q0.re && q1.im == q0.im) {
            if (p1.re != p0.re || p1.im != p0.im) return false;         // "Tolerance of ... reached"
            *root = 0.5 * (p1 + p0);
            return true;
        }
        cx p;
        if (hypot(q1.re, q1.im) > hypot(q0.re, q0.im)) {
            const cx r = div_safe(q0, q1);
            p = div_safe(p0 - r * p1, cx(1.0 - r.re, -r.im));
        } else {
            const cx r = div_safe(q1, q0);
            p = div_safe(p1 - r * p0, cx(1.0 - r.re, -r.im));
        }
        if (hypot(p.re - p1.re, p.im - p1.im) <= 1.48e-8) { *root = p; return true; }
        p0 = p1; q0 = q1; p1 = p;
        q1 = tv_disp(P, p1, s, kx, kz);
    }
    return false;                                                       // "Failed to converge"
}

// growing root of the single-size cubic at tau = tau_max, in the z variable (:272-294); the three roots by
// Ehrlich-Aberth (the reference takes the companion matrix's eigenvalues)
PSI_HD cx tv_single_size_z(const TvParams& P, double kx, double kz) {
    const double K2 = kx * kx + kz * kz, fg = 1.0 - P.fd, tm = P.tau_max;
    const cx B(2.0 * fg * fg * kx * tm, P.fd * tm + P.diff * K2);
    const cx C(P.fd * tm * P.diff * K2 - kz * kz / K2, 0.0);
    const cx D(2.0 * fg * (P.fd - fg) * kz * kz * kx * tm / K2, -P.diff * kz * kz);
    const double rad = 1.0 + fmax(hypot(B.re, B.im), fmax(hypot(C.re, C.im), hypot(D.re, D.im)));
    cx x[3] = {cx(0.4 * rad, 0.9 * rad), cx(-0.65 * rad, 0.72 * rad), cx(0.13 * rad, -0.97 * rad)};
    for (int it = 0; it < 200; ++it) {
        double moved = 0.0;
        for (int k = 0; k < 3; ++k) {
            const cx z = x[k];
            const cx pz = ((z + B) * z + C) * z + D, dp = (3.0 * z + 2.0 * B) * z + C;
            const cx newton = div_safe(pz, dp);
            cx rep(0.0, 0.0);
            for (int l = 0; l < 3; ++l)
                if (l != k) rep = rep + div_safe(cx(1.0, 0.0), z - x[l]);
            const cx one_minus = cx(1.0, 0.0) - newton * rep;
            const cx step = div_safe(newton, one_minus);
            if (step.re == step.re && step.im == step.im) {
                x[k] = z - step;
                moved = fmax(moved, hypot(step.re, step.im) / (hypot(z.re, z.im) + 1.0e-300));
            }
        }
        if (moved < 4.0e-16) break;
    }
    cx w = x[0];
    if (x[1].im > w.im) w = x[1];
    if (x[2].im > w.im) w = x[2];
    const double fac = 2.0 * fg * tm * kx;
    return (1.0 / fac) * cx(w.re - P.fd * fac, w.im + P.diff * K2);
}

// PSIModeTV.find_roots for one minimum Stokes number.  Returns 0 (mode found) or 1 (gave up: z = 0).
PSI_HD int tv_find_root(const TvParams& P, double kx, double kz, cx* omega) {
    cx z_start = tv_single_size_z(P, kx, kz), z = z_start;
    double step = 0.1, s_now = 1.0;
    int tries = 0, status = 0;
    while (s_now > P.s_want) {
        if (++tries > P.max_it) { z = cx(0.0, 0.0); status = 1; break; }
        const double s = fmax(s_now - step, P.s_want);
        const bool ok = tv_secant(P, s, kx, kz, z_start, &z);
        if (!ok) z = z_start;
        if (!ok || z.im < 0.0 || z.im > 0.5 || z.im != z.im) {
            step = 0.1 * step;
        } else {
            step = sqrt(2.0) * step;
            s_now = s;
            z_start = z;
        }
    }
    *omega = tv_z_to_omega(P, z, tv_average_tau(P, P.s_want), kx, kz);
    return status;
}

}  // namespace psi
