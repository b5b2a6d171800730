"""Extended-precision (mpmath, 40 digits) evaluation of D(omega) -- an independent yardstick for the cases where
the reference's own double-precision value is only good to ~1e-10 (LU cancellation in det(P+M), rounding noise
~1e-16/tau of its integrands, psi_dispersion.py:212-373, :451).  Test infrastructure only.

Same mathematics as PSIDispersion.calculate (psi_dispersion.py:375-456): M_ij = mu * int F(s) g_ij(taumax s) ds over
[taumin/taumax, 1] in u = ln s, split at the size-density poles and the resonance poles, D = det(P + M); J0, J1 and
the background velocities (:74-84) in the same precision."""
import mpmath as mp

mp.mp.dps = 40


def _powerbump_sigma(p):
    aP, aL, aR, amin, bumpfac, beta = (mp.mpf(p[k]) for k in ("aP", "aL", "aR", "amin", "bumpfac", "beta"))
    aBT = mp.mpf(p["aBT"]) if p.get("aBT") is not None else None            # PowerBumpTail (power_bump.py:193-212)
    width = max(min(abs(aR - aP), abs(aL - aP))/mp.sqrt(mp.log(2)), mp.mpf("0.1")*aP)     # power_bump.py:63-68
    norm_pl = aP**(-beta)
    amp = bumpfac*norm_pl*aL**beta
    curv = -1/width**2

    def sigma(a):
        pl = norm_pl*a**beta
        if aBT is not None and not a > aBT:
            pl = norm_pl*aBT**beta*mp.sqrt(a/aBT)
        bb = amp*mp.e**(curv*(a - aP)**2)
        v = pl
        if a > aL:
            v = max(pl, bb)
        if a > aP:
            v = bb
        return v*a**3
    return sigma


class TruthDispersion:
    def __init__(self, mu, stokes_range, c=20.0, power=3.5, bump=None, poles=(), rhod=None, j01=None):
        """``bump``: dict(amin, aP, aL, aR, bumpfac, beta=-3.5) for a PowerBump size density (its discontinuity goes
        into ``poles``); None: the power law a^(3-power).  ``rhod``: the normalisation the reference computed on the
        host (SizeDensity.rhod comes from scipy.integrate.quad at its default tolerance, sizedensity.py:37-38, and
        is an INPUT of the hot path); None: the exact integral of sigma.  ``j01``: (J0, J1) as the reference's
        __init__ computes them -- it splits those two integrals at the UNSCALED poles (psi_dispersion.py:67,142), so
        a kink of sigma is not at a split point and J0, J1 carry a ~1e-8 quadrature error that every later
        evaluation inherits; None: the exact integrals."""
        self.mu, self.c = mp.mpf(mu), mp.mpf(c)
        self.max_err = 0.0                 # largest relative error estimate of any quadrature so far
        self.t0, self.t1 = mp.mpf(stokes_range[0]), mp.mpf(stokes_range[1])
        self.sd_poles = [mp.mpf(p) for p in poles]
        if bump is None:
            q = 3 - mp.mpf(power)
            self.sigma = lambda a: a**q
        else:
            b = dict(bump)
            b.setdefault("beta", -3.5)
            self.sigma = _powerbump_sigma(b)
            # quadrature aid only: split also at the other switch points of the piecewise sigma
            self.sd_poles += [mp.mpf(b[k]) for k in ("aL", "aP", "aBT") if b.get(k) is not None]
        self.rhod = self._int(lambda a: self.sigma(a), self.t0, self.t1, self.sd_poles, log=True) if rhod is None \
            else mp.mpf(rhod)
        if j01 is None:
            j0 = self._avg(lambda t: 1/(1 + t*t))
            j1 = self._avg(lambda t: t/(1 + t*t))
        else:
            j0, j1 = mp.mpf(j01[0]), mp.mpf(j01[1])
        self.j0, self.j1 = j0, j1
        self.denom = 1/((1 + j0)**2 + j1**2)
        self.vgx = 2*j1*self.denom
        self.vgy = -(1 + j0)*self.denom

    def _int(self, f, a, b, splits, log=False):
        pts = sorted([a] + [p for p in splits if a < p < b] + [b])
        if log:      # integrate in u = ln a
            lp = [mp.log(p) for p in pts]
            val, err = mp.quad(lambda u: f(mp.e**u)*mp.e**u, lp, maxdegree=12, error=True)
        else:
            val, err = mp.quad(f, pts, maxdegree=12, error=True)
        self.max_err = max(self.max_err, float(err)/max(float(abs(val)), 1e-300))
        return val

    def _avg(self, g, extra=()):
        """mu/rhod * int sigma(a) g(a) da over the Stokes range (psi_dispersion.py:124-157)."""
        splits = list(self.sd_poles) + list(extra)
        return self.mu/self.rhod*self._int(lambda a: self.sigma(a)*g(a), self.t0, self.t1, splits, log=True)

    def calculate(self, w, kx, kz, alpha=0.0, return_M=False):
        w = mp.mpc(complex(w))
        kx, kz = mp.mpf(kx), mp.mpf(kz)
        nu = mp.mpf(alpha)*self.c**2
        k2 = kx*kx + kz*kz
        dn, j0, j1, vgx, vgy = self.denom, self.j0, self.j1, self.vgx, self.vgy
        c0 = 1/(w - kx*vgx)
        I = mp.mpc(0, 1)

        def parts(tau):
            ux = 2*dn*(j1 - tau*(1 + j0))/(1 + tau*tau)
            uy = -dn*(1 + j0 + tau*j1)/(1 + tau*tau)
            dux, duy = ux - vgx, uy - vgy
            it = I/tau
            d = kx*ux - w - it
            den = 1/(1 - d*d)
            a11, a12, a21, a33 = -d*den, 2*I*den, -I*den/2, 1/d
            d11, d13 = it*(dux*kx*c0 - 1), it*dux*kz*c0
            d21, d22, d23, d33 = it*duy*kx*c0, -it, it*duy*kz*c0, -it
            ad11, ad12, ad13 = a11*d11 + a12*d21, a12*d22, a11*d13 + a12*d23
            ad21, ad22, ad23 = a21*d11 + a11*d21, a11*d22, a21*d13 + a11*d23
            ad33 = a33*d33
            dtau = (1 + tau + 4*tau*tau)*nu/(1 + tau*tau)**2
            dv = 1/(w - kx*ux + I*(k2*dtau + mp.mpf("1e-10")))
            v11, v13, v21, v23 = 1 + dux*kx*dv, dux*kz*dv, duy*kx*dv, duy*kz*dv
            q = I*dtau*k2*dv*c0
            return [(v11*ad11 + dux*kx*q - 1)*it, (v11*ad12)*it, (v11*ad13 + v13*ad33 + dux*kz*q)*it,
                    (v21*ad11 + ad21 + duy*kx*q)*it, (v21*ad12 + ad22 - 1)*it,
                    (v21*ad13 + ad23 + v23*ad33 + duy*kz*q)*it, (ad33 - 1)*it]

        # resonance poles: real roots of Re(w) - kx*ux(tau) = 0 in the Stokes range (:408-446)
        c2 = w.real/kx
        cc0, cc1 = -2*dn*j1 + c2, 2*dn*(1 + j0)
        extra = []
        if c2 == 0:
            extra = [-cc0/cc1]
        else:
            disc = cc1*cc1 - 4*c2*cc0
            if disc >= 0:
                extra = [(-cc1 + mp.sqrt(disc))/(2*c2), (-cc1 - mp.sqrt(disc))/(2*c2)]
        extra = [r for r in extra if self.t0 < r < self.t1]
        m = [self._avg(lambda t, k=k: parts(t)[k], extra) for k in range(7)]
        r = w - kx*vgx
        d = kx*vgx - w - I*nu*k2
        c = self.c
        P = mp.matrix([[d + (kx*c)**2/r - I*kx**2*nu/3, 2*I, kx*c**2*kz/r - I*kx*kz*nu/3],
                       [-I/2, d, 0],
                       [kz*c**2*kx/r - I*kx*kz*nu/3, 0, d + (kz*c)**2/r - I*kz**2*nu/3]])
        M = mp.matrix([[m[0], m[1], m[2]], [m[3], m[4], m[5]], [0, 0, m[6]]])
        D = mp.det(P + M)
        if return_M:
            return complex(D), [complex(x) for x in m]
        return complex(D)
