"""CPU oracle for the psitools D(omega) / psi_mode / psi_grid_refine hot path.

TEST INFRASTRUCTURE ONLY.  This is a numpy restatement ("port") of the reference algorithm,
used as the checker for the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``psitools_public_b200/`` imports it and there is no CPU fallback in the product.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the unmodified reference (imported
from /root/reference through ``tests/golden/refshim.py``) in the build container and commits its
outputs under ``tests/golden/*.json``; ``tests/test_oracle_golden.py`` checks this file against
those fixtures and against the golden numbers pasted into the reference's own tests
(test_psi_mode.py:34-189, test_psi_mode_mpi.py:38-171, test_complex_roots_mpi.py:32-102).

Differences from the reference that do NOT change results beyond rounding:
  * the seven M_ij integrands (psi_dispersion.py:354-373) are evaluated jointly from shared
    sub-expressions instead of seven separate closure trees; each component still has its own
    tanh-sinh early exit (tanhsinh.py:120-126), so exit levels are the reference's;
  * ``tanhsinh_integrator=None`` (QUADPACK in the reference, psi_dispersion.py:159-181) is served
    by TanhSinh(15, 12); the goldens agree to <= 1.5e-13 relative.
Third-party numerics are called exactly as the reference calls them: numpy.linalg.svd/det,
numpy.polynomial.Polynomial.roots, scipy.linalg.eig, scipy.integrate.quad, scipy.optimize.bisect,
numpy.random (legacy MT19937 global state).
"""
import numpy as np
import scipy.integrate
import scipy.linalg
import scipy.optimize


# --------------------------------------------------------------------------------------
# tanh-sinh quadrature                                   (reference: psitools/tanhsinh.py)
# --------------------------------------------------------------------------------------
class NodeTable:
    """Abscissae/weights at step 2**-max_m.  Follows tanhsinh.py:34-66."""

    def __init__(self, precision_digit=15, max_level=12):
        self.p = precision_digit
        self.max_m = max_level
        self.eps = np.power(10.0, -self.p)
        hmin = np.power(2.0, -self.max_m)
        # tanhsinh.py:46 -- upper estimate of the number of terms
        nmax = int(np.arcsinh((2.0/np.pi)*np.arctanh(1 - self.eps))/hmin) + 10
        jh = hmin*np.arange(nmax)
        s = 0.5*np.pi*np.sinh(jh)
        c = 0.5*np.pi*np.cosh(jh)
        cs = np.cosh(s)
        x = 1.0 - 1.0/(np.exp(s)*cs)                       # tanhsinh.py:54
        w = c/(cs*cs)                                      # tanhsinh.py:55
        keep = np.logical_and(x < 1.0 - self.eps, w > self.eps).nonzero()   # :58-62
        self.xj = x[keep]
        self.wj = c[keep]/(cs[keep]*cs[keep])              # tanhsinh.py:66


def abscissa(x, a, b):
    """tanhsinh.py:68-70 -- operation order matters for the node positions."""
    return 0.5*(x*(b - a) + b + a)


def integrate_joint(fun, a, b, tab, ncomp, tol=None, stats=None):
    """Level-doubling tanh-sinh of ``ncomp`` integrands sharing one evaluation.

    ``fun(u)`` maps a 1-D array of abscissae to an (ncomp, len(u)) array.  Follows
    tanhsinh.py:72-128: nodes whose image reaches b - eps are dropped (:95-98); level 0 uses
    stride 2**max_m with the x = 0 node counted once (:101-104); level m adds the odd multiples
    of stride 2**(max_m - m) (:107-118); a component stops -- WITHOUT adding the last
    increment -- as soon as |increment| < tol (:120-126).
    """
    if tol is None or tol < tab.eps:
        tol = tab.eps
    keep = (abscissa(tab.xj, a, b) < b - tab.eps).nonzero()
    xs = tab.xj[keep]
    ws = tab.wj[keep]
    half = 0.5*(b - a)

    def fx(x):
        return half*np.reshape(fun(abscissa(x, a, b)), (ncomp, -1))

    step = 2**tab.max_m
    res = np.sum(ws[::step]*fx(xs[::step]), axis=1) + \
        np.sum(ws[step::step]*fx(-xs[step::step]), axis=1)
    nev = len(xs[::step]) + len(xs[step::step])
    active = np.ones(ncomp, dtype=bool)
    levels = np.full(ncomp, tab.max_m, dtype=int)
    for m in range(1, tab.max_m + 1):
        h = np.power(2.0, -m)
        step = 2**(tab.max_m - m)
        w = ws[::step][1::2]
        x = xs[::step][1::2]
        dres = -0.5*res + h*np.sum(w*fx(x), axis=1) + h*np.sum(w*fx(-x), axis=1)
        nev += 2*len(x)
        for k in range(ncomp):
            if not active[k]:
                continue
            if np.abs(dres[k]) < tol:
                active[k] = False
                levels[k] = m
            else:
                res[k] = res[k] + dres[k]
        if not active.any():
            break
    if stats is not None:
        stats['nev'] = stats.get('nev', 0) + nev
        stats.setdefault('levels', []).append(levels)
    return res


def integrate(func, a, b, tab, tol=None):
    """Scalar wrapper with the signature of TanhSinh.integrate (tanhsinh.py:72)."""
    return integrate_joint(lambda u: np.asarray(func(u))[None, :], a, b, tab, 1, tol)[0]


# --------------------------------------------------------------------------------------
# dust size distributions    (reference: sizedensity.py, power_bump.py, psi_dispersion.py:57-66)
# --------------------------------------------------------------------------------------
class DustDist:
    """An evaluable size distribution + the PSIDispersion constructor arguments.

    ``F(s)`` is SizeDensity.__call__ (sizedensity.py:41-48): amax*sigma(amax*s)/rho_d.
    ``poles`` is SizeDensity.poles (unscaled, in units of stopping time).
    """

    def __init__(self, kind, mu, stokes_range, c=20.0, single_size=False):
        self.kind = kind
        self.mu = mu
        self.taumin = stokes_range[0]
        self.taumax = stokes_range[1]
        self.c = c
        self.ssf = single_size
        self.poles = []
        self.rhod = None

    # -- constructors ---------------------------------------------------------------
    @classmethod
    def power_law(cls, mu, stokes_range, c=20.0, power=3.5, single_size=False):
        """psi_dispersion.py:57-66 (default MRN size density with closed-form rho_d)."""
        d = cls('powerlaw', mu, stokes_range, c, single_size)
        d.power = power
        d.rhod = (stokes_range[0]**(4.0 - power) - stokes_range[1]**(4.0 - power))/(power - 4.0)
        d.sigma = lambda a: np.power(a, 3 - power)
        return d

    @classmethod
    def power_bump(cls, mu, amin, aP, aL=None, aR=None, bumpfac=2.0, beta=-3.5, epstot=None,
                   aBT=None, c=20.0, with_pole=True):
        """power_bump.py:38-130 (PowerBump) and :133-212 (PowerBumpTail when aBT is given),
        wrapped as in test_psi_mode.py:89-105: SizeDensity(pb.sigma0, [amin, aR]) with
        rho_d from scipy.integrate.quad (sizedensity.py:37-38) and poles=[discontinuity]."""
        aL = 2.0*aP/3.0 if aL is None else aL
        aR = 1.56*aP if aR is None else aR
        d = cls('powerbump' if aBT is None else 'powerbumptail', mu, [amin, aR], c)
        width = np.max((np.min((np.abs(aR - aP), np.abs(aL - aP)))/np.sqrt(np.log(2.0)), 0.1*aP))
        norm_pl = aP**(-beta)
        if aBT is None:
            def plaw(a):
                return norm_pl*a**beta
        else:
            def plaw(a):
                a = np.asarray(a, dtype=float)
                hi = norm_pl*a**beta
                lo = norm_pl*aBT**beta*(a/aBT)**0.5
                return np.where(a > aBT, hi, lo)
        amp = bumpfac*float(plaw(aL))
        curv = -1/width**2

        def bump(a):
            return amp*np.exp(curv*(a - aP)**2)

        def mnn(a):
            a = np.asarray(a, dtype=float)
            pl = np.array(plaw(a), dtype=float)
            bb = bump(a)
            out = np.where(a > aL, np.maximum(pl, bb), pl)
            out = np.where(a > aP, bb, out)
            return out*a**3

        d.params = dict(amin=amin, aP=aP, aL=aL, aR=aR, bumpfac=bumpfac, beta=beta, aBT=aBT,
                        norm_pl=norm_pl, amp=amp, curv=curv)
        # discontinuity between power law and bump (power_bump.py:87-95)
        g = lambda a: float(plaw(a)) - float(bump(a))
        d.discontinuity = aL if g(aL)*g(aP) > 0.0 else scipy.optimize.bisect(g, aL, aP)
        pts = [d.discontinuity] if aBT is None else [aBT, d.discontinuity]
        d.mnorm = scipy.integrate.quad(lambda a: float(mnn(a)), amin, aR, epsrel=1e-15,
                                       limit=100, points=pts)[0]
        if epstot is None:
            d.sigma = mnn
            d.scale = 1.0
        else:
            d.scale = epstot/d.mnorm
            d.sigma = lambda a: mnn(a)/d.mnorm*epstot
        d.rhod = scipy.integrate.quad(lambda a: float(d.sigma(a)), amin, aR)[0]
        if with_pole:
            d.poles = [d.discontinuity]
        return d

    @classmethod
    def from_callable(cls, mu, sigma, size_range, c=20.0, sigma_integral=None, poles=()):
        d = cls('callable', mu, size_range, c)
        d.sigma = sigma
        d.rhod = scipy.integrate.quad(sigma, size_range[0], size_range[1])[0] \
            if sigma_integral is None else sigma_integral
        d.poles = list(poles)
        return d

    def F(self, s):
        return self.taumax*self.sigma(self.taumax*s)/self.rhod


# --------------------------------------------------------------------------------------
# dispersion relation                               (reference: psitools/psi_dispersion.py)
# --------------------------------------------------------------------------------------
class Dispersion:
    """Restatement of PSIDispersion (psi_dispersion.py:30-456) on top of ``integrate_joint``."""

    def __init__(self, dist, tab=None):
        self.dist = dist
        self.tab = NodeTable() if tab is None else tab
        self.n_node_evals = 0
        # psi_dispersion.py:74-84; note the UNSCALED poles of __init__ (:67) are used here
        j0 = np.real(self.average(lambda x: (np.power(x, 0)/(1 + x**2))[None, :], 1,
                                  list(dist.poles))[0])
        j1 = np.real(self.average(lambda x: (np.power(x, 1)/(1 + x**2))[None, :], 1,
                                  list(dist.poles))[0])
        self.j0, self.j1 = j0, j1
        self.denom = 1/((1 + j0)**2 + j1**2)
        self.vgx = 2*j1*self.denom
        self.vgy = -(1 + j0)*self.denom
        self.base_poly = [-2*self.denom*j1, 2*self.denom*(1 + j0), 0]

    def ux(self, x):
        return 2*self.denom*(self.j1 - x*(1 + self.j0))/(1 + x*x)

    def uy(self, x):
        return -self.denom*(1 + self.j0 + x*self.j1)/(1 + x*x)

    def average(self, g, ncomp, poles):
        """psi_dispersion.py:124-157.  ``g(tau)`` -> (ncomp, n); ``poles`` in units of s."""
        d = self.dist
        if d.ssf:
            return d.mu*np.reshape(g(np.asarray([d.taumax], dtype=float)), (ncomp,))
        t = d.taumax
        a = np.log(d.taumin/d.taumax)
        b = 0
        lp = np.log(np.asarray(poles, dtype=float)) if len(poles) else np.asarray([])

        def f(u):
            e = np.exp(u)
            return g(t*e)*e*d.F(e)

        st = {}
        if len(lp) == 0:
            ret = integrate_joint(f, a, b, self.tab, ncomp, stats=st)
        else:
            lo = np.concatenate((np.asarray([a]), lp))
            hi = np.concatenate((lp, np.asarray([b])))
            ret = np.zeros(ncomp, dtype=complex)
            for aa, bb in zip(lo, hi):
                ret = ret + integrate_joint(f, aa, bb, self.tab, ncomp, stats=st)
        self.n_node_evals += st.get('nev', 0)
        return d.mu*ret

    def resonance_poles(self, w, kx):
        """psi_dispersion.py:408-446: size-density poles / taumax plus the real roots of
        Re(w) - kx*ux(tau) = 0 inside (taumin, taumax), sorted, in units of s."""
        d = self.dist
        poles = list(np.array(d.poles)/d.taumax)
        coeff = [self.base_poly[0] + np.real(w)/kx, self.base_poly[1], np.real(w)/kx]
        roots = np.polynomial.Polynomial(coeff).roots()
        for r in roots:
            if np.real(r) > d.taumin and np.real(r) < d.taumax and np.isreal(r):
                poles.append(r/d.taumax)
        poles.sort()
        return poles

    def m_integrands(self, tau, w, kx, kz, nu):
        """The seven (V A^-1 D + W)_ij * i/tau of psi_dispersion.py:212-373, jointly."""
        ux = self.ux(tau)
        uy = self.uy(tau)
        dux = ux - self.vgx
        duy = uy - self.vgy
        itau = 1j/tau
        c0 = 1/(w - kx*self.vgx)
        d = kx*ux - w - itau
        den = 1/(1 - d**2)
        a11 = -d*den
        a12 = 2*1j*den
        a21 = -0.5*1j*den
        a33 = 1/d
        d11 = itau*(dux*kx*c0 - 1)
        d13 = itau*dux*kz*c0
        d21 = itau*duy*kx*c0
        d22 = -itau
        d23 = itau*duy*kz*c0
        d33 = -itau
        ad11 = a11*d11 + a12*d21
        ad12 = a12*d22
        ad13 = a11*d13 + a12*d23
        ad21 = a21*d11 + a11*d21
        ad22 = a11*d22
        ad23 = a21*d13 + a11*d23
        ad33 = a33*d33
        k2 = kx**2 + kz**2
        dtau = (1 + tau + 4*tau*tau)*nu/(1 + tau*tau)**2
        dv = 1/(w - kx*ux + 1j*(k2*dtau + 1.0e-10))
        v11 = 1 + dux*kx*dv
        v13 = dux*kz*dv
        v21 = duy*kx*dv
        v23 = duy*kz*dv
        q = 1j*dtau*k2*dv*c0
        out = np.empty((7, len(tau)), dtype=complex)
        out[0] = (v11*ad11 + dux*kx*q - 1)*itau
        out[1] = (v11*ad12)*itau
        out[2] = (v11*ad13 + v13*ad33 + dux*kz*q)*itau
        out[3] = (v21*ad11 + ad21 + duy*kx*q)*itau
        out[4] = (v21*ad12 + ad22 - 1)*itau
        out[5] = (v21*ad13 + ad23 + v23*ad33 + duy*kz*q)*itau
        out[6] = (ad33 - 1)*itau
        return out

    def matrix_M(self, w, kx, kz, nu):
        poles = self.resonance_poles(w, kx)
        m = self.average(lambda tau: self.m_integrands(tau, w, kx, kz, nu), 7, poles)
        return np.asarray([[m[0], m[1], m[2]], [m[3], m[4], m[5]], [0, 0, m[6]]])

    def matrix_P(self, w, kx, kz, nu):
        """psi_dispersion.py:189-210."""
        c = self.dist.c
        r = w - kx*self.vgx
        d = kx*self.vgx - w - 1j*nu*(kx**2 + kz**2)
        return np.asarray([
            [d + (kx*c)**2/r - 1j*kx**2*nu/3, 2*1j, kx*c**2*kz/r - 1j*kx*kz*nu/3],
            [-0.5*1j, d, 0],
            [kz*c**2*kx/r - 1j*kx*kz*nu/3, 0, d + (kz*c)**2/r - 1j*kz**2*nu/3]])

    def calculate(self, w, kx, kz, viscous_alpha=0, return_M=False):
        """psi_dispersion.py:375-456."""
        nu = viscous_alpha*self.dist.c**2
        w = np.asarray(w)
        shape = w.shape
        flat = np.ravel(w).astype(complex)
        ret = 0.0*flat
        ms = np.zeros((len(flat), 7), dtype=complex)
        for i, wi in enumerate(flat):
            M = self.matrix_M(wi, kx, kz, nu)
            P = self.matrix_P(wi, kx, kz, nu)
            ret[i] = np.linalg.det(P + M)
            ms[i] = [M[0, 0], M[0, 1], M[0, 2], M[1, 0], M[1, 1], M[1, 2], M[2, 2]]
        out = np.squeeze(ret) if w.ndim == 0 else np.reshape(ret, shape)
        if return_M:
            return out, ms
        return out


# --------------------------------------------------------------------------------------
# contour root counting               (reference: complex_roots.py:517-627, ClosedPath)
# --------------------------------------------------------------------------------------
def count_roots(f, z_list, max_iter=100, tol=0.1, max_step=0.1, trace=None):
    """Winding number of f around the closed polygon z_list by adaptive midpoint insertion.

    Edge (i-1, i) is split when |arg f_i - arg f_{i-1}| > tol (angles NOT unwrapped), when it is
    more than twice as long as either neighbour edge, or longer than max_step -- but never once
    it is shorter than 1e-8 (complex_roots.py:556-595).  When a pass adds nothing the winding sum
    must be within 1e-6 of an integer (:616-623).  Raises like the reference.
    Returns (count, points, f_points)."""
    pts = np.asarray(z_list)
    fv = np.asarray(f(pts))
    for _ in range(max_iter):
        prev = np.roll(pts, 1)
        mid = 0.5*(pts + prev)
        da = np.angle(fv) - np.angle(np.roll(fv, 1))
        dz = np.abs(pts - prev)
        jumpy = (dz > 2*np.roll(dz, 1)) | (dz > 2*np.roll(dz, -1))
        want = ((np.abs(da) > tol) | jumpy) | (dz > max_step)
        pick = (want & (dz > 1.0e-8)).nonzero()[0]
        new_z = mid[pick]
        new_f = np.asarray(f(new_z)) if len(pick) else np.zeros(0, dtype=complex)
        if trace is not None:
            trace.append(len(pick))
        if not np.all(np.isfinite(new_f)):
            raise ValueError("Infinite value in f_add")
        pts = np.insert(pts, pick, new_z)
        fv = np.insert(fv, pick, new_f)
        if len(pick) == 0:
            g = fv + 1.0e-30*1j
            a = 0.5*np.sum(np.angle(np.roll(g, -1)/g))/np.pi
            n = np.rint(a)
            if np.abs(a - n) < 1.0e-6:
                return int(n), pts, fv
            raise RuntimeError("No convergence to integer number of roots")
    raise RuntimeError("Maximum number of iterations exceeded")


# --------------------------------------------------------------------------------------
# secant iteration        (third party: scipy 1.18 optimize/_zeros_py.py newton(), secant branch,
#                          as driven by psi_mode.py:438-466 and complex_roots.py:216-233)
# --------------------------------------------------------------------------------------
class SecantResult:
    def __init__(self, root, converged, iterations, function_calls, degenerate=False):
        self.root = root
        self.converged = converged
        self.iterations = iterations
        self.function_calls = function_calls
        self.degenerate = degenerate      # q1 == q0 (scipy emits a RuntimeWarning here)


def secant(f, x0, x1, maxiter=50, xtol=1.48e-8):
    if xtol <= 0:
        raise ValueError("tol too small (%g <= 0)" % xtol)
    p0 = complex(x0)
    p1 = complex(x1)
    q0 = complex(f(p0))
    q1 = complex(f(p1))
    calls = 2
    if abs(q1) < abs(q0):
        p0, p1, q0, q1 = p1, p0, q1, q0
    p = p1
    for itr in range(maxiter):
        if q1 == q0:
            return SecantResult((p1 + p0)/2.0, False, itr + 1, calls, degenerate=(p1 != p0))
        if abs(q1) > abs(q0):
            p = (-q0/q1*p1 + p0)/(1 - q0/q1)
        else:
            p = (-q1/q0*p0 + p1)/(1 - q1/q0)
        if abs(p - p1) <= xtol:
            return SecantResult(p, True, itr + 1, calls)
        p0, q0 = p1, q1
        p1 = p
        q1 = complex(f(p1))
        calls += 1
    return SecantResult(p, False, maxiter, calls)


def offset_start(z0, eps):
    """Second secant point: z0*(1+eps) +/- eps (psi_mode.py:431-433, complex_roots.py:218-222)."""
    z1 = z0*(1 + eps)
    z1 += (eps if z1.real >= 0 else -eps)
    return z1


# --------------------------------------------------------------------------------------
# AAA rational approximation      (reference: complex_roots.py:38-368, RationalApproximation)
# --------------------------------------------------------------------------------------
def unique_within_tol(a, tol=1e-12):
    """complex_roots.py:29-35 (sorts lexicographically by (re, im); keeps first of each run)."""
    if a.size == 0:
        return a
    b = a.copy()
    b.sort()
    d = np.append(True, np.diff(b))
    return b[d > tol]


class Rect:
    """complex_roots.py:391-425."""

    def __init__(self, xr, yr):
        self.xmin, self.xmax = xr[0], xr[1]
        self.ymin, self.ymax = yr[0], yr[1]

    def sample(self, n):
        x = np.random.uniform(self.xmin, self.xmax, n)      # x first, then y (:404-409)
        y = np.random.uniform(self.ymin, self.ymax, n)
        return x + 1j*y

    def is_in(self, z):
        return ((np.real(z) > self.xmin) & (np.real(z) < self.xmax) &
                (np.imag(z) > self.ymin) & (np.imag(z) < self.ymax))


class AAA:
    def __init__(self, domain, tol=1.0e-13, clean_tol=1.0e-10):
        self.domain = domain
        self.tol = tol
        self.clean_tol = clean_tol

    # nodes = samples with mask==1, rest = mask==0, both in sample order
    def _split(self):
        on = self.mask.astype(bool)
        return self.F[~on], self.Z[~on], self.F[on], self.Z[on]

    def _fit(self):
        """Loewner matrix + SVD + residuals (complex_roots.py:56-90)."""
        Fr, Zr, fn, zn = self._split()
        C = 1/(Zr[:, None] - zn[None, :])
        L = np.matmul(np.diag(Fr), C) - np.matmul(C, np.diag(fn))
        _, _, vh = np.linalg.svd(L, compute_uv=True)
        self.weights = vh.T.conjugate()[:, len(fn) - 1]
        num = np.matmul(C, self.weights*fn)
        den = np.matmul(C, self.weights)
        self.R = np.abs(Fr - num/den)
        return np.max(self.R)/np.max(np.abs(self.F))

    def _add_node(self):
        """complex_roots.py:100-124 (LinAlgError branch included)."""
        order = np.argsort(self.R)
        i = np.argmax(self.R)
        free = np.cumsum(1 - self.mask)
        k = np.searchsorted(free, i + 1)
        self.mask[k] = 1
        try:
            return self._fit()
        except np.linalg.LinAlgError:
            self.mask[k] = 0
            free = np.cumsum(1 - self.mask)
            k = np.searchsorted(free, order[-2] + 1)
            self.mask[k] = 1
            return self._fit()

    def calculate(self, F, Z):
        """complex_roots.py:126-176."""
        if np.shape(F) != np.shape(Z):
            raise TypeError('Function samples and sample points need to have the same shape')
        if not np.isfinite(np.sum(F) + np.sum(Z)):
            raise ValueError('Nan or Inf in sample points or function samples')
        self.F = np.asarray(F)
        self.Z = np.asarray(Z)
        self.M = len(F)
        self.mask = np.zeros(self.M)
        self.mask[0] = 1
        self.mask[1] = 1
        self._fit()
        for _ in range(1, int(len(self.F)/2)):
            if self._add_node() < self.tol:
                break
        while self.cleanup() != 0:
            pass
        return {"n_nodes": np.sum(self.mask),
                "max_residual": np.max(self.R)/np.max(np.abs(self.F))}

    def _arrowhead_eig(self, top):
        _, _, _, zn = self._split()
        m = len(zn)
        A = np.diag(np.insert(zn, 0, 0))
        A[0, 1:] = top
        A[1:, 0] = np.ones(m)
        B = np.diag(np.insert(np.ones(m), 0, 0))
        e, _ = scipy.linalg.eig(A, B)
        return e

    def poles_in_domain(self):
        """complex_roots.py:264-284 (drops the FIRST two eigenvalues, unsorted)."""
        e = self._arrowhead_eig(self.weights)[2:]
        return e[np.asarray(self.domain.is_in(e)).nonzero()]

    def cleanup(self):
        """Froissart-doublet removal, complex_roots.py:286-329."""
        poles = self.poles_in_domain()
        L = 1.0e-5
        dz = L*np.exp(-0.25*1j*np.pi + np.arange(4)*0.5*np.pi*1j)
        drop = []
        for p in poles:
            zc = p + dz
            fc = self.evaluate(zc)
            ci = 0.5*(zc[1] - zc[0])*(fc[1] + fc[0]) + \
                0.5*(zc[2] - zc[1])*(fc[2] + fc[1]) + \
                0.5*(zc[3] - zc[2])*(fc[3] + fc[2]) + \
                0.5*(zc[0] - zc[3])*(fc[0] + fc[3])
            if np.abs(ci)/np.max(np.abs(self.F)) < self.clean_tol:
                _, _, _, zn = self._split()
                j = np.argmin(np.abs(p - zn))
                used = np.cumsum(self.mask)
                drop.append(np.searchsorted(used, j + 1))
        for j in drop:
            self.mask[j] = 0
        self._fit()
        return len(drop)

    def evaluate(self, z):
        """Barycentric evaluation, complex_roots.py:331-368."""
        z = np.asarray(z)
        scalar = z.ndim == 0
        zz = np.ravel(z[None] if scalar else z)
        _, _, fn, zn = self._split()
        out = 0.0*zz
        nz = np.asarray(self.weights != 0).nonzero()
        for i in range(len(zz)):
            d = np.copy(self.weights)
            d[nz] = d[nz]/(zz[i] - zn[nz])
            out[i] = np.sum(d*fn)/np.sum(d)
        return np.squeeze(out) if scalar else np.reshape(out, z.shape)

    def find_zeros(self):
        """complex_roots.py:178-240 (method='arrowhead', secant_improve=True)."""
        _, _, fn, _ = self._split()
        e = self._arrowhead_eig(self.weights*fn)
        zeros = np.atleast_1d(np.sort(e)[:-2])
        if len(zeros) != len(self.weights) - 1:
            raise RuntimeError('Number of roots of rational approximation '
                               'not equal to number of nodes used')
        for i in range(len(zeros)):
            z0 = zeros[i]
            z1 = offset_start(z0, 1.0e-4)
            # the reference turns RuntimeWarnings into errors here and keeps the start value
            with np.errstate(divide='raise', over='raise', invalid='raise'):
                try:
                    r = secant(self.evaluate, z0, z1, maxiter=50, xtol=1.48e-8)
                    if not r.degenerate:
                        zeros[i] = r.root
                except (FloatingPointError, ZeroDivisionError, ValueError):
                    pass
        if len(zeros) > 1:
            zeros = unique_within_tol(zeros)
        return zeros


# --------------------------------------------------------------------------------------
# PSIMode                                               (reference: psitools/psi_mode.py)
# --------------------------------------------------------------------------------------
def guess_domain_size(x):
    """psi_mode.py:34-37."""
    return min((1e-3, np.abs(x.imag)*4))


class Mode:
    """Restatement of PSIMode (psi_mode.py:40-491)."""

    def __init__(self, dispersion, real_range, imag_range, n_sample=20, max_zoom_domains=1,
                 tol=1.0e-13, clean_tol=1.0e-4, max_secant_iterations=100):
        self.disp_obj = dispersion
        self.domain = Rect(real_range, imag_range)
        self.n_sample = n_sample
        self.max_zoom_level = max_zoom_domains
        self.max_zoom_domains_per_level = 4
        self.max_secant_iterations = max_secant_iterations
        self.minimum_interpolation_nodes = 4
        self.n_function_call = 0
        self.max_convergence_iterations = 0
        self.ra = AAA(self.domain, tol=tol, clean_tol=clean_tol)

    def _extra_domain(self, size, centre):
        """psi_mode.py:354-397."""
        dm = self.domain
        if np.real(centre) < dm.xmin + 0.5*size[0]:
            centre = dm.xmin + 0.5*size[0] + 1j*np.imag(centre)
        if np.real(centre) > dm.xmax - 0.5*size[0]:
            centre = dm.xmax - 0.5*size[0] + 1j*np.imag(centre)
        if np.imag(centre) < dm.ymin + 0.5*size[1]:
            centre = np.real(centre) + 1j*(dm.ymin + 0.5*size[1])
        if np.imag(centre) > dm.ymax - 0.5*size[1]:
            centre = np.real(centre) + 1j*(dm.ymax - 0.5*size[1])
        box = Rect([np.real(centre) - 0.5*size[0], np.real(centre) + 0.5*size[0]],
                   [np.imag(centre) - 0.5*size[1], np.imag(centre) + 0.5*size[1]])
        z = box.sample(self.n_sample)
        fz = self.disp(z)
        self.z_sample = np.concatenate((self.z_sample, z))
        self.f_sample = np.concatenate((self.f_sample, fz))
        self.n_function_call += self.n_sample

    def _polish(self, starts, max_iter, initial_iterations=5):
        """psi_mode.py:340-352 + 411-491: find_dispersion_roots calls find_root with ONE scalar at a time, so
        the xtol of :438-444 / :460-464 looks only at the current start (or its restart point)."""
        w0 = np.ravel(np.asarray(starts, dtype=complex)).copy()
        ret = 0*w0
        for i in range(len(w0)):
            eps = 1.0e-6
            xtol = np.min(np.hstack([[1.48e-8], np.abs(w0[i:i + 1].imag*1e-5)]))
            z = secant(self.disp, w0[i], offset_start(w0[i], eps), maxiter=initial_iterations,
                       xtol=xtol)
            self.n_function_call += z.function_calls
            if (not z.converged and np.abs(z.root - w0[i])/np.abs(w0[i]) < 1
                    and max_iter > initial_iterations):
                w0[i] = z.root
                xtol = np.min(np.hstack([[1.48e-8], np.abs(w0[i:i + 1].imag*1e-5)]))
                z = secant(self.disp, w0[i], offset_start(w0[i], eps), maxiter=max_iter,
                           xtol=xtol)
                self.n_function_call += z.function_calls
            if z.converged and bool(self.domain.is_in(z.root)):
                self.max_convergence_iterations = max(self.max_convergence_iterations,
                                                      z.iterations)
                ret[i] = z.root
            else:
                ret[i] = -1j
        return ret

    def _polish_in_domain(self, zeros, max_iter):
        """psi_mode.py:340-352."""
        zeros = np.atleast_1d(zeros)
        out = self._polish(zeros, max_iter) if len(zeros) else zeros
        return out[np.asarray(self.domain.is_in(out)).nonzero()]

    def _zeros_in_domain(self):
        z = self.ra.find_zeros()
        return z[np.asarray(self.domain.is_in(z)).nonzero()]

    def calculate(self, kx, kz, viscous_alpha=0, guess_roots=(), count=False):
        """psi_mode.py:106-333."""
        self.disp = lambda z: self.disp_obj.calculate(z, kx, kz, viscous_alpha)
        dm = self.domain
        self.z_sample = dm.sample(self.n_sample)
        self.n_function_call = self.n_sample
        self.f_sample = self.disp(self.z_sample)
        if count:
            corners = [dm.xmin + 1j*dm.ymin, dm.xmax + 1j*dm.ymin,
                       dm.xmax + 1j*dm.ymax, dm.xmin + 1j*dm.ymax]
            self.n_roots, cp, cf = count_roots(self.disp, corners)
            self.z_sample = np.concatenate((self.z_sample, cp))
            self.f_sample = np.concatenate((self.f_sample, cf))
        for centre in guess_roots:
            s = guess_domain_size(centre)
            self._extra_domain([s, s], centre)

        ret = np.zeros(0, dtype=complex)
        for n in range(0, self.max_zoom_level + 1):
            self.ra.calculate(self.f_sample, self.z_sample)
            zeros = self._zeros_in_domain()
            # psi_mode.py:174-217: halve clean_tol until enough nodes and one growing zero
            min_nodes = self.minimum_interpolation_nodes*(1 + n + len(guess_roots))
            keep_tol = self.ra.clean_tol
            while (np.sum(self.ra.mask) < min_nodes or
                   len(zeros[np.asarray(np.imag(zeros) > 0.0).nonzero()]) == 0):
                self.ra.clean_tol = 0.5*self.ra.clean_tol
                if self.ra.clean_tol < self.ra.tol:
                    break
                self.ra.calculate(self.f_sample, self.z_sample)
                zeros = self._zeros_in_domain()
            self.ra.clean_tol = keep_tol

            zeros = self.ra.find_zeros()
            cand = zeros[np.asarray((np.imag(zeros) < dm.ymax) & (np.real(zeros) < dm.xmax) &
                                    (np.real(zeros) > dm.xmin)).nonzero()]
            zoom_dx = np.power(10.0, -1 - 0.5*n)
            if len(cand) > 0:
                cand = cand[np.argsort(np.imag(cand))]
                zoom = [cand[-1]]
                for i in range(1, len(cand)):
                    z = cand[-1 - i]
                    if all(not (np.abs(np.real(z) - np.real(q)) < zoom_dx) for q in zoom):
                        zoom.append(z)
                zoom = zoom[:self.max_zoom_domains_per_level]
                least_damped = zoom[0]
            else:
                least_damped = None
                zoom = cand
            zeros = np.asarray(zeros[np.asarray(dm.is_in(zeros)).nonzero()])
            max_iter = self.max_secant_iterations
            if n < self.max_zoom_level:
                max_iter = self.n_sample
            if len(zeros) == 0 and least_damped is not None:
                zeros = np.asarray([least_damped])
                max_iter = 5
            ret = self._polish_in_domain(zeros, max_iter)
            growing = ret[np.asarray(np.imag(ret) > 0.0).nonzero()]
            if len(growing) > 0 or n == self.max_zoom_level or len(zoom) == 0:
                return ret
            for centre in zoom:
                sy = np.min([0.1, np.abs(np.imag(centre))])
                self._extra_domain([zoom_dx, sy], centre)
        return ret


# --------------------------------------------------------------------------------------
# schedulers / grid refinement   (reference: psi_mode_mpi.py:162-192, complex_roots_mpi.py:160-175,
#                                 psi_grid_refine.py)
# --------------------------------------------------------------------------------------
def dist_from_init(kw, tab=None):
    """Map PSIMode/PSIDispersion constructor kwargs (with an oracle DustDist passed as
    'size_density', or None for the power law) to a Dispersion."""
    sd = kw.get('size_density')
    if sd is None:
        dist = DustDist.power_law(kw['dust_to_gas_ratio'], kw['stokes_range'],
                                  kw.get('sound_speed_over_eta', 20.0),
                                  kw.get('size_distribution_power', 3.5),
                                  kw.get('single_size_flag', False))
    else:
        dist = sd
    return Dispersion(dist, tab)


def run_mode_item(item, tab=None, cache=None):
    """MpiScheduler.runcompute (psi_mode_mpi.py:162-178) for one arg-dict."""
    np.random.seed(item.get('random_seed', 0))
    init = item['__init__']
    disp = dist_from_init(init, tab)
    md = Mode(disp, init['real_range'], init['imag_range'], init.get('n_sample', 20),
              init.get('max_zoom_domains', 1), init.get('tol', 1e-13),
              init.get('clean_tol', 1e-4), init.get('max_secant_iterations', 100))
    c = item['calculate']
    return md.calculate(c['wave_number_x'], c['wave_number_z'], c.get('viscous_alpha', 0),
                        c.get('guess_roots', ()), c.get('count_roots', False)).copy()


def run_count_item(item, tab=None):
    """complex_roots_mpi.MpiScheduler.runcompute (complex_roots_mpi.py:160-175)."""
    disp = dist_from_init(item['PSIDispersion.__init__'], tab)
    d = item['dispersion']
    f = lambda z: disp.calculate(z, d['wave_number_x'], d['wave_number_z'],
                                 d.get('viscous_alpha', 0))
    try:
        return count_roots(f, item['z_list'], **item.get('count_roots', {}))[0]
    except RuntimeError:
        return -666


def spreadgrid(old, const_axis=0):
    """psi_grid_refine.py:36-56: 2n-1 refinement with log10 midpoints."""
    new = np.zeros([2*n - 1 for n in old.shape])
    new[::2, ::2] = old
    if const_axis == 0:
        new[1:-1:2, ::2] = new[:-1:2, ::2]
        new[:, 1:-1:2] = 10**(0.5*(np.log10(new[:, :-2:2]) + np.log10(new[:, 2::2])))
    elif const_axis == 1:
        new[::2, 1:-1:2] = new[::2, :-1:2]
        new[1:-1:2, :] = 10**(0.5*(np.log10(new[:-2:2, :]) + np.log10(new[2::2, :])))
    else:
        raise ValueError('Bad const_axis')
    return new


def prune_eps(values, epsilon=1e-5):
    """psi_grid_refine.py:68-84."""
    if values is None or len(values) <= 1:
        return values
    keep = [values[0]]
    for v in values[1:]:
        if all(not (np.abs(v - u) < epsilon) for u in keep):
            keep.append(v)
    return np.array(keep)
