"""Rational approximation (AAA), rectangular domains and the argument-principle root counter --
drop-in for the hot-path part of psitools.complex_roots (reference complex_roots.py:29-425, 517-627).

What runs where: every evaluation of the dispersion relation requested by these classes is a batched
GPU call (kernel K1).  The control logic is written as *generators* that yield batches of points and
receive the function values, so that ``_engine.run_lockstep`` can advance thousands of contours /
root searches together with one kernel launch per pass and no per-point host round trip.  The plain
classes below (``ClosedPath``, ``RationalApproximation``) keep the reference's call signatures and
drive the same generators with a user supplied callable.

Out of scope (never called by the hot path): Circle, RootFollower,
find_zeros(method='polyroots').
"""
import numpy as np
import scipy.linalg


def unique_within_tol(a, tol=1e-12):
    """Sorted copy of ``a`` without entries closer than ``tol`` to their predecessor."""
    if a.size == 0:
        return a
    srt = np.sort(a)
    gaps = np.append(True, np.diff(srt))
    return srt[gaps > tol]


class Rectangle:
    """Rectangular domain (x_range, y_range) in the complex plane."""

    def __init__(self, x_range, y_range):
        self.xmin, self.xmax = x_range[0], x_range[1]
        self.ymin, self.ymax = y_range[0], y_range[1]

    def generate_random_sample_points(self, n_sample, rng=None):
        """n_sample uniform points: all real parts are drawn first, then all imaginary parts
        (the order fixes the random stream).  ``rng`` defaults to numpy's global legacy state."""
        rng = np.random if rng is None else rng
        re = rng.uniform(self.xmin, self.xmax, n_sample)
        im = rng.uniform(self.ymin, self.ymax, n_sample)
        return re + 1j*im

    def is_in(self, z):
        """Strictly inside?"""
        z = np.asarray(z)
        return ((z.real > self.xmin) & (z.real < self.xmax) &
                (z.imag > self.ymin) & (z.imag < self.ymax))

    def corners(self):
        """Counter-clockwise corner list, as used for root counting."""
        return [self.xmin + 1j*self.ymin, self.xmax + 1j*self.ymin,
                self.xmax + 1j*self.ymax, self.xmin + 1j*self.ymax]

    def grid_xy(self, N=100):
        return np.linspace(self.xmin, self.xmax, N), np.linspace(self.ymin, self.ymax, N)


# ------------------------------------------------------------------------------------------------
# secant iteration with scipy's update and termination rules (scipy.optimize.newton, secant branch,
# which the reference reaches through root_scalar(method='secant'))
# ------------------------------------------------------------------------------------------------
class SecantState:
    __slots__ = ("root", "converged", "iterations", "function_calls", "flagged")


def second_point(z0, eps):
    """Companion start value z0*(1+eps) +- eps (sign of the real part)."""
    z1 = z0*(1 + eps)
    return z1 + (eps if z1.real >= 0 else -eps)


def secant_machine(x0, x1, maxiter, xtol):
    """Generator: yields 1-element arrays of points, receives f there, returns a SecantState."""
    if xtol <= 0:
        raise ValueError("tol too small (%g <= 0)" % xtol)
    out = SecantState()
    out.flagged = False
    p0, p1 = complex(x0), complex(x1)
    q = yield np.array([p0, p1])
    q0, q1 = complex(q[0]), complex(q[1])
    calls = 2
    if abs(q1) < abs(q0):
        p0, p1, q0, q1 = p1, p0, q1, q0
    p = p1
    for it in range(maxiter):
        if q1 == q0:
            out.root, out.converged, out.iterations, out.function_calls = (p1 + p0)/2.0, False, it + 1, calls
            out.flagged = p1 != p0
            return out
        if abs(q1) > abs(q0):
            ratio = q0/q1
            p = (-ratio*p1 + p0)/(1 - ratio)
        else:
            ratio = q1/q0
            p = (-ratio*p0 + p1)/(1 - ratio)
        if abs(p - p1) <= xtol:
            out.root, out.converged, out.iterations, out.function_calls = p, True, it + 1, calls
            return out
        p0, q0, p1 = p1, q1, p
        q1 = complex((yield np.array([p1]))[0])
        calls += 1
    out.root, out.converged, out.iterations, out.function_calls = p, False, maxiter, calls
    return out


def drive(machine, func):
    """Run a generator-machine to completion against a plain callable."""
    try:
        req = next(machine)
        while True:
            req = machine.send(np.asarray(func(req)))
    except StopIteration as stop:
        return stop.value


class RootFollower:
    """Follow a root of func(z, k) while the parameter k moves away from k_start, where func(z_start, k_start) = 0
    (reference complex_roots.py:428-515).  ``calculate(k_want)`` returns the root at every wanted k (sorted
    ascending; values below and above k_start are marched separately, nearest first), 0 where the march gave up.

    The march, as the reference defines it: try the whole remaining distance; a step is accepted when scipy's
    secant rule (10 iterations, started from z_start and its companion point -- the start values are NOT moved
    along with k) converges to a root on the same side of the real axis as z_start; a rejected step is cut
    tenfold, an accepted one is followed by a step sqrt(2) times longer, clipped to the target.  Two properties of
    the reference are kept because results depend on them: (1) once a clipped step would reach the target the
    march stops WITHOUT solving there, so the value reported for a target that needed more than one step is the
    root at the last intermediate k; (2) the give-up test ``k_step < 1e-6`` is signed, so marching towards
    smaller k gives up at the first rejected step."""

    def __init__(self, func, z_start, k_start, verbose=False):
        self.func = func
        self.z_start = z_start
        self.k_start = k_start
        self.verbose = verbose

    def calculate(self, k_want):
        k_want = np.atleast_1d(k_want)
        ret = np.zeros(len(k_want), dtype=np.complex128)
        for side in (k_want < self.k_start, k_want > self.k_start):
            sel = np.asarray(side).nonzero()
            ret[sel] = self._calculate(k_want[sel])
        return ret

    def _solve(self, k):
        z0 = self.z_start
        return drive(secant_machine(z0, second_point(z0, 1.0e-4), 10, 1.48e-8),
                     lambda pts: [self.func(z, k) for z in pts])

    def _calculate(self, k_want):
        ret = np.zeros(len(k_want), dtype=np.complex128)
        if len(k_want) == 0:
            return ret
        downwards = k_want[0] < self.k_start
        order = range(len(k_want) - 1, -1, -1) if downwards else range(len(k_want))
        k = self.k_start
        for idx in order:
            target = k_want[idx]
            step = target - k
            arrived = False
            while not arrived:
                while True:
                    if self.verbose:
                        print("Trying to find root at k = ", k + step, step)
                    res = self._solve(k + step)
                    if res.converged and res.root.imag*self.z_start.imag > 0:
                        break
                    step = step/10
                    if step < 1.0e-6:
                        return ret
                k = k + step
                step = step*np.sqrt(2)
                if (step > 0 and k + step > target) or (step <= 0 and k + step < target) or step == 0:
                    step = target - k          # (step == 0: a target equal to the current k; the reference
                    arrived = True             # loops forever there)
            ret[idx] = res.root
        return ret


# ------------------------------------------------------------------------------------------------
# AAA
# ------------------------------------------------------------------------------------------------
class RationalApproximation:
    """AAA rational approximation (Nakatsukasa, Sete & Trefethen 2018) of samples F at points Z.

    Args:
        domain: domain object with ``is_in`` (poles outside it are never cleaned up)
        tol: residual tolerance relative to max|F| (default 1e-13)
        clean_tol: Froissart-doublet threshold (default 1e-10)

    Attributes kept for callers: ``F, Z, f, z, M, maskF`` (1 = support point), ``weights``, ``R``.
    """

    def __init__(self, domain, tol=1.0e-13, clean_tol=1.0e-10):
        self.domain = domain
        self.tol = tol
        self.clean_tol = clean_tol
        self._greedy_cache = None

    # -- pieces ---------------------------------------------------------------------------
    def _nodes(self):
        on = self.maskF.astype(bool)
        return self.f[on], self.z[on]

    def calc_weights_residuals(self):
        """Weights = right singular vector of the Loewner matrix for its smallest singular value;
        residuals at the points not used as support points.  Returns max residual / max|F|."""
        on = self.maskF.astype(bool)
        Fr, Zr, fn, zn = self.F[~on], self.Z[~on], self.f[on], self.z[on]
        cauchy = 1/(Zr[:, None] - zn[None, :])
        loewner = Fr[:, None]*cauchy - cauchy*fn[None, :]
        vh = np.linalg.svd(loewner, compute_uv=True)[2]
        self.weights = vh.T.conjugate()[:, len(fn) - 1]
        num = cauchy @ (self.weights*fn)
        den = cauchy @ self.weights
        self.R = np.abs(Fr - num/den)
        return np.max(self.R)/np.max(np.abs(self.F))

    def _free_index(self, k):
        """Position in the sample arrays of the k-th point that is not a support point."""
        return int(np.searchsorted(np.cumsum(1 - self.maskF), k + 1))

    def step(self):
        """Promote the worst-fitted sample to a support point and refit."""
        ranking = np.argsort(self.R)
        pos = self._free_index(int(np.argmax(self.R)))
        self.maskF[pos] = 1
        try:
            return self.calc_weights_residuals()
        except np.linalg.LinAlgError:
            self.maskF[pos] = 0
            pos = self._free_index(int(ranking[-2]))
            self.maskF[pos] = 1
            return self.calc_weights_residuals()

    def calculate(self, F, Z):
        """Greedy AAA to ``tol`` with at most len(F)/2 support points, then doublet cleanup."""
        if np.shape(F) != np.shape(Z):
            raise TypeError('Function samples and sample points need to have the same shape')
        if not np.isfinite(np.sum(F) + np.sum(Z)):
            raise ValueError('Nan or Inf in sample points or function samples')
        F = np.asarray(F)
        Z = np.asarray(Z)
        self.F = self.f = F
        self.Z = self.z = Z
        self.M = len(F)
        cache = self._greedy_cache
        if cache is not None and cache[0].shape == F.shape and np.array_equal(cache[0], F) \
                and np.array_equal(cache[1], Z) and cache[2] == self.tol:
            # the greedy phase depends on (F, Z, tol) only: PSIMode's clean_tol halving loop re-runs
            # calculate() on identical samples dozens of times, so its outcome is replayed
            self.maskF, self.weights, self.R = cache[3].copy(), cache[4].copy(), cache[5].copy()
        else:
            self.maskF = np.zeros(self.M)
            self.maskF[0] = 1
            self.maskF[1] = 1
            self.calc_weights_residuals()
            for _ in range(1, int(len(self.F)/2)):
                if self.step() < self.tol:
                    break
            self._greedy_cache = (F.copy(), Z.copy(), self.tol, self.maskF.copy(),
                                  self.weights.copy(), self.R.copy())
        while self.cleanup() != 0:
            pass
        return {"n_nodes": np.sum(self.maskF),
                "max_residual": np.max(self.R)/np.max(np.abs(self.F))}

    def _arrowhead(self, first_row):
        _, zn = self._nodes()
        m = len(zn)
        A = np.diag(np.insert(zn, 0, 0))
        A[0, 1:] = first_row
        A[1:, 0] = np.ones(m)
        B = np.diag(np.insert(np.ones(m), 0, 0))
        return scipy.linalg.eig(A, B)[0]

    def find_poles(self):
        return self._arrowhead(self.weights)[2:]

    def find_poles_in_domain(self):
        e = self.find_poles()
        return e[np.asarray(self.domain.is_in(e)).nonzero()]

    def cleanup(self):
        """Remove the support point nearest to every pole (inside the domain) whose residue,
        measured by a 4-point contour integral of side 1e-5, is below clean_tol*max|F|.
        Returns the number of support points removed."""
        side = 1.0e-5
        ring = side*np.exp(-0.25j*np.pi + 0.5j*np.pi*np.arange(4))
        fmax = np.max(np.abs(self.F))
        doomed = []
        for pole in self.find_poles_in_domain():
            zc = pole + ring
            fc = self.evaluate(zc)
            integral = sum(0.5*(zc[(k + 1) % 4] - zc[k])*(fc[(k + 1) % 4] + fc[k]) for k in range(4))
            if np.abs(integral)/fmax < self.clean_tol:
                _, zn = self._nodes()
                nearest = int(np.argmin(np.abs(pole - zn)))
                doomed.append(int(np.searchsorted(np.cumsum(self.maskF), nearest + 1)))
        for pos in doomed:
            self.maskF[pos] = 0
        self.calc_weights_residuals()
        return len(doomed)

    def evaluate(self, z):
        """Barycentric evaluation at scalar or array z."""
        z = np.asarray(z)
        scalar = z.ndim == 0
        pts = np.ravel(z[None] if scalar else z)
        fn, zn = self._nodes()
        live = np.asarray(self.weights != 0).nonzero()
        out = 0.0*pts
        for i in range(len(pts)):
            terms = np.copy(self.weights)
            terms[live] = terms[live]/(pts[i] - zn[live])
            out[i] = np.sum(terms*fn)/np.sum(terms)
        return np.squeeze(out) if scalar else np.reshape(out, z.shape)

    def find_zeros(self, method='arrowhead', secant_improve=True):
        """Zeros of the approximant: generalised arrowhead eigenvalues, each polished by a secant
        iteration on the approximant itself, duplicates (within 1e-12) merged."""
        if method != 'arrowhead':
            raise NotImplementedError("only method='arrowhead' is on the hot path")
        fn, _ = self._nodes()
        zeros = np.atleast_1d(np.sort(self._arrowhead(self.weights*fn))[:-2])
        if len(zeros) != len(self.weights) - 1:
            raise RuntimeError('Number of roots of rational approximation '
                               'not equal to number of nodes used')
        if secant_improve:
            for i in range(len(zeros)):
                start = zeros[i]
                # the reference promotes RuntimeWarnings to errors here and then keeps the start
                with np.errstate(divide='raise', over='raise', invalid='raise'):
                    try:
                        res = drive(secant_machine(start, second_point(start, 1.0e-4), 50, 1.48e-8),
                                    self.evaluate)
                        if not res.flagged:
                            zeros[i] = res.root
                    except (FloatingPointError, ZeroDivisionError, ValueError):
                        pass
            if len(zeros) > 1:
                zeros = unique_within_tol(zeros)
        return zeros


# ------------------------------------------------------------------------------------------------
# argument principle
# ------------------------------------------------------------------------------------------------
def contour_machine(z_list, max_iter=100, tol=0.1, max_step=0.1, state=None):
    """Generator form of ClosedPath.count_roots: yields batches of contour points, receives f,
    returns the winding number.  ``state`` (optional dict) receives 'points' and 'f_points'.

    An edge (previous point -> point) is bisected when the argument of f changes by more than
    ``tol`` across it (plain difference of principal values), when it is more than twice as long
    as a neighbouring edge, or longer than ``max_step`` -- unless it is already shorter than 1e-8.
    """
    pts = np.asarray(z_list, dtype=complex)
    fv = np.asarray((yield pts), dtype=complex)
    state = {} if state is None else state
    state.update(points=pts, f_points=fv, success=False)
    for _ in range(max_iter):
        before = np.roll(pts, 1)
        arg_jump = np.angle(fv) - np.angle(np.roll(fv, 1))
        length = np.abs(pts - before)
        uneven = (length > 2*np.roll(length, 1)) | (length > 2*np.roll(length, -1))
        split = (((np.abs(arg_jump) > tol) | uneven) | (length > max_step)) & (length > 1.0e-8)
        where = split.nonzero()[0]
        if len(where) == 0:
            shifted = fv + 1.0e-30j
            winding = 0.5*np.sum(np.angle(np.roll(shifted, -1)/shifted))/np.pi
            nearest = np.rint(winding)
            if np.abs(winding - nearest) < 1.0e-6:
                state["success"] = True
                return int(nearest)
            raise RuntimeError("No convergence to integer number of roots")
        mids = 0.5*(pts + before)[where]
        fm = np.asarray((yield mids), dtype=complex)
        if not np.all(np.isfinite(fm)):
            raise ValueError("Infinite value in f_add")
        pts = np.insert(pts, where, mids)
        fv = np.insert(fv, where, fm)
        state.update(points=pts, f_points=fv)
    raise RuntimeError("Maximum number of iterations exceeded")


class ClosedPath:
    """Piecewise linear closed contour (counter-clockwise ``z_list``) around which the roots of
    ``f`` are counted with the argument principle."""

    def __init__(self, f, z_list):
        self.func = f
        self.points = np.asarray(z_list)
        self.f_points = np.asarray(self.func(z_list))
        self.centre = np.mean(self.points)
        self.success = False

    def sum_angles(self):
        shifted = self.f_points + 1.0e-30j
        return 0.5*np.sum(np.angle(np.roll(shifted, -1)/shifted))/np.pi

    @staticmethod
    def interweave(a, b):
        """a[0], b[0], a[1], b[1], ... (reference complex_roots.py:540-546)."""
        a, b = np.asarray(a), np.asarray(b)
        out = np.empty(a.size + b.size, dtype=np.result_type(a, b))
        out[0::2], out[1::2] = a, b
        return out

    def refine(self):
        """Double the number of points: the mid-point of every edge (point -> next point) is inserted behind its
        first end (reference complex_roots.py:548-554); one batched evaluation of f."""
        mids = 0.5*(self.points + np.roll(self.points, -1))
        self.points = self.interweave(self.points, mids)
        self.f_points = self.interweave(self.f_points, np.asarray(self.func(mids)))

    def refine_select(self, tol=0.1, max_step=0.1):
        """ONE refinement pass of count_roots (reference complex_roots.py:556-595): bisect the edges across which
        the argument of f jumps by more than ``tol``, that are more than twice as long as a neighbour or longer
        than ``max_step`` (never below 1e-8).  Returns the number of points added; ValueError on non-finite f."""
        st = {}
        mach = contour_machine(self.points, 1, tol, max_step, st)
        next(mach)
        try:
            req = mach.send(np.asarray(self.f_points, dtype=complex))      # the pass asks for its mid-points ...
        except StopIteration:
            return 0                                                       # ... or finds nothing to split
        except RuntimeError:                                               # winding sum not an integer: not this
            return 0                                                       # method's business
        n_add = len(req)
        try:
            mach.send(np.asarray(self.func(req)))
        except RuntimeError:                                               # "maximum number of iterations": the one
            pass                                                           # pass is done
        self.points, self.f_points = st["points"], st["f_points"]
        return n_add

    def count_roots(self, verbose=False, max_iter=100, tol=0.1, max_step=0.1):
        """Number of roots inside the contour.  RuntimeError when the winding sum does not settle on
        an integer or max_iter passes are exhausted; ValueError on non-finite function values."""
        self.success = False
        st = {}
        mach = contour_machine(self.points, max_iter, tol, max_step, st)
        next(mach)                                     # corners were evaluated by the constructor
        try:
            req = mach.send(self.f_points)
            while True:
                vals = np.asarray(self.func(req))
                if verbose:
                    print(len(st["points"]), len(req))
                req = mach.send(vals)
        except StopIteration as stop:
            self.points, self.f_points, self.success = st["points"], st["f_points"], st["success"]
            return stop.value
        finally:
            if "points" in st:
                self.points, self.f_points = st["points"], st["f_points"]
