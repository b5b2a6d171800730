"""Tanh-sinh (double exponential) quadrature tables -- host-side API object.

Drop-in for ``psitools.tanhsinh`` (reference tanhsinh.py:27-141).  The hot path never calls
``integrate`` here: ``PSIDispersion`` uploads ``xj``/``wj`` to the GPU once and the level loop runs
in the CUDA kernel (csrc/psi_quad.cuh).  ``integrate`` is kept because the reference exposes it for
arbitrary Python callables (e.g. user-side normalisation integrals), which cannot run on a device.
"""
import warnings

import numpy as np


class TanhSinh:
    """Abscissae ``xj`` and weights ``wj`` of the tanh-sinh rule at step 2**-max_level.

    Args:
        precision_digit: digits of precision requested (default 15, about double precision).
        max_level: finest level; the smallest step is 2**-max_level (default 12).
    """

    def __init__(self, precision_digit=15, max_level=12):
        self.p = precision_digit
        self.max_m = max_level
        self.eps = np.power(10.0, -self.p)
        hmin = np.power(2.0, -self.max_m)
        # enough terms to reach 1 - eps (reference tanhsinh.py:46)
        count = int(np.arcsinh((2.0/np.pi)*np.arctanh(1 - self.eps))/hmin) + 10
        arg = hmin*np.arange(count)
        half_pi_sinh = 0.5*np.pi*np.sinh(arg)
        half_pi_cosh = 0.5*np.pi*np.cosh(arg)
        outer = np.cosh(half_pi_sinh)
        absc = 1.0 - 1.0/(np.exp(half_pi_sinh)*outer)
        wts = half_pi_cosh/(outer*outer)
        ok = np.logical_and(absc < 1.0 - self.eps, wts > self.eps).nonzero()
        self.xj = absc[ok]
        self.wj = half_pi_cosh[ok]/(outer[ok]*outer[ok])

    def t(self, x, a, b):
        """Map [-1, 1] onto [a, b]."""
        return 0.5*(x*(b - a) + b + a)

    def integrate(self, func, a, b, tol=None):
        """Integrate a vectorised Python callable over [a, b] (host utility; see module note)."""
        if tol is None:
            tol = self.eps
        elif tol < self.eps:
            warnings.warn('TanhSinh: Requested tolerance too small, setting tol = {:e}'
                          .format(self.eps))
            tol = self.eps
        self.f = lambda x: 0.5*(b - a)*func(self.t(x, a, b))
        inside = np.asarray(self.t(self.xj, a, b) < b - self.eps).nonzero()
        xs, ws = self.xj[inside], self.wj[inside]
        stride = 2**self.max_m
        total = np.sum(ws[::stride]*self.f(xs[::stride])) + \
            np.sum(ws[stride::stride]*self.f(-xs[stride::stride]))
        for level in range(1, self.max_m + 1):
            stride = 2**(self.max_m - level)
            xo, wo = xs[::stride][1::2], ws[::stride][1::2]
            h = np.power(2.0, -level)
            change = -0.5*total + h*np.sum(wo*self.f(xo)) + h*np.sum(wo*self.f(-xo))
            if np.abs(change) < tol:
                break
            total = total + change
        return total


class TanhSinhNoDeepCopy(TanhSinh):
    """``copy.deepcopy`` returns the instance itself, so arg-dict campaigns (psi_grid_refine,
    MpiScheduler) share one table (reference tanhsinh.py:131-141)."""

    def __deepcopy__(self, memo):
        return self
