"""Size density wrapper -- host-side API object (drop-in for psitools.sizedensity, reference
sizedensity.py:25-48).  The device evaluates the analytic family behind ``sigma``; see
``_device.describe_size_density``."""
import scipy.integrate as integrate


class SizeDensity:
    """sigma(a) on [amin, amax]; calling the instance with s = a/amax returns
    amax*sigma(amax*s)/rho_d.

    Args:
        sigma: size density function of size a
        size_range: (amin, amax)
        sigma_integral: integral of sigma over the range; computed with QUADPACK when omitted
    """

    def __init__(self, sigma, size_range, sigma_integral=None):
        self.amin = size_range[0]
        self.amax = size_range[1]
        self.sigma = sigma
        if sigma_integral is None:
            self.rhod = integrate.quad(sigma, self.amin, self.amax)[0]
        else:
            self.rhod = sigma_integral
        self.f = lambda x: self.amax*sigma(self.amax*x)/self.rhod
        self.poles = []

    def __call__(self, x):
        return self.f(x)
