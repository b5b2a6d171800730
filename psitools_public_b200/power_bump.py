"""Power law + cratering bump dust distributions (Birnstiel et al. 2011, 2015) -- host-side API
objects, drop-in for psitools.power_bump (reference power_bump.py:31-212).

The instances keep every derived constant as an attribute (``norm_pl``, ``amp``, ``curv``,
``beta``, ``bumpfac``) so that ``_device.describe_size_density`` can hand the exact same numbers to
the CUDA kernel, which evaluates the piecewise formula per quadrature node (csrc/psi_math.cuh,
``sigma_of``)."""
import numpy as np
import scipy.integrate
import scipy.optimize


def lognormpdf(y, s, loc, scale=1.0):
    """Log-normal pdf without scipy.stats (reference power_bump.py:31-35)."""
    x = (y - loc)/scale
    return np.exp(-np.log(x)**2/(2*s**2) - np.log(s*x*np.sqrt(2.0*np.pi)))


class PowerBump:
    """Normalised MRN-like power law with a Gaussian bump near the maximum size.

    amin: minimum monomer size; aL: left edge of the bump; aP: peak of the bump; aR: maximum size;
    bumpfac: bump height over the power law at aL (2.0 = Birnstiel model); beta: power-law index;
    epstot: optional total dust mass fraction to normalise to.
    """

    def __init__(self, amin, aP, aL=None, aR=None, bumpfac=2.0, beta=-3.5, epstot=None):
        self._setup(amin, aP, aL, aR, bumpfac, beta)
        self.mnorm = scipy.integrate.quad(self.mnn, amin, self.aR, epsrel=1e-15, limit=100,
                                          points=[self.get_discontinuity()])[0]
        self.epstot = epstot

    def _setup(self, amin, aP, aL, aR, bumpfac, beta):
        self.amin = amin
        self.aP = aP
        self.aL = 2.0*aP/3.0 if aL is None else aL
        self.aR = 1.56*aP if aR is None else aR
        self.bumpfac = bumpfac
        self.beta = beta
        width = np.max((np.min((np.abs(self.aR - aP), np.abs(self.aL - aP)))/np.sqrt(np.log(2.0)),
                        0.1*aP))
        self.norm_pl = aP**(-beta)
        self.amp = bumpfac*float(self.fnn(self.aL))
        self.curv = -1/width**2

    def fnn(self, a):
        """Power-law branch."""
        return self.norm_pl*a**self.beta

    def b(self, a):
        """Gaussian bump branch."""
        return self.amp*np.exp(self.curv*(a - self.aP)**2)

    def get_discontinuity(self):
        """Size where the bump overtakes the power law between aL and aP (the kink of sigma)."""
        diff = lambda x: float(self.fnn(x)) - float(self.b(x))
        if diff(self.aL)*diff(self.aP) > 0.0:
            return self.aL
        return scipy.optimize.bisect(diff, self.aL, self.aP)

    def mnn(self, a):
        """Un-normalised mass density a^3 n(a): power law below aL, max(power law, bump) up to aP,
        bump above."""
        arr = np.asarray(a, dtype=float)
        scalar = arr.ndim == 0
        flat = np.ravel(arr[None] if scalar else arr)
        pl = np.array(self.fnn(flat), dtype=float)
        bump = self.b(flat)
        val = np.where(flat > self.aL, np.maximum(pl, bump), pl)
        val = np.where(flat > self.aP, bump, val)*flat**3
        if scalar:
            return np.squeeze(val).item()
        return np.reshape(val, arr.shape)

    def sigma0(self, a):
        ret = self.mnn(a)
        if self.epstot is not None:
            ret /= self.mnorm
            ret *= self.epstot
        return ret


class PowerBumpTail(PowerBump):
    """PowerBump with a Brownian-motion tail n ~ a^0.5 below aBT (Birnstiel et al. 2015)."""

    def __init__(self, amin, aP, aL=None, aR=None, aBT=1e-6, bumpfac=2.0, beta=-3.5, epstot=None):
        self.aBT = aBT
        self.bpower = 0.5
        self._setup(amin, aP, aL, aR, bumpfac, beta)
        self.mnorm = scipy.integrate.quad(self.mnn, amin, self.aR, epsrel=1e-15, limit=100,
                                          points=[aBT, self.get_discontinuity()])[0]
        self.epstot = epstot

    def fnn(self, a):
        arr = np.asarray(a, dtype=float)
        scalar = arr.ndim == 0
        flat = np.ravel(arr[None] if scalar else arr)
        steep = self.norm_pl*flat**self.beta
        tail = self.norm_pl*self.aBT**self.beta*(flat/self.aBT)**self.bpower
        val = np.where(flat > self.aBT, steep, tail)
        if scalar:
            return np.squeeze(val).item()
        return np.reshape(val, arr.shape)
