"""PSIDispersion -- drop-in for psitools.psi_dispersion.PSIDispersion (reference
psi_dispersion.py:30-456) whose ``calculate`` runs on the GPU.

Same constructor, same ``calculate(w, wave_number_x, wave_number_z, viscous_alpha=0)`` signature and
return-shape semantics.  Underneath, the background integrals J0/J1 (kernel K0) and every D(omega)
evaluation (kernel K1: pole finding, tanh-sinh of the seven joint M_ij integrands per sub-interval,
det(P+M)) are CUDA kernels reached through the C ABI of libpsi_b200.so.

Deviation from the reference, documented in DESIGN.md: ``tanhsinh_integrator=None`` selects
TanhSinh(15, 12) instead of QUADPACK (agreement <= 1.5e-13 relative at the reference's goldens).
"""
import numpy as np

from . import _device
from .sizedensity import SizeDensity
from .tanhsinh import TanhSinh

_default_integrator = None


def default_integrator():
    global _default_integrator
    if _default_integrator is None:
        _default_integrator = TanhSinh()
    return _default_integrator


class PSIDispersion:
    """The polydisperse streaming instability dispersion relation f(w) = 0; ``calculate`` gives f.

    Args:
        dust_to_gas_ratio: dust to gas ratio mu
        stokes_range: (tau_min, tau_max)
        sound_speed_over_eta: gas sound speed / eta (default 1/0.05)
        size_distribution_power: power-law index (default 3.5 = MRN); ignored with size_density
        single_size_flag: single size at tau_max
        size_density: any SizeDensity(sigma, size_range) -- PowerBump.sigma0 / PowerBumpTail.sigma0 are evaluated
            analytically on the device, every other callable through a table (_device.SizeDensityTable) -- or None
        tanhsinh_integrator: TanhSinh instance whose tables are uploaded to the device
    """

    def __init__(self, dust_to_gas_ratio, stokes_range, sound_speed_over_eta=1/0.05,
                 size_distribution_power=3.5, single_size_flag=False, size_density=None,
                 tanhsinh_integrator=None, device=None):
        self.mu = dust_to_gas_ratio
        self.taumin = stokes_range[0]
        self.taumax = stokes_range[1]
        self.c = sound_speed_over_eta
        self.ssf = single_size_flag
        self.tanhsinh_integrator = tanhsinh_integrator
        integ = default_integrator() if tanhsinh_integrator is None else tanhsinh_integrator
        self.desc, self.table = _device.describe_size_density(dust_to_gas_ratio, stokes_range, sound_speed_over_eta,
                                                              size_distribution_power, single_size_flag,
                                                              size_density)
        self.size_density = size_density
        if size_density is None:
            q = 3 - size_distribution_power
            self.size_density = SizeDensity(lambda x: np.power(x, q), stokes_range,
                                            sigma_integral=self.desc.rhod)
        self.poles = list(self.size_density.poles)
        self.ctx = _device.get_context(integ, device)
        self.dist_id = self.ctx.add_dist(self.desc, self.table)
        bg = self.ctx.background(self.dist_id)           # computed on the device (K0)
        j0, j1, denom = bg["j0"], bg["j1"], bg["denom"]
        self.j0, self.j1 = j0, j1
        self.vgx = bg["vgx"]
        self.vgy = bg["vgy"]
        self.ux = lambda x: 2*denom*(j1 - x*(1 + j0))/(1 + x*x)
        self.uy = lambda x: -denom*(1 + j0 + x*j1)/(1 + x*x)
        self.base_poly = [-2*denom*j1, 2*denom*(1 + j0), 0]

    def calculate(self, w, wave_number_x, wave_number_z, viscous_alpha=0, return_M=False):
        """Value of the dispersion relation at complex frequency ``w`` (scalar or any-shape array)."""
        self.kx = wave_number_x
        self.kz = wave_number_z
        self.nu = viscous_alpha*self.c**2
        w = np.asarray(w)
        flat = np.ravel(w).astype(np.complex128)
        res = self.ctx.disp_eval(self.dist_id, float(wave_number_x), float(wave_number_z),
                                 float(viscous_alpha), flat, want_M=return_M)
        D, M = res if return_M else (res, None)
        out = np.squeeze(D) if w.ndim == 0 else np.reshape(D, w.shape)
        if return_M:
            return out, M
        return out

    # -- the reference's public helpers (psi_dispersion.py:124-352), thin: none of them is on the hot path -------
    def average(self, g):
        """(1/rho_g) * integral of sigma(a) g(tau) da for a HOST callable g(tau) (reference psi_dispersion.py:124-181):
        mu*g(taumax) for a single size, else the tanh-sinh rule in u = ln s split at the size density's poles --
        on the host, with the same tables the device uses (a Python callable cannot run on the GPU; the product's own
        integrals never come through here)."""
        if self.ssf:
            return self.mu*g(self.taumax)
        integ = self.tanhsinh_integrator if self.tanhsinh_integrator is not None else default_integrator()
        t = self.taumax
        f = lambda u: g(t*np.exp(u))*np.exp(u)*self.size_density(np.exp(u))
        cuts = [np.log(self.taumin/self.taumax)] + [np.log(p) for p in self.poles] + [0.0]
        return self.mu*sum(integ.integrate(f, a, b) for a, b in zip(cuts[:-1], cuts[1:]))

    def int_j(self, alpha):
        """J_alpha = average(tau^alpha/(1 + tau^2)) (reference psi_dispersion.py:183-187)."""
        return self.average(lambda x: np.power(x, alpha)/(1 + x**2))

    def D(self, x):
        """Dust diffusion coefficient at stopping time x for the viscosity of the last calculate call (:390)."""
        return (1 + x + 4*x*x)*self.nu/(1 + x*x)**2

    def matrix_P(self, w):
        """3x3 matrix P of the linear gas momentum equation without drag, at the wave numbers and viscosity of the
        last calculate call (reference psi_dispersion.py:189-210)."""
        kx, kz, nu, c = self.kx, self.kz, self.nu, self.c
        r = w - kx*self.vgx
        d = kx*self.vgx - w - 1j*nu*(kx**2 + kz**2)
        third = 1j*nu/3
        return np.asarray([[d + (kx*c)**2/r - third*kx**2, 2j, kx*c**2*kz/r - third*kx*kz],
                           [-0.5j, d, 0],
                           [kz*c**2*kx/r - third*kx*kz, 0, d + (kz*c)**2/r - third*kz**2]])

    def _tau_matrices(self, w):
        """All stopping-time dependent 3x3 blocks of the drag term at once (arrays over x): A^-1, D, V, W."""
        kx, kz = self.kx, self.kz
        k2 = kx**2 + kz**2
        c0 = 1/(w - kx*self.vgx)

        def blocks(x):
            x = np.asarray(x, dtype=float)
            zero, one = 0*x + 0j, 0*x + 1.0 + 0j
            dux, duy = self.ux(x) - self.vgx, self.uy(x) - self.vgy
            it = 1j/x
            d = kx*self.ux(x) - w - it
            den = 1/(1 - d**2)
            Ainv = [[-d*den, 2j*den, zero], [-0.5j*den, -d*den, zero], [zero, zero, 1/d]]
            Dm = [[it*(dux*kx*c0 - 1), zero, it*dux*kz*c0], [it*duy*kx*c0, -it, it*duy*kz*c0], [zero, zero, -it]]
            dv = 1/(w - kx*self.ux(x) + 1j*(k2*self.D(x) + 1.0e-10))
            V = [[1 + dux*kx*dv, zero, dux*kz*dv], [duy*kx*dv, one, duy*kz*dv], [zero, zero, one]]
            q = 1j*self.D(x)*k2*dv*c0
            W = [[dux*kx*q - 1, zero, dux*kz*q], [duy*kx*q, -one, duy*kz*q], [zero, zero, -one]]
            return Ainv, Dm, V, W
        return blocks

    @staticmethod
    def _entrywise(fn):
        """3x3 nested list of callables x -> entry (the shape the reference's matrix_* methods return)."""
        return [[(lambda x, i=i, j=j: fn(x)[i][j]) for j in range(3)] for i in range(3)]

    @staticmethod
    def _matmul(a, b):
        return [[a[i][0]*b[0][j] + a[i][1]*b[1][j] + a[i][2]*b[2][j] for j in range(3)] for i in range(3)]

    def matrix_inverse_A(self, w):
        """Entries of A^-1 (linear dust momentum equation A u = D v_g) as callables of the stopping time (:212-227)."""
        blocks = self._tau_matrices(w)
        return self._entrywise(lambda x: blocks(x)[0])

    def matrix_AD(self, A, w):
        """A^-1 D as callables of the stopping time; ``A`` as returned by matrix_inverse_A (:229-278)."""
        blocks = self._tau_matrices(w)
        return self._entrywise(lambda x: self._matmul([[A[i][j](x) for j in range(3)] for i in range(3)], blocks(x)[1]))

    def matrix_VAD(self, AD, w):
        """V A^-1 D as callables of the stopping time; ``AD`` as returned by matrix_AD (:280-330)."""
        blocks = self._tau_matrices(w)
        return self._entrywise(lambda x: self._matmul(blocks(x)[2], [[AD[i][j](x) for j in range(3)] for i in range(3)]))

    def matrix_W(self, w):
        """Entries of W (dust diffusion terms) as callables of the stopping time (:332-352)."""
        blocks = self._tau_matrices(w)
        return self._entrywise(lambda x: blocks(x)[3])

    def matrix_M(self, w):
        """3x3 drag matrix at the wave numbers of the last ``calculate`` call."""
        _, M = self.calculate(w, self.kx, self.kz, self.nu/self.c**2, return_M=True)
        m = M[0]
        return np.asarray([[m[0], m[1], m[2]], [m[3], m[4], m[5]], [0, 0, m[6]]])
