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

    def matrix_M(self, w):
        """3x3 drag matrix at the wave numbers of the last ``calculate`` call."""
        _, M = self.calculate(w, self.kx, self.kz, self.nu/self.c**2, return_M=True)
        m = M[0]
        return np.asarray([[m[0], m[1], m[2]], [m[3], m[4], m[5]], [0, 0, m[6]]])
