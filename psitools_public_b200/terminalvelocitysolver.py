"""Polydisperse streaming instability in the terminal-velocity approximation: a closed-form dispersion
relation for power-law size distributions whose growing mode is followed from the single-size limit
(tau_min = tau_max) down to the wanted tau_min.  Same classes, arguments and results as the reference's
``psitools/terminalvelocitysolver.py`` (``TerminalVelocitySolver`` :27-193, ``PSIModeTV`` :195-418).

Why it is here: the terminal-velocity mode is a cheap GUESS for the root
searches of the full dispersion relation on (Kx, Kz) maps.  ``PSIModeTV.find_roots`` is the reference's
call (one wave-number pair, host arithmetic); ``PSIModeTV.find_roots_map`` runs the same continuation for
a whole map on the GPU, one thread per wave-number pair (``psi_tv_roots`` in ``include/psi_b200.h``,
``csrc/psi_tv.cu``) and needs the CUDA library.

The continuation (reference :135-172, :317-366): step s = tau_min/tau_max down from 1 in steps that grow by
sqrt(2) after an accepted root and shrink tenfold after a rejected one (secant failed, Im z < 0, Im z > 0.5
or NaN); every root is found by scipy's secant rule (``scipy.optimize.newton`` without a derivative) started
at the last accepted root; ``maximum_iterations`` rejected-or-accepted steps in total, then z = 0 ("mode
does not exist") for this and all smaller tau_min."""
import numpy as np
import scipy.optimize as opt


def _atanc(z):
    """arctan for complex arguments through logarithms (the reference's branch choice, :63-64)."""
    return 0.5j*(np.log(1 - 1j*z) - np.log(1 + 1j*z))


class _PowerLawTV:
    """What both classes share: F(z, s)/(1 + z) in closed form for b = -(3 + beta) in {0, 1/2, 3/2, 5/2},
    the size-averaged stopping time and the step-size controlled continuation in s."""
    guard = 0.0                  # added to the denominator 1 - s^(1-b)  (PSIModeTV: 1e-30)
    s_one = None                 # PSIModeTV: F = 1/(1+z) when 1 - s < s_one

    def _setup(self, dust_gas_ratio, maximum_stokes, beta, maximum_iterations):
        self.fd = dust_gas_ratio/(1.0 + dust_gas_ratio)
        self.tau_max = maximum_stokes
        self.b = -(3 + beta)
        self.max_it = maximum_iterations

    def atanc(self, z):
        return _atanc(z)

    def _closed_form(self, z, s):
        if self.s_one is not None and 1 - s < self.s_one:
            return 1/(1 + z)
        b = self.b
        p1, p2 = np.power(s, 1 - b), np.power(s, 2 - b)
        w = z*(b - 1)*(1 - p2)/((b - 2)*(1 - p1 + self.guard))
        arg = (1 - np.sqrt(s))/(np.sqrt(w) + np.sqrt(s/w))
        if b == 0.5:
            return 1 - np.sqrt(w)*_atanc(arg)/(1 - np.sqrt(s))
        if b == 1.5:
            return -_atanc(arg)/(np.sqrt(w)*(1 - p1))
        if b == 2.5:
            return (1 + _atanc(arg)/(np.sqrt(w)*(1 - p2)))/z
        if b == 0.0:
            return 1 - w*np.log((w + 1)/(w + s))/(1 - s)
        raise NotImplementedError(('No form for the I2 integral for this'
                                   ' parameter b {:e}').format(b))

    def average_tau(self, s):
        """Size-averaged stopping time for tau_min/tau_max = s (regularised at s -> 0)."""
        s = s*s/(s + 1.0e-10)
        num = (self.b - 1)*(1 - s**(self.b - 2))
        den = (self.b - 2)*(1 - s**(self.b - 1))
        return s*self.tau_max*num/den

    def _follow(self, residual, s_want, s_start, z_start):
        """Root of residual(z, s_want) by continuation from the known root z_start at s_start > s_want."""
        step, s_now, z = 0.1, s_start, z_start
        tries = 0
        while s_now > s_want:
            tries += 1
            if tries > self.max_it:
                print('Warning: maximum iterations reached at z = ', z, ', assuming z=0')
                return 0
            s = max(s_now - step, s_want)             # end exactly on s_want
            try:
                z, failed = opt.newton(lambda x: residual(x, s), z_start), False
            except (ValueError, RuntimeError):
                z, failed = z_start, True
            if failed or np.imag(z) < 0 or np.imag(z) > 0.5 or np.isnan(np.imag(z)):
                step = 0.1*step                       # rejected: smaller step from the same start
            else:
                step = np.sqrt(2)*step
                s_now, z_start = s, z
        return z


class TerminalVelocitySolver(_PowerLawTV):
    """Inviscid form with the wave numbers and the tau_min array fixed at construction (reference :27-193)."""

    def __init__(self, dust_gas_ratio, wave_number_x, wave_number_z, minimum_stopping_time,
                 maximum_stopping_time, power_law_exponent_size_distribution, maximum_iterations=100):
        self._setup(dust_gas_ratio, maximum_stopping_time, power_law_exponent_size_distribution,
                    maximum_iterations)
        self.kx, self.kz = wave_number_x, wave_number_z
        self.tau_min = minimum_stopping_time
        self.I2(1.0, 2.0)                             # raises for an unsupported exponent

    def I2(self, z, s):
        return self._closed_form(z, s)

    def single_size_z(self):
        """Growing root of the single-size cubic at tau = tau_max, in the z variable."""
        fg, k2 = 1 - self.fd, self.kx*self.kx + self.kz*self.kz
        roots = np.roots([1, (2*fg*fg*self.kx + 1j*self.fd)*self.tau_max, -self.kz**2/k2,
                          2*fg*(self.fd - fg)*self.kz**2*self.kx*self.tau_max/k2])
        w = roots[np.argmax(np.imag(roots))]
        return (w - 2*fg*self.fd*self.tau_max*self.kx)/(2*fg*self.kx*self.tau_max)

    def func(self, z, s):
        tau, fg = self.average_tau(s), 1 - self.fd
        k2 = self.kx*self.kx + self.kz*self.kz
        omega = 2*fg*self.kx*tau*(z + self.fd)
        A = 2*fg*k2*tau*tau*self.kx/(self.kz*self.kz)
        lhs = 1.0 - k2*omega*omega/(self.kz*self.kz)
        return self.fd*(1j*A*np.power(z + self.fd, 2) + 1)*self.I2(z, s) - lhs

    def find_root(self, s_want, s_start, z_start):
        return self._follow(self.func, s_want, s_start, z_start)

    def find_roots(self):
        """Mode frequency for every tau_min (ascending), worked from the largest down."""
        n = len(self.tau_min)
        z = np.zeros(n, dtype=np.complex128)
        s_start, z0 = 1.0, self.single_size_z()
        for j in range(n - 1, -1, -1):
            s_want = self.tau_min[j]/self.tau_max
            z[j] = self.find_root(s_want, s_start, z0)
            s_start, z0 = s_want, z[j]
            if z0 == 0:
                break
        tau = self.average_tau(self.tau_min/self.tau_max)
        return 2*(1 - self.fd)*self.kx*tau*(z + self.fd)


class PSIModeTV(_PowerLawTV):
    """Terminal-velocity PSI mode for power-law size distributions; wave numbers, tau_min and the dust
    diffusion (viscous_alpha) are arguments of ``find_roots`` (reference :195-418)."""
    guard = 1.0e-30
    s_one = 1.0e-12

    def __init__(self, dust_gas_ratio, maximum_stokes, power_law_exponent_size_distribution=-3.5,
                 maximum_iterations=100):
        self._setup(dust_gas_ratio, maximum_stokes, power_law_exponent_size_distribution, maximum_iterations)
        self.D = 0.0
        self.F_integral(1.0, 0.1)                     # raises for an unsupported exponent

    def F_integral(self, z, s):
        """F(z)/(1 + z)."""
        return self._closed_form(z, s)

    def _scale(self, tau, Kx):
        return 2*(1 - self.fd)*tau*Kx

    def omega_to_z(self, w, tau, Kx, Kz):
        fac = self._scale(tau, Kx)
        return (w - self.fd*fac + 1j*self.D*(Kx**2 + Kz**2))/fac

    def z_to_omega(self, z, tau, Kx, Kz):
        return self._scale(tau, Kx)*(z + self.fd) - 1j*self.D*(Kx**2 + Kz**2)

    def single_size_z(self, Kx, Kz):
        """Growing mode of the single-size cubic at tau = tau_max (with dust diffusion), in z."""
        K2, fg = Kx*Kx + Kz*Kz, 1 - self.fd
        roots = np.roots([1, (2*fg*fg*Kx + 1j*self.fd)*self.tau_max + 1j*self.D*K2,
                          self.fd*self.tau_max*self.D*K2 - Kz**2/K2,
                          2*fg*(self.fd - fg)*Kz**2*Kx*self.tau_max/K2 - 1j*self.D*Kz**2])
        return self.omega_to_z(roots[np.argmax(np.imag(roots))], self.tau_max, Kx, Kz)

    def disp(self, z, s, Kx, Kz):
        """Dispersion relation, zero at a mode."""
        tau = self.average_tau(s)
        omega = self.z_to_omega(z, tau, Kx, Kz)
        K2 = Kx*Kx + Kz*Kz
        lhs = 1.0 - K2*omega**2/Kz**2
        return self.fd*(1j*omega*tau*K2*(z + self.fd)/Kz**2 + 1)*self.F_integral(z, s) - lhs

    def find_root(self, s_want, s_start, z_start, Kx, Kz):
        return self._follow(lambda z, s: self.disp(z, s, Kx, Kz), s_want, s_start, z_start)

    def find_roots(self, minimum_stokes, Kx, Kz, viscous_alpha=0, c_over_eta=20):
        """Mode frequency for the minimum Stokes number(s) (ascending) at one wave-number pair."""
        self.D = viscous_alpha*c_over_eta**2
        tau_min = np.asarray(minimum_stokes)
        scalar = tau_min.ndim == 0
        shape = np.shape(minimum_stokes)
        tau_min = tau_min[None] if scalar else np.ravel(minimum_stokes)
        z = np.zeros(np.shape(tau_min), dtype=np.complex128)
        s_start, z0 = 1.0, self.single_size_z(Kx, Kz)
        for j in range(len(tau_min) - 1, -1, -1):
            s_want = tau_min[j]/self.tau_max
            z[j] = self.find_root(s_want, s_start, z0, Kx, Kz)
            s_start, z0 = s_want, z[j]
            if z0 == 0:
                break
        ret = self.z_to_omega(z, self.average_tau(tau_min/self.tau_max), Kx, Kz)
        return np.squeeze(ret) if scalar else np.reshape(ret, shape)

    def find_roots_map(self, minimum_stokes, Kx, Kz, viscous_alpha=0, c_over_eta=20, device=None):
        """``find_roots`` for ONE minimum Stokes number on arrays of wave-number pairs, on the GPU (one thread
        per pair runs the whole continuation, csrc/psi_tv.cu).  Returns (omega, status): omega has the shape of
        Kx; status 0 = mode found, 1 = gave up after maximum_iterations (omega is then the value the reference
        returns for z = 0)."""
        from . import _device
        Kx, Kz = np.broadcast_arrays(np.asarray(Kx, dtype=float), np.asarray(Kz, dtype=float))
        w, st = _device.tv_roots(self.fd, self.tau_max, self.b, float(minimum_stokes),
                                 float(viscous_alpha*c_over_eta**2), int(self.max_it), Kx.ravel(), Kz.ravel(),
                                 device=device)
        return w.reshape(Kx.shape), st.reshape(Kx.shape)
