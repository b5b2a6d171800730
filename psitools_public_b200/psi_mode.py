"""PSIMode -- drop-in for psitools.psi_mode.PSIMode (reference psi_mode.py:40-491).

The root search (random samples -> AAA rational approximation -> zeros -> zoom domains -> secant
polish on the exact dispersion relation) is written once as the generator ``mode_machine``: it
yields batches of frequencies and receives D there.  ``PSIMode.calculate`` drives one machine;
``solve_modes`` drives thousands of them in lock-step (one K1 launch per pass, see _engine.py), which
is what ``MpiScheduler.run`` and ``PSIGridRefiner`` use for maps.

``plot_dispersion(show_exact=True)`` (reference psi_mode.py:493-558) evaluates the exact D on the whole
plot grid in ONE K1 launch (``dispersion_map``); matplotlib is imported only when a figure is drawn.
"""
import numpy as np

from . import complex_roots as cr
from . import psi_dispersion as psid
from ._engine import Job, run_lockstep


def guess_domain_size(x):
    """Size of the zoom box put around a guessed root (may be used from other modules)."""
    return min((1e-3, np.abs(x.imag)*4))


class ModeSettings:
    """The PSIMode constructor knobs that steer the search."""

    def __init__(self, real_range, imag_range, n_sample=20, max_zoom_domains=1, tol=1.0e-13,
                 clean_tol=1.0e-4, max_secant_iterations=100):
        self.domain = cr.Rectangle(real_range, imag_range)
        self.n_sample = n_sample
        self.max_zoom_level = max_zoom_domains
        self.max_zoom_domains_per_level = 4
        self.max_secant_iterations = max_secant_iterations
        self.minimum_interpolation_nodes = 4
        self.force_at_least_one_root = True
        self.tol = tol
        self.clean_tol = clean_tol


class ModeTrace:
    """What PSIMode exposes after a call: counters, samples and the rational approximation."""

    def __init__(self):
        self.n_function_call = 0
        self.max_convergence_iterations = 0
        self.z_sample = None
        self.f_sample = None
        self.ra = None
        self.n_roots = None


def _boxed_centre(dom, centre, size):
    """Shift a zoom box centre so that the whole box lies inside the search domain."""
    re, im = np.real(centre), np.imag(centre)
    if re < dom.xmin + 0.5*size[0]:
        re = dom.xmin + 0.5*size[0]
    if re > dom.xmax - 0.5*size[0]:
        re = dom.xmax - 0.5*size[0]
    if im < dom.ymin + 0.5*size[1]:
        im = dom.ymin + 0.5*size[1]
    if im > dom.ymax - 0.5*size[1]:
        im = dom.ymax - 0.5*size[1]
    return re + 1j*im


def mode_machine(cfg, guess_roots=(), count_roots=False, rng=None, trace=None, log=None):
    """Generator implementing PSIMode.calculate for one (Kx, Kz, alpha).

    Yields arrays of complex frequencies, receives the dispersion relation there, returns the
    array of roots inside the search rectangle (reference psi_mode.py:106-333).  ``rng`` supplies
    ``uniform`` (numpy's global legacy state by default, a per-item RandomState in batched runs).
    """
    trace = ModeTrace() if trace is None else trace
    say = (lambda *a: None) if log is None else log
    dom = cfg.domain
    ra = cr.RationalApproximation(dom, tol=cfg.tol, clean_tol=cfg.clean_tol)
    trace.ra = ra

    zs = dom.generate_random_sample_points(cfg.n_sample, rng)
    fs = yield zs
    trace.n_function_call = cfg.n_sample

    if count_roots:
        st = {}
        trace.n_roots = yield from cr.contour_machine(dom.corners(), state=st)
        say('Number of roots: {}'.format(trace.n_roots))
        zs = np.concatenate((zs, st["points"]))
        fs = np.concatenate((fs, st["f_points"]))

    def extra_domain(zs, fs, size, centre):
        centre = _boxed_centre(dom, centre, size)
        say('Adding zoom domain around {}'.format(centre))
        box = cr.Rectangle([centre.real - 0.5*size[0], centre.real + 0.5*size[0]],
                           [centre.imag - 0.5*size[1], centre.imag + 0.5*size[1]])
        zn = box.generate_random_sample_points(cfg.n_sample, rng)
        fn = yield zn
        trace.n_function_call += cfg.n_sample
        return np.concatenate((zs, zn)), np.concatenate((fs, fn))

    for centre in guess_roots:
        side = guess_domain_size(centre)
        zs, fs = yield from extra_domain(zs, fs, [side, side], centre)

    roots = np.zeros(0, dtype=complex)
    for level in range(0, cfg.max_zoom_level + 1):
        trace.z_sample, trace.f_sample = zs, fs
        ra.calculate(fs, zs)
        zeros = ra.find_zeros()
        zeros = zeros[np.asarray(dom.is_in(zeros)).nonzero()]

        # too few support points or no growing zero: halve clean_tol and redo (down to tol)
        need = cfg.minimum_interpolation_nodes*(1 + level + len(guess_roots))
        saved = ra.clean_tol
        while (np.sum(ra.maskF) < need or
               (not np.any(np.imag(zeros) > 0.0) and cfg.force_at_least_one_root)):
            ra.clean_tol = 0.5*ra.clean_tol
            if ra.clean_tol < ra.tol:
                break
            say('Reducing clean_tol to {}'.format(ra.clean_tol))
            ra.calculate(fs, zs)
            zeros = ra.find_zeros()
            zeros = zeros[np.asarray(dom.is_in(zeros)).nonzero()]
        ra.clean_tol = saved
        say('Number of nodes used: {}'.format(np.sum(ra.maskF)))

        zeros = ra.find_zeros()
        say('All zeros: {}'.format(zeros))
        # zoom candidates: anything below the top of the box and inside in Re (may be damped)
        cand = zeros[np.asarray((zeros.imag < dom.ymax) & (zeros.real < dom.xmax) &
                                (zeros.real > dom.xmin)).nonzero()]
        box_dx = np.power(10.0, -1 - 0.5*level)
        if len(cand) > 0:
            cand = cand[np.argsort(cand.imag)]
            zoom = [cand[-1]]
            for z in cand[-2::-1]:
                if not any(np.abs(z.real - q.real) < box_dx for q in zoom):
                    zoom.append(z)
            zoom = zoom[:cfg.max_zoom_domains_per_level]
            least_damped = zoom[0]
        else:
            zoom, least_damped = [], None

        zeros = np.asarray(zeros[np.asarray(dom.is_in(zeros)).nonzero()])
        say('All zeros in domain: {}'.format(zeros))
        max_iter = cfg.max_secant_iterations
        if level < cfg.max_zoom_level:
            max_iter = cfg.n_sample
        if len(zeros) == 0 and least_damped is not None:
            zeros = np.asarray([least_damped])
            max_iter = 5

        roots = yield from polish_machine(cfg, zeros, max_iter, trace, say)
        if np.any(roots.imag > 0.0) or level == cfg.max_zoom_level or len(zoom) == 0:
            return roots
        for centre in zoom:
            zs, fs = yield from extra_domain(zs, fs, [box_dx, np.min([0.1, np.abs(centre.imag)])],
                                             centre)
    return roots


def polish_machine(cfg, starts, max_iter, trace, say=lambda *a: None, initial_iterations=5,
                   select_inside_domain=True, per_start=True, keep_all=False):
    """Secant polish of every start value on the exact dispersion relation, one after the other
    (reference psi_mode.py:340-352, 411-491): five iterations first; if that has not converged but
    stayed within |start| of the start, continue from there for up to ``max_iter``.  Failed or
    outside-domain results are dropped.  ``per_start=True`` (find_dispersion_roots, which hands find_root
    one scalar at a time): the step tolerance min(1.48e-8, 1e-5*|Im start|) is that of the current start
    value alone, and of its restart point after a restart; ``per_start=False`` (the public find_root called
    with an array, :438-444): the minimum runs over the whole current start array."""
    w0 = np.atleast_1d(np.asarray(starts, dtype=complex)).ravel().copy()
    out = np.zeros(len(w0), dtype=complex)
    eps = 1.0e-6
    for i in range(len(w0)):
        say('Starting iteration at z = {}'.format(w0[i]))
        xtol = np.min(np.hstack([[1.48e-8], np.abs((w0[i:i + 1] if per_start else w0).imag*1e-5)]))
        res = yield from cr.secant_machine(w0[i], cr.second_point(w0[i], eps), initial_iterations, xtol)
        trace.n_function_call += res.function_calls
        if (not res.converged and np.abs(res.root - w0[i])/np.abs(w0[i]) < 1
                and max_iter > initial_iterations):
            say("Starting further iteration")
            w0[i] = res.root
            xtol = np.min(np.hstack([[1.48e-8], np.abs((w0[i:i + 1] if per_start else w0).imag*1e-5)]))
            res = yield from cr.secant_machine(w0[i], cr.second_point(w0[i], eps), max_iter, xtol)
            trace.n_function_call += res.function_calls
        elif not res.converged:
            say("Secant is running away; no further iterations")
        inside = bool(cfg.domain.is_in(res.root)) or not select_inside_domain
        if res.converged and inside:
            trace.max_convergence_iterations = max(trace.max_convergence_iterations, res.iterations)
            out[i] = res.root
        else:
            out[i] = -1j                      # sentinel, outside every search rectangle
    if keep_all:
        return out
    return out[np.asarray(cfg.domain.is_in(out)).nonzero()]


class PSIMode:
    """Calculates PSI modes (complex frequencies of the polydisperse streaming instability).

    Args: as psitools.psi_mode.PSIMode -- dust_to_gas_ratio, stokes_range, real_range, imag_range,
    sound_speed_over_eta=1/0.05, size_distribution_power=3.5, single_size_flag=False, n_sample=20,
    max_zoom_domains=1, verbose_flag=False, size_density=None, tol=1e-13, clean_tol=1e-4,
    max_secant_iterations=100, tanhsinh_integrator=None.
    """

    def __init__(self, dust_to_gas_ratio, stokes_range, real_range, imag_range,
                 sound_speed_over_eta=1/0.05, size_distribution_power=3.5, single_size_flag=False,
                 n_sample=20, max_zoom_domains=1, verbose_flag=False, size_density=None,
                 tol=1.0e-13, clean_tol=1.0e-4, max_secant_iterations=100,
                 tanhsinh_integrator=None, device=None, engine=None):
        self.psi_dispersion = psid.PSIDispersion(dust_to_gas_ratio, stokes_range, sound_speed_over_eta,
                                                 size_distribution_power, single_size_flag,
                                                 size_density, tanhsinh_integrator, device=device)
        self.dispersion = self.psi_dispersion.calculate
        # "device" (default): a calculate() call is one search of the K4 pipeline; "python": the generator form
        # below (LAPACK AAA on the host, D on the GPU) -- what verbose and count_roots=True calls always use
        import os
        self.engine = engine or os.environ.get("PSI_B200_ENGINE", "device")
        if self.engine == "native":
            self.engine = "python"
        self.cfg = ModeSettings(real_range, imag_range, n_sample, max_zoom_domains, tol, clean_tol,
                                max_secant_iterations)
        self.domain = self.cfg.domain
        self.n_sample = n_sample
        self.max_zoom_level = max_zoom_domains
        self.max_zoom_domains_per_level = self.cfg.max_zoom_domains_per_level
        self.max_secant_iterations = max_secant_iterations
        self.verbose_flag = verbose_flag
        self.max_convergence_iterations = 0
        self.minimum_interpolation_nodes = self.cfg.minimum_interpolation_nodes
        self.force_at_least_one_root = True
        self.n_function_call = 0
        self.guess_domain_size = guess_domain_size
        self.ra = cr.RationalApproximation(self.domain, tol=tol, clean_tol=clean_tol)

    def log_print(self, arg):
        if self.verbose_flag:
            print(arg)

    def _sync_cfg(self):
        c = self.cfg
        c.n_sample, c.max_zoom_level = self.n_sample, self.max_zoom_level
        c.max_zoom_domains_per_level = self.max_zoom_domains_per_level
        c.max_secant_iterations = self.max_secant_iterations
        c.minimum_interpolation_nodes = self.minimum_interpolation_nodes
        c.force_at_least_one_root = self.force_at_least_one_root
        c.tol, c.clean_tol = self.ra.tol, self.ra.clean_tol

    def make_job(self, wave_number_x, wave_number_z, viscous_alpha=0, guess_roots=(),
                 count_roots=False, rng=None, trace=None):
        """A lock-step job for this mode search (used by the batched drivers)."""
        self._sync_cfg()
        mach = mode_machine(self.cfg, guess_roots, count_roots, rng, trace,
                            self.log_print if self.verbose_flag else None)
        return Job(mach, self.psi_dispersion.dist_id, wave_number_x, wave_number_z, viscous_alpha)

    def _device_search_ok(self, guess_roots, count_roots):
        """A single search goes through the K4 pipeline (everything on the GPU) unless it needs what only the
        generator form offers: verbose logging, the count_roots coupling, non-default internal knobs, more than 512
        samples, or a context without the pipeline."""
        ctx = self.psi_dispersion.ctx
        if self.engine != "device" or not hasattr(ctx, "mode_search_detail"):
            return False
        if self.verbose_flag or count_roots:
            return False
        if (self.max_zoom_domains_per_level, self.minimum_interpolation_nodes, self.force_at_least_one_root) != \
                (4, 4, True):
            return False
        return self.n_sample >= 2 and self.n_sample*(1 + len(guess_roots) + 4*self.max_zoom_level) <= 512

    def _calculate_on_device(self, wave_number_x, wave_number_z, viscous_alpha, guess_roots):
        """PSIMode.calculate as ONE search of the K4 pipeline (C ABI psi_mode_search_detail).  The sample points come
        from numpy's global legacy stream exactly as in the reference (complex_roots.py:404-409): the numbers the
        longest possible search could need are drawn ahead, and afterwards the generator is put where the reference
        would have left it (only the numbers actually consumed count)."""
        ns, mz = int(self.n_sample), int(self.max_zoom_level)
        guesses = [complex(g) for g in guess_roots]
        state = np.random.get_state()
        u = np.random.random_sample(2*ns*(1 + len(guesses) + 4*mz))
        d = self.domain
        res = self.psi_dispersion.ctx.mode_search_detail(
            self.psi_dispersion.dist_id, float(wave_number_x), float(wave_number_z), float(viscous_alpha),
            [d.xmin, d.xmax, d.ymin, d.ymax], ns, mz, int(self.max_secant_iterations), float(self.ra.tol),
            float(self.ra.clean_tol), u, guesses)
        np.random.set_state(state)
        np.random.random_sample(res["n_uniforms_used"])
        self.n_function_call = res["n_function_call"]
        self.z_sample, self.f_sample = res["z_sample"], res["f_sample"]
        if res["status"] == 1:
            raise ValueError('Nan or Inf in sample points or function samples')
        if res["status"] == 3:
            raise ValueError('tol too small (0 <= 0)')
        if res["status"] not in (0, 4):
            raise RuntimeError('device root search failed with status %d' % res["status"])
        self.max_convergence_iterations = max(self.max_convergence_iterations, res["max_convergence_iterations"])
        # the rational approximation callers look at (ra.maskF, ra.evaluate, find_minimum, plot_dispersion): support
        # points from the device, weights refitted on the host (one small SVD; the device's own weights differ from
        # them by a phase only)
        ra = cr.RationalApproximation(self.domain, tol=self.ra.tol, clean_tol=self.ra.clean_tol)
        ra.F = ra.f = res["f_sample"]
        ra.Z = ra.z = res["z_sample"]
        ra.M = len(res["z_sample"])
        ra.maskF = res["mask"].astype(float)
        if ra.maskF.sum() >= 1 and ra.M > ra.maskF.sum():
            ra.calc_weights_residuals()
        self.ra = ra
        return res["roots"]

    def calculate(self, wave_number_x, wave_number_z, viscous_alpha=0, guess_roots=[],
                  count_roots=False):
        """Complex mode frequencies at (Kx, Kz) and viscosity parameter viscous_alpha."""
        self.disp = lambda z: self.dispersion(z, wave_number_x, wave_number_z, viscous_alpha)
        pd = self.psi_dispersion                     # the reference's PSIDispersion.calculate leaves these behind
        pd.kx, pd.kz, pd.nu = wave_number_x, wave_number_z, viscous_alpha*pd.c**2
        if self._device_search_ok(guess_roots, count_roots):
            try:
                return self._calculate_on_device(wave_number_x, wave_number_z, viscous_alpha, guess_roots)
            except Exception:
                print(('Root search failed at Kx = {}, Kz = {}, alpha = {}, guess_roots = {}')
                      .format(wave_number_x, wave_number_z, viscous_alpha, guess_roots))
                raise
        trace = ModeTrace()
        job = self.make_job(wave_number_x, wave_number_z, viscous_alpha, guess_roots, count_roots,
                            None, trace)
        try:
            run_lockstep(self.psi_dispersion.ctx, [job])
        except Exception:
            print(('Root search failed at Kx = {}, Kz = {}, alpha = {}, guess_roots = {}')
                  .format(wave_number_x, wave_number_z, viscous_alpha, guess_roots))
            raise
        finally:
            self.n_function_call = trace.n_function_call
            self.max_convergence_iterations = max(self.max_convergence_iterations,
                                                  trace.max_convergence_iterations)
            self.z_sample, self.f_sample = trace.z_sample, trace.f_sample
            if trace.ra is not None:
                self.ra = trace.ra
            if trace.n_roots is not None:
                self.n_roots = trace.n_roots
        return job.result

    def find_dispersion_roots(self, zeros, max_iter):
        trace = ModeTrace()
        self._sync_cfg()
        job = Job(polish_machine(self.cfg, zeros, max_iter, trace), self.psi_dispersion.dist_id,
                  self.psi_dispersion.kx, self.psi_dispersion.kz,
                  self.psi_dispersion.nu/self.psi_dispersion.c**2)
        run_lockstep(self.psi_dispersion.ctx, [job])
        self.n_function_call += trace.n_function_call
        return job.result

    def add_extra_domain(self, extra_domain_size=[0.1, 0.1], centre=0.0):
        """n_sample more random samples in a box around ``centre`` (shifted to lie inside the search domain), appended
        to z_sample / f_sample; their D values are ONE K1 launch (reference psi_mode.py:354-397).  Needs a preceding
        calculate call (or ``self.disp``) like the reference."""
        centre = _boxed_centre(self.domain, complex(centre), extra_domain_size)
        self.log_print('Adding zoom domain around {}'.format(centre))
        box = cr.Rectangle([centre.real - 0.5*extra_domain_size[0], centre.real + 0.5*extra_domain_size[0]],
                           [centre.imag - 0.5*extra_domain_size[1], centre.imag + 0.5*extra_domain_size[1]])
        z_new = box.generate_random_sample_points(self.n_sample)
        f_new = self.disp(z_new)
        self.z_sample = np.concatenate((self.z_sample, z_new))
        self.f_sample = np.concatenate((self.f_sample, f_new))
        self.n_function_call += self.n_sample

    def find_root(self, w0, max_iter=1000, select_inside_domain=True, initial_iterations=5):
        """Secant polish of one start value or of an array of them on the exact dispersion relation of the last
        calculate call (reference psi_mode.py:411-491): ``initial_iterations`` first, then up to ``max_iter`` from where
        that stopped if it has not run away; failures and (with ``select_inside_domain``) roots outside the search
        rectangle give -1j.  The step tolerance min(1.48e-8, 1e-5*min|Im w0|) runs over the WHOLE start array, as in
        the reference (find_dispersion_roots calls this with one scalar at a time).  Same shape out as in."""
        w0 = np.asarray(w0, dtype=complex)
        trace = ModeTrace()
        self._sync_cfg()
        pd = self.psi_dispersion
        job = Job(polish_machine(self.cfg, w0, max_iter, trace, self.log_print if self.verbose_flag else (lambda *a: None),
                                 initial_iterations, select_inside_domain, per_start=False, keep_all=True),
                  pd.dist_id, pd.kx, pd.kz, pd.nu/pd.c**2)
        run_lockstep(pd.ctx, [job])
        self.n_function_call += trace.n_function_call
        self.max_convergence_iterations = max(self.max_convergence_iterations, trace.max_convergence_iterations)
        out = np.asarray(job.result, dtype=complex)
        return np.squeeze(out) if w0.ndim == 0 else out.reshape(w0.shape)

    def find_minimum(self):
        """Location of the minimum of |rational approximation| in the domain (reference
        psi_mode.py:399-409; host-side scipy differential evolution on the approximant, no D calls)."""
        from scipy import optimize as opt
        bounds = ((self.domain.xmin, self.domain.xmax), (self.domain.ymin, self.domain.ymax))
        ret = opt.differential_evolution(lambda x: np.abs(self.ra.evaluate(x[0] + 1j*x[1])), bounds)
        self.log_print(ret)
        return ret.x[0] + 1j*ret.x[1]

    def dispersion_map(self, wave_number_x, wave_number_z, viscous_alpha=0, N=100, show_exact=False,
                       x=None, y=None):
        """The arrays ``plot_dispersion`` draws (reference psi_mode.py:507-517): ``x, y``, the rational
        approximation on the meshgrid and, with ``show_exact``, the exact D there -- one K1 launch over
        the N*N frequencies instead of the reference's per-frequency Python loop."""
        if x is None or y is None:
            x, y = self.domain.grid_xy(N=N)
        xv, yv = np.meshgrid(x, y)
        z = xv + 1j*yv
        f_appro = self.ra.evaluate(z)
        f_exact = None
        if show_exact:
            f_exact = self.dispersion(z, wave_number_x, wave_number_z, viscous_alpha=viscous_alpha)
        return x, y, f_appro, f_exact

    def plot_dispersion(self, wave_number_x, wave_number_z, viscous_alpha=0, N=100, show_exact=False,
                        x=None, y=None):
        """Plot the approximate dispersion relation over the domain and, with ``show_exact``, the exact
        one beside it (reference psi_mode.py:493-558).  Needs matplotlib."""
        import matplotlib.pyplot as plt
        x, y, f_appro, f_exact = self.dispersion_map(wave_number_x, wave_number_z, viscous_alpha, N,
                                                     show_exact, x, y)
        if show_exact:
            plt.subplot(122)
            plt.contourf(x, y, np.log10(np.abs(f_exact)), 20)
            plt.colorbar()
            plt.subplot(121)
        plt.contourf(x, y, np.log10(np.abs(f_appro)), 20)
        plt.colorbar()

        def inside(zs):
            zs = np.asarray(zs)
            sel = (zs.real > x[0]) & (zs.real < x[-1]) & (zs.imag > y[0]) & (zs.imag < y[-1])
            return zs[sel]

        nodes = inside(self.z_sample)
        plt.plot(nodes.real, nodes.imag, linestyle='', marker='.', color='black')
        nodes = inside(np.asarray(self.z_sample)[np.asarray(self.ra.maskF).astype(bool)])
        plt.plot(nodes.real, nodes.imag, linestyle='', marker='.', color='red')
        plt.show()
