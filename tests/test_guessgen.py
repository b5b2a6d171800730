"""the guess generators for map sweeps -- RootFollower (reference
complex_roots.py:428-515) and the terminal-velocity mode (terminalvelocitysolver.py) -- against values of the
unmodified reference (tests/golden/guessgen.json, section `guessgen` of make_golden.py) and, for the batched
device form, against the host form."""
import contextlib
import io

import numpy as np
import pytest

from common import carr, cplx, load


def _follow_func(z, k):
    return z**3 - (1.0 + 0.5j)*z - (k + 0.2j)


def test_root_follower_matches_reference():
    from psitools_public_b200 import RootFollower
    for case in load("guessgen.json")["follow"]:
        rf = RootFollower(_follow_func, cplx(case["z_start"]), case["k_start"])
        got = rf.calculate(np.array(case["k_want"]))
        np.testing.assert_allclose(got, carr(case["roots"]), rtol=1e-12, atol=1e-14)
    # a target equal to the current k ends the march instead of looping (the reference never returns here)
    rf = RootFollower(_follow_func, cplx(case["z_start"]), case["k_start"])
    assert np.all(np.isfinite(rf.calculate(np.array([0.6, 0.6]))))


def test_terminal_velocity_classes_match_reference():
    from psitools_public_b200 import PSIModeTV, TerminalVelocitySolver
    gold = load("guessgen.json")
    for c in gold["psimodetv"]:
        with contextlib.redirect_stdout(io.StringIO()):
            w = PSIModeTV(c["mu"], c["tau_max"], c["beta"]).find_roots(np.array(c["tau_min"]), c["kx"], c["kz"],
                                                                       viscous_alpha=c["alpha"])
            w1 = PSIModeTV(c["mu"], c["tau_max"], c["beta"]).find_roots(c["tau_min"][0], c["kx"], c["kz"],
                                                                        viscous_alpha=c["alpha"])
        np.testing.assert_allclose(w, carr(c["omega"]), rtol=1e-10, atol=1e-16)
        assert np.ndim(w1) == 0 and abs(w1 - cplx(c["omega_scalar"])) <= 1e-10*abs(w1)
    tau_min = np.logspace(-8, -2, 100)
    w = TerminalVelocitySolver(3, 46, 1, tau_min, 0.01, -3.5).find_roots()
    np.testing.assert_allclose(w[::9], carr(gold["tvsolver_test_0"]), rtol=1e-10)
    # the reference's own known answers (psitools/test_terminalvelocitysolver.py)
    ref = np.array((0.01000783281810628+0.024248160108983886j, 0.009256948351079859+0.024541272898239125j,
                    0.008859085339836027+0.024878670596489044j))
    np.testing.assert_allclose(w[-3:], ref, rtol=1e-10)
    with pytest.raises(NotImplementedError):
        PSIModeTV(3.0, 0.01, power_law_exponent_size_distribution=-3.7)


def _tv_map_cases():
    gold = load("guessgen.json")["psimodetv"]
    return [(c["mu"], c["tau_max"], c["beta"], c["alpha"]) for c in gold]


def _check_tv_map(tv_roots):
    """Batched form == host class on a map of wave numbers, every fixture family; and == the reference's values at
    the fixture's own wave numbers."""
    from psitools_public_b200 import PSIModeTV
    gold = load("guessgen.json")["psimodetv"]
    ax = np.logspace(-1, 3, 9)
    kx, kz = [a.ravel() for a in np.meshgrid(ax, ax)]
    n_found = 0
    for c in gold:
        tv = PSIModeTV(c["mu"], c["tau_max"], c["beta"])
        tmin, D = c["tau_min"][0], c["alpha"]*20**2
        w, st = tv_roots(tv.fd, tv.tau_max, tv.b, tmin, D, tv.max_it, [c["kx"]], [c["kz"]])
        ref = cplx(c["omega_scalar"])
        assert abs(w[0] - ref) <= 1e-8*abs(ref), (c, w[0], ref)
        tv = PSIModeTV(c["mu"], c["tau_max"], c["beta"], maximum_iterations=40)      # (bounds the give-ups' time)
        w, st = tv_roots(tv.fd, tv.tau_max, tv.b, tmin, D, tv.max_it, kx, kz)
        for i in range(0, len(kx), 3):
            with contextlib.redirect_stdout(io.StringIO()) as said:
                host = complex(tv.find_roots(tmin, kx[i], kz[i], viscous_alpha=c["alpha"]))
            gave_up = "maximum iterations reached" in said.getvalue()
            assert bool(st[i]) == bool(gave_up), (c, kx[i], kz[i], st[i], host, w[i])
            assert abs(w[i] - host) <= 1e-8*max(abs(host), 1e-300), (c, kx[i], kz[i], w[i], host)
            n_found += not gave_up
    assert n_found > 40


def test_tv_map_logic_on_cpu_harness(built):
    import hostsim_util as hs
    _check_tv_map(hs.tv_roots)


@pytest.mark.gpu
def test_tv_map_on_gpu(built):
    from psitools_public_b200 import PSIModeTV, _device
    _check_tv_map(_device.tv_roots)
    # the public call: shapes, status, and that it is the kernel that ran
    tv = PSIModeTV(3.0, 0.1)
    ax = np.logspace(-1, 3, 64)
    kx, kz = np.meshgrid(ax, ax)
    w, st = tv.find_roots_map(1e-8, kx, kz)
    assert w.shape == st.shape == kx.shape and np.all(np.isfinite(w[st == 0])) and (st == 0).mean() > 0.25
    with contextlib.redirect_stdout(io.StringIO()):
        one = complex(tv.find_roots(1e-8, kx[40, 20], kz[40, 20]))
    assert abs(w[40, 20] - one) <= 1e-8*abs(one)
    with pytest.raises(_device.PsiError):
        _device.tv_roots(0.75, 0.1, 0.7, 1e-8, 0.0, 100, [1.0], [1.0])


def _refine_with_tv(tv):
    import time
    from psitools_public_b200 import PSIGridRefiner, TanhSinhNoDeepCopy, psi_mode_mpi
    base = {'__init__': {'stokes_range': [1e-8, 1.0], 'dust_to_gas_ratio': 3.0, 'real_range': [-2.0, 2.0],
                         'imag_range': [1e-8, 1.0], 'n_sample': 20, 'max_zoom_domains': 1, 'tol': 1.0e-13,
                         'clean_tol': 1e-4, 'tanhsinh_integrator': TanhSinhNoDeepCopy()},
            'calculate': {'wave_number_x': None, 'wave_number_z': None, 'guess_roots': []}, 'random_seed': 2}
    with contextlib.redirect_stdout(io.StringIO()):
        gr = PSIGridRefiner('x', baseargs=base, nbase=(3, 3), reruns=1, krange=(1.0, 1.6), tv_guesses=tv,
                            scheduler=psi_mode_mpi.MpiScheduler(time.time(), 3600, verbose=False))
        gr.run_basegrid()
    g = gr.grids[-1]
    return gr, {(iz, ix): g['results'][g['runs'][iz, ix]] for iz in range(g['runs'].shape[0])
                for ix in range(g['runs'].shape[1]) if g['runs'][iz, ix] >= 0}


def _check_refiner_guesses():
    """PSIGridRefiner(tv_guesses=True): the base map's searches get the terminal-velocity mode as a guess; every
    root found without the guesses is found with them too (same values), and the guesses are real modes of the
    approximation inside the search domain."""
    plain, roots0 = _refine_with_tv(False)
    guided, roots1 = _refine_with_tv(True)
    g = guided.grids[0]
    flat, off = guided._terminal_velocity_guesses(g['Kx'].ravel(), g['Kz'].ravel())
    assert off[-1] == len(flat) and len(flat) > 0
    assert np.all((flat.imag > 1e-8) & (flat.imag < 1.0) & (np.abs(flat.real) < 2.0))
    for pix, r0 in roots0.items():
        for z in r0:
            assert any(abs(z - y) < 1e-8 for y in roots1[pix]), (pix, z, roots1[pix])
    assert sum(len(r) > 0 for r in roots1.values()) >= sum(len(r) > 0 for r in roots0.values())
    bad = dict(plain.baseargs)
    bad['__init__'] = dict(bad['__init__'], size_density=object())
    from psitools_public_b200 import PSIGridRefiner
    with pytest.raises(ValueError):
        PSIGridRefiner('x', baseargs=bad, nbase=(3, 3), tv_guesses=True)._terminal_velocity_guesses([1.0], [1.0])


def test_refiner_tv_guesses_on_cpu_harness(hostsim_backend, monkeypatch):
    import hostsim_util as hs
    from psitools_public_b200 import _device
    monkeypatch.setattr(_device, "tv_roots", hs.tv_roots)
    _check_refiner_guesses()


@pytest.mark.gpu
def test_refiner_tv_guesses_on_gpu(built):
    _check_refiner_guesses()
