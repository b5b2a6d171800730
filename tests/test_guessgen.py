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
