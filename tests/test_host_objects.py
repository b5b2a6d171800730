"""Host-side API objects (TanhSinh, SizeDensity, PowerBump*) against reference fixtures, and the
descriptor translation, including duck-typed reference-style objects."""
import copy
import hashlib

import numpy as np

from common import load, product_desc, product_kwargs
from psitools_public_b200 import TanhSinh, TanhSinhNoDeepCopy, _device


def test_tanhsinh_tables_bit_identical():
    g = load("nodes.json")
    for key in ("15_12", "15_10", "15_13", "12_8"):
        p, m = map(int, key.split("_"))
        t = TanhSinh(p, m)
        assert hashlib.sha256(np.ascontiguousarray(t.xj).tobytes()).hexdigest() == g[key]["sha_x"]
        assert hashlib.sha256(np.ascontiguousarray(t.wj).tobytes()).hexdigest() == g[key]["sha_w"]


def test_tanhsinh_integrate_and_nodeepcopy():
    t = TanhSinhNoDeepCopy()
    assert copy.deepcopy(t) is t
    g = load("nodes.json")["known"]
    assert abs(t.integrate(lambda x: x*np.log(1 + x), 0.0, 1.0) - g["xlog1px"]) < 1e-15
    assert abs(t.integrate(lambda x: np.log(x)**2, 0.0, 1.0) - g["log2x"]) < 1e-14


def test_size_density_probes():
    for name, g in load("background.json").items():
        kw = product_kwargs(g["spec"])
        sd = kw.get("size_density")
        if sd is None:
            continue
        assert abs(sd.rhod - g["rhod"]) <= 1e-15*g["rhod"], name
        np.testing.assert_allclose(sd(np.array(g["s_probe"])), g["F_probe"], rtol=1e-14)
        assert list(sd.poles) == g["poles"], name


def test_descriptor_fields():
    g = load("background.json")
    d = product_desc(g["mrn_1e-2"]["spec"])
    assert d.kind == _device.KIND_POWERLAW and d.q == -0.5 and d.n_poles == 0
    d = product_desc(g["bumptail"]["spec"])
    assert d.kind == _device.KIND_POWERBUMPTAIL and d.aBT == 1e-6 and d.n_poles == 1
    d = product_desc(g["bump_eps"]["spec"])
    assert d.kind == _device.KIND_POWERBUMP and d.scale != 1.0


def test_descriptor_from_reference_style_objects():
    """A SizeDensity/PowerBump written the way the reference writes them (constants only inside
    lambda closures, reference power_bump.py:75-80, sizedensity.py:41) is recognised too."""
    class RefStyleBump:
        def __init__(self, amin, aP, aL, aR, bumpfac=2.0, beta=-3.5):
            self.amin, self.aP, self.aL, self.aR = amin, aP, aL, aR
            sigma = np.max((np.min((abs(aR - aP), abs(aL - aP)))/np.sqrt(np.log(2.0)), 0.1*aP))
            fac_fnn = aP**(-beta)
            self.fnn = lambda a: fac_fnn*a**(beta)
            fac_b = bumpfac*self.fnn(aL)
            fac_b2 = -1/sigma**2
            self.b = lambda a: fac_b*np.exp(fac_b2*(a - aP)**2)
            self.epstot = None

        def sigma0(self, a):
            return np.where(a > self.aP, self.b(a), np.maximum(self.fnn(a), self.b(a)))*a**3

    class RefStyleDensity:
        def __init__(self, sigma, size_range, rhod):
            self.amin, self.amax, self.rhod = size_range[0], size_range[1], rhod
            self.f = lambda x: self.amax*sigma(self.amax*x)/self.rhod
            self.poles = []

    spec = load("background.json")["bump"]["spec"]
    mine = product_desc(spec)
    pb = RefStyleBump(spec["amin"], spec["aP"], spec["aL"], spec["aR"], spec["bumpfac"])
    sd = RefStyleDensity(pb.sigma0, [spec["amin"], spec["aR"]], mine.rhod)
    sd.poles = list(mine.poles[:mine.n_poles])
    other = _device.describe_size_density(spec["mu"], [spec["amin"], spec["aR"]], 20.0, 3.5, False, sd)
    assert other.key() == mine.key()


def test_opaque_callable_rejected():
    from psitools_public_b200 import SizeDensity
    import pytest
    sd = SizeDensity(lambda a: a**-0.5, [1e-8, 1.0])
    with pytest.raises(NotImplementedError):
        _device.describe_size_density(1.0, [1e-8, 1.0], 20.0, 3.5, False, sd)
