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

        def __call__(self, x):                      # sizedensity.py:46-48
            return self.f(x)

    spec = load("background.json")["bump"]["spec"]
    mine = product_desc(spec)
    pb = RefStyleBump(spec["amin"], spec["aP"], spec["aL"], spec["aR"], spec["bumpfac"])
    sd = RefStyleDensity(pb.sigma0, [spec["amin"], spec["aR"]], mine.rhod)
    sd.poles = list(mine.poles[:mine.n_poles])
    other, table = _device.describe_size_density(spec["mu"], [spec["amin"], spec["aR"]], 20.0, 3.5, False, sd)
    assert table is None and other.key() == mine.key()


def test_opaque_callable_is_tabulated():
    """Any callable the reference accepts is served: kind TABLE with the quadrature weight as piecewise cubics that
    reproduce the callable to 2e-13 (power law: exactly linear in ln w); a PowerBump SUBCLASS with its own sigma0
    is not mistaken for the stock bump (the analytic route is taken only after a numerical probe)."""
    from psitools_public_b200 import PowerBump, SizeDensity
    sd = SizeDensity(lambda a: a**-0.5, [1e-8, 1.0])
    d, tab = _device.describe_size_density(1.0, [1e-8, 1.0], 20.0, 3.5, False, sd)
    assert d.kind == _device.KIND_TABLE and tab is not None and tab.max_rel_err < 2e-13
    u = np.linspace(np.log(1e-8), 0.0, 1001)[:-1]
    k = np.minimum((u - tab.brk[0])/(tab.brk[1] - tab.brk[0])*tab.n_cells[0], tab.n_cells[0] - 1).astype(int)
    t = (u - tab.brk[0])/(tab.brk[1] - tab.brk[0])*tab.n_cells[0] - k
    c = tab.coef.reshape(-1, 4)[k]
    got = np.exp(((c[:, 3]*t + c[:, 2])*t + c[:, 1])*t + c[:, 0])
    np.testing.assert_allclose(got, tab.weight(u), rtol=1e-13)

    class MyBump(PowerBump):
        def sigma0(self, a):
            return 2.0*PowerBump.sigma0(self, a)*(1 + a)
    pb = MyBump(amin=1e-8, aP=0.1, aL=2.0/3.0*0.1, aR=0.156, bumpfac=2.0)
    sd = SizeDensity(pb.sigma0, [1e-8, 0.156])
    sd.poles = [pb.get_discontinuity()]
    d, tab = _device.describe_size_density(10.0, [1e-8, 0.156], 20.0, 3.5, False, sd)
    assert d.kind == _device.KIND_TABLE and len(tab.n_cells) == 2
    stock = PowerBump(amin=1e-8, aP=0.1, aL=2.0/3.0*0.1, aR=0.156, bumpfac=2.0)
    sd2 = SizeDensity(stock.sigma0, [1e-8, 0.156])
    sd2.poles = [stock.get_discontinuity()]
    d2, tab2 = _device.describe_size_density(10.0, [1e-8, 0.156], 20.0, 3.5, False, sd2)
    assert d2.kind == _device.KIND_POWERBUMP and tab2 is None
