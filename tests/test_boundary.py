"""Public methods of the boundary classes beyond calculate / count_roots, against values produced by
the UNMODIFIED reference (tests/golden/make_golden.py: sec_boundary, boundary.json): an import-swap of a script that
calls them must keep working.  Reference: psi_dispersion.py:124-352, psi_mode.py:354-491, complex_roots.py:540-595,
psi_grid_refine.py:336-386."""
import types

import numpy as np
import pytest

from common import carr, cplx, load


def _ts():
    from psitools_public_b200 import TanhSinhNoDeepCopy
    return TanhSinhNoDeepCopy()


def test_dispersion_matrices(backend):
    from psitools_public_b200 import PSIDispersion
    g = load("boundary.json")
    pd = PSIDispersion(3.0, [1e-8, 1.0], tanhsinh_integrator=_ts())
    w = cplx(g["w"])
    D = pd.calculate(w, g["kx"], g["kz"], g["alpha"])
    assert abs(D - cplx(g["D"])) <= 1e-9*abs(D)
    P = pd.matrix_P(w)
    np.testing.assert_allclose(P, np.array([carr(r) for r in g["P"]]), rtol=1e-13, atol=1e-300)
    taus = np.array(g["taus"])
    A = pd.matrix_inverse_A(w)
    AD = pd.matrix_AD(A, w)
    VAD = pd.matrix_VAD(AD, w)
    W = pd.matrix_W(w)
    for name, mat in (("Ainv", A), ("AD", AD), ("VAD", VAD), ("W", W)):
        for i in range(3):
            for j in range(3):
                got = np.asarray(mat[i][j](taus)) + 0*taus
                np.testing.assert_allclose(got, carr(g[name][i][j]), rtol=1e-9, atol=1e-300, err_msg="%s[%d][%d]" % (name, i, j))
    assert abs(pd.int_j(0).real - g["int_j"][0]) <= 1e-13*abs(g["int_j"][0])
    assert abs(pd.int_j(1).real - g["int_j"][1]) <= 1e-13*abs(g["int_j"][1])
    assert abs(pd.j0 - g["int_j"][0]) <= 1e-13*abs(g["int_j"][0])          # ... and equals the device's K0 value
    a = pd.average(lambda x: x**2/(1 + 1j*x))
    assert abs(a - cplx(g["average_tau2"])) <= 1e-13*abs(a)
    # M from the device equals the reference's integrals of (VAD + W) i/tau -- spot check of one entry on the host
    m11 = pd.average(lambda x: (VAD[0][0](x) + W[0][0](x))*1j/x)
    assert abs(m11 - pd.matrix_M(w)[0, 0]) <= 1e-9*abs(m11)


def test_find_root_and_extra_domain(backend):
    from psitools_public_b200 import PSIMode
    g = load("boundary.json")
    np.random.seed(3)
    pm = PSIMode(3.0, [1e-8, 1.0], [-2, 2], [1e-8, 1], tanhsinh_integrator=_ts())
    roots = pm.calculate(wave_number_x=0.1, wave_number_z=10.0)
    ref_roots = carr(g["mode_roots"])
    assert len(roots) == len(ref_roots) == 1 and abs(roots[0] - ref_roots[0]) <= 1e-8*abs(ref_roots[0])
    fr = g["find_root"]
    n0 = pm.n_function_call
    r1 = pm.find_root(cplx(fr["scalar_start"]), max_iter=50)
    assert np.ndim(r1) == 0 and abs(r1 - cplx(fr["scalar"])) <= 1e-8*abs(r1)
    assert pm.n_function_call - n0 == fr["calls_scalar"]
    n1 = pm.n_function_call
    r2 = pm.find_root(carr(fr["array_start"]), max_iter=30)
    ref2 = carr(fr["array"])
    assert r2.shape == ref2.shape
    for a, b in zip(r2, ref2):                              # sentinel -1j where the reference has it
        assert (a == -1j) == (b == -1j) and abs(a - b) <= 1e-8*abs(b)
    assert pm.n_function_call - n1 == fr["calls_array"]
    n2 = pm.n_function_call
    r3 = pm.find_root(carr(fr["outside_start"]), max_iter=8, select_inside_domain=False)
    ref3 = carr(fr["outside"])
    assert r3.shape == (1,) and (r3[0] == -1j) == (ref3[0] == -1j)
    assert pm.n_function_call - n2 == fr["calls_outside"]
    # add_extra_domain: same random points (global numpy stream, continuing after calculate), D there, counters
    ex = g["extra_domain"]
    n3, m3 = pm.n_function_call, len(pm.z_sample)
    pm.add_extra_domain(extra_domain_size=[0.05, 0.02], centre=ref_roots[0])
    np.testing.assert_array_equal(pm.z_sample[m3:], carr(ex["z_new"]))
    np.testing.assert_allclose(pm.f_sample[m3:], carr(ex["f_new"]), rtol=1e-9)
    assert pm.n_function_call - n3 == ex["calls"] == pm.n_sample


def test_closed_path_passes(backend):
    from psitools_public_b200 import PSIDispersion
    from psitools_public_b200.complex_roots import ClosedPath
    g = load("boundary.json")["closed_path"]
    pd = PSIDispersion(3.0, [1e-8, 1.0], tanhsinh_integrator=_ts())
    f = lambda z: pd.calculate(z, 0.1, 10.0)
    cp = ClosedPath(f, list(carr(g["z_list"])))
    added = [cp.refine_select(tol=0.1, max_step=0.1) for _ in range(3)]
    assert added == g["added"]
    np.testing.assert_array_equal(cp.points, carr(g["points"]))
    np.testing.assert_allclose(cp.f_points, carr(g["f_points"]), rtol=1e-9)
    cp2 = ClosedPath(f, list(carr(g["z_list"])))
    cp2.refine()
    np.testing.assert_array_equal(cp2.points, carr(g["refined_points"]))
    assert list(cp2.interweave(np.array([1.0, 2.0, 3.0]), np.array([10.0, 20.0, 30.0]))) == g["interweave"]
    assert cp.count_roots() == 1                       # and the count can still be finished from there


class _FakeH5Group:
    def __init__(self, name, log):
        self.name, self.attrs, self.log = name, {}, log

    def create_dataset(self, name, data=None):
        self.log.append(("dataset", self.name, name, np.asarray(data).shape, np.asarray(data).dtype.kind))


class _FakeH5File(_FakeH5Group):
    def __init__(self, path, mode):
        _FakeH5Group.__init__(self, "/", [])
        self.path, self.mode, self.groups, self.closed = path, mode, {}, False

    def create_group(self, name):
        g = self.groups[name] = _FakeH5Group(name, self.log)
        return g

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.closed = True


def test_to_hdf5_layout_with_h5py(monkeypatch, tmp_path):
    """The h5py branch of to_hdf5 / write_grid (h5py is absent here and on the GPU box: a 20-line stand-in records
    what is written): groups batchname and batchname/<igrid>, attribute and dataset names of the reference
    (psi_grid_refine.py:336-386)."""
    import sys
    from psitools_public_b200 import PSIGridRefiner
    from psitools_public_b200.psi_grid_refine import RootStore
    files = []

    def File(path, mode):
        files.append(_FakeH5File(path, mode))
        return files[-1]
    monkeypatch.setitem(sys.modules, "h5py", types.SimpleNamespace(File=File))
    gr = PSIGridRefiner.__new__(PSIGridRefiner)
    gr.batchname, gr.datafile_hdf5 = "mybatch", str(tmp_path / "mybatch.hdf5")
    gr.root, gr.wall_start = True, 0.0
    gr.baseargs = {'__init__': {'stokes_range': [1e-8, 0.1], 'dust_to_gas_ratio': 3.0, 'real_range': [-2, 2],
                                'imag_range': [1e-8, 1]}, 'calculate': {'viscous_alpha': 1e-6}}
    kz, kx = np.meshgrid([1.0, 2.0], [3.0, 4.0, 5.0], indexing='ij')
    store = RootStore({0: [0.1 + 0.2j], 1: [], 2: [0.3 + 0.1j, 0.4 + 0.5j]}, width=2)
    gr.grids = [{'Kx': kx, 'Kz': kz, 'runs': np.array([[0, 1, 2], [-1, 2, 0]]), 'results': store, 'nguess': None},
                {'Kx': kx, 'Kz': kz, 'runs': np.array([[0, 1, 2], [-1, 2, 0]]),
                 'results': {0: np.array([0.1 + 0.2j]), 1: np.array([]), 2: np.array([0.3 + 0.1j, 0.4 + 0.5j])},
                 'nguess': None}]
    assert gr.to_hdf5() == gr.datafile_hdf5
    f = files[0]
    assert f.mode == 'w' and f.closed and sorted(f.groups) == ["mybatch", "mybatch/0", "mybatch/1"]
    assert "run_walltime" in f.groups["mybatch"].attrs
    for k in ("mybatch/0", "mybatch/1"):
        assert sorted(f.groups[k].attrs) == ["dust_to_gas_ratio", "imag_range", "real_range", "stokes_range",
                                             "viscous_alpha"]
    names = [(grp, name) for kind, grp, name, _, _ in f.log]
    assert names == [(k, n) for k in ("mybatch/0", "mybatch/1")
                     for n in ("Kxgrid", "Kzgrid", "root_real", "root_imag", "error")]
    shapes = {(grp, name): shp for _, grp, name, shp, _ in f.log}
    assert shapes[("mybatch/0", "root_real")] == (2, 3, 2) == shapes[("mybatch/1", "root_imag")]
    assert shapes[("mybatch/0", "error")] == (2, 3)
    # the array form (RootStore) and the dict form write the same numbers
    a, b = gr.grid_arrays(gr.grids[0]), gr.grid_arrays(gr.grids[1])
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert a["error"].tolist() == [[1, 0, 2], [0, 2, 1]]
