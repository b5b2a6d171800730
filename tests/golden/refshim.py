"""Import shim for the UNMODIFIED reference (psitools) in the build container.

Test infrastructure only.  /root/reference does not exist on the GPU box, so nothing that
runs there (``-m gpu`` tests, smoke(), bench.py) may import this module; it is used by
``make_golden.py`` (fixture generation) and by the optional ``refcheck`` tests that skip
when the reference tree is absent.

Why a shim is needed:
  * psitools/__init__.py:20-30 imports mpi4py and h5py, both absent here -> register an
    empty ``psitools`` package whose __path__ points at the reference and stub the two
    modules;
  * numpy >= 1.24 removed ``np.int`` which complex_roots.py:159 and psi_grid_refine.py use;
  * the MPI master/worker loop (psi_mode_mpi.py:57-160) needs >1 rank; ``SerialScheduler``
    below maps the scheduler's own ``runcompute`` bodies over the arg list instead.
No arithmetic of the reference is changed.
"""
import os
import sys
import types
import importlib

REF_ROOT = os.environ.get("PSI_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "psitools"))


class _FakeComm:
    size = 2

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 2

    def bcast(self, obj, root=0):
        return obj


def load():
    """Return a namespace with the reference hot-path modules."""
    if not available():
        raise RuntimeError("reference tree not present at " + REF_ROOT)
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int
    if "mpi4py" not in sys.modules:
        m = types.ModuleType("mpi4py")
        mpi = types.ModuleType("mpi4py.MPI")
        mpi.COMM_WORLD = _FakeComm()
        mpi.ANY_SOURCE = -1
        mpi.Status = object
        mpi.Finalize = lambda: None
        m.MPI = mpi
        sys.modules["mpi4py"] = m
        sys.modules["mpi4py.MPI"] = mpi
    if "h5py" not in sys.modules:
        sys.modules["h5py"] = types.ModuleType("h5py")
    if "psitools" not in sys.modules:
        pkg = types.ModuleType("psitools")
        pkg.__path__ = [os.path.join(REF_ROOT, "psitools")]
        sys.modules["psitools"] = pkg
    ns = types.SimpleNamespace()
    for name in ("tanhsinh", "sizedensity", "power_bump", "psi_dispersion", "complex_roots",
                 "psi_mode", "psi_mode_mpi", "complex_roots_mpi", "psi_grid_refine"):
        setattr(ns, name, importlib.import_module("psitools." + name))
    return ns


_ITEMS = []     # arg list shared with forked workers (arg dicts hold unpicklable lambdas)


def _run_index(i):
    return _run_item(_ITEMS[i])


def _run_item(args):
    """Worker body: one scheduler item through the reference's own runcompute."""
    kind, item = args
    ns = load()
    if kind == "mode":
        s = ns.psi_mode_mpi.MpiScheduler.__new__(ns.psi_mode_mpi.MpiScheduler)
        s.verbose = False
        s.rank = 1
        s.PSIModeArgs = {}
        return ns.psi_mode_mpi.MpiScheduler.runcompute(s, item)
    if kind == "disp":
        s = ns.psi_mode_mpi.DispersionRelationMpiScheduler.__new__(
            ns.psi_mode_mpi.DispersionRelationMpiScheduler)
        s.verbose = False
        s.PSIModeArgs = {}
        return ns.psi_mode_mpi.DispersionRelationMpiScheduler.runcompute(s, item)
    if kind == "count":
        s = ns.complex_roots_mpi.MpiScheduler.__new__(ns.complex_roots_mpi.MpiScheduler)
        s.verbose = False
        s.PSIDispersionArgs = {}
        return ns.complex_roots_mpi.MpiScheduler.runcompute(s, item)
    raise ValueError(kind)


class SerialScheduler:
    """Stand-in for MpiScheduler.run: same per-item work and seeding (psi_mode_mpi.py:162-178),
    mapped over a process pool instead of MPI ranks."""

    def __init__(self, kind="mode", procs=1):
        self.kind = kind
        self.procs = procs

    def run(self, arglist):
        import copy
        items = [(self.kind, copy.deepcopy(a)) for a in arglist]
        if self.procs > 1:
            import multiprocessing as mp
            global _ITEMS
            _ITEMS = items
            with mp.get_context("fork").Pool(self.procs) as pool:
                res = pool.map(_run_index, range(len(items)), chunksize=1)
        else:
            res = [_run_item(it) for it in items]
        return {i: r for i, r in enumerate(res)}
