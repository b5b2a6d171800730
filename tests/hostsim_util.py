"""Loader for the TEST-ONLY lane-emulation harness (tests/hostsim/hostsim.cpp): the device headers
compiled by g++ with a one-lane warp policy.  Lets the GPU-less container check the kernel logic
against the oracle.  Never imported by the product package."""
import ctypes as C
import os
import subprocess

import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostsim", "hostsim.cpp")
OUT = os.path.join(HERE, "hostsim", "_build", "libpsi_hostsim.so")
CSRC = os.path.join(HERE, "..", "psitools_public_b200", "csrc")
sys.path.insert(0, os.path.join(HERE, ".."))


def build(force=False):
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)
                    if f.endswith((".cuh", ".hpp")) or f == "psi_engine.cu"]
    deps.append(os.path.join(HERE, "..", "include", "psi_b200.h"))
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-fopenmp",
                           "-x", "c++", "-o", OUT, SRC, "-ldl", "-lpthread"])
    return OUT


_lib = None


def lib():
    global _lib
    if _lib is None:
        # PSI_HOSTSIM_LIB: an alternative build of hostsim.cpp, e.g. with -fsanitize=address (tools/asan_hostsim.sh)
        _lib = C.CDLL(os.environ.get("PSI_HOSTSIM_LIB") or build())
        _lib.hs_ctx_create.restype = C.c_void_p
        _lib.hs_last_error.restype = C.c_char_p
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


from psitools_public_b200._device import EngineAPI, PsiError  # noqa: E402


class HostsimContext(EngineAPI):
    """Test double with the interface of psitools_public_b200._device.Context, backed by the
    lane-emulation harness.  Lets the CPU suite run the product's host-side state machines and the
    native lock-step engine end to end without a GPU."""

    def __init__(self, integ):
        self.integ = integ
        self.lib = lib()
        xj = np.ascontiguousarray(integ.xj, dtype=np.float64)
        wj = np.ascontiguousarray(integ.wj, dtype=np.float64)
        self.handle = C.c_void_p(self.lib.hs_ctx_create(_p(xj), _p(wj), len(xj), int(integ.max_m),
                                                        C.c_double(float(integ.eps))))
        self._dists = {}
        self.node_evals = 0
        self.launch_count = 0

    def _check(self, rc, what):
        if rc < 0:
            raise PsiError("%s failed (code %d): %s" % (what, rc, self._last_error()))
        return rc

    _exit_rule = "reference"

    @property
    def exit_rule(self):
        return self._exit_rule

    @exit_rule.setter
    def exit_rule(self, rule):
        assert rule in ("predicted", "reference", "fast")
        self.lib.hs_set_exit_rule(self.handle, {"reference": 0, "predicted": 1, "fast": 2}[rule])
        self._exit_rule = rule

    def set_level_cap(self, cap):
        """Fast-path emulation: with exit_rule 'fast' and a cap the one-lane code is exactly what one thread
        of disp_eval_thread_kernel runs (NaN = not converged within the cap)."""
        self.lib.hs_set_level_cap(self.handle, int(cap))

    def _last_error(self):
        return self.lib.hs_last_error().decode()

    def add_dist(self, desc, table=None):
        k = desc.key() + (table.key() if table is not None else b"")
        if k not in self._dists:
            if table is None:
                self._dists[k] = self.lib.hs_dist_add(self.handle, C.byref(desc))
            else:
                self._dists[k] = self.lib.hs_dist_add_table(self.handle, C.byref(desc), C.byref(table.struct()))
        return self._dists[k]

    def background(self, dist_id):
        out = np.zeros(5)
        self.lib.hs_dist_background(self.handle, int(dist_id), _p(out))
        return dict(j0=out[0], j1=out[1], denom=out[2], vgx=out[3], vgy=out[4])

    def disp_eval(self, dist, kx, kz, alpha, w, want_M=False, out=None):
        w = np.ascontiguousarray(np.ravel(w), dtype=np.complex128)
        n = len(w)
        dist = np.ascontiguousarray(np.broadcast_to(np.asarray(dist, dtype=np.int32), (n,)))
        kx = np.ascontiguousarray(np.broadcast_to(np.asarray(kx, dtype=float), (n,)))
        kz = np.ascontiguousarray(np.broadcast_to(np.asarray(kz, dtype=float), (n,)))
        al = np.ascontiguousarray(np.broadcast_to(np.asarray(alpha, dtype=float), (n,)))
        D = np.empty(n, dtype=np.complex128) if out is None else out
        M = np.empty((n, 7), dtype=np.complex128)
        ne = C.c_ulonglong(0)
        self.launch_count += 1
        self.lib.hs_disp_eval(self.handle, C.c_longlong(n), _p(dist), _p(kx), _p(kz), _p(al), _p(w), _p(D),
                              _p(M), C.byref(ne))
        self.node_evals += ne.value
        return (D, M) if want_M else D


def background(desc, integ):
    ctx = HostsimContext(integ)
    return ctx.background(ctx.add_dist(desc))


def disp_eval(desc, integ, w, kx, kz, alpha, want_M=False):
    ctx = HostsimContext(integ)
    res = ctx.disp_eval(ctx.add_dist(desc), kx, kz, alpha, w, want_M=True)
    return (res[0], res[1], ctx.node_evals) if want_M else res[0]


def tv_roots(fd, tau_max, b, tau_min, diffusion, max_it, kx, kz, device=None):
    """Lane-emulation form of _device.tv_roots (csrc/psi_tv.cuh compiled by g++)."""
    kx = np.ascontiguousarray(kx, dtype=np.float64)
    kz = np.ascontiguousarray(kz, dtype=np.float64)
    w = np.zeros(len(kx), dtype=np.complex128)
    st = np.zeros(len(kx), dtype=np.int32)
    lib().hs_tv_roots(C.c_double(fd), C.c_double(tau_max), C.c_double(b), C.c_double(tau_min), C.c_double(diffusion),
                      C.c_int(max_it), C.c_longlong(len(kx)), _p(kx), _p(kz), _p(w), _p(st))
    return w, st
