"""Loader for the TEST-ONLY lane-emulation harness (tests/hostsim/hostsim.cpp): the device headers
compiled by g++ with a one-lane warp policy.  Lets the GPU-less container check the kernel logic
against the oracle.  Never imported by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostsim", "hostsim.cpp")
OUT = os.path.join(HERE, "hostsim", "_build", "libpsi_hostsim.so")
CSRC = os.path.join(HERE, "..", "psitools_public_b200", "csrc")


def build(force=False):
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp"))]
    deps.append(os.path.join(HERE, "..", "include", "psi_b200.h"))
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off",
                           "-o", OUT, SRC])
    return OUT


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def background(desc, integ):
    out = np.zeros(5)
    xj = np.ascontiguousarray(integ.xj)
    wj = np.ascontiguousarray(integ.wj)
    lib().hs_background(C.byref(desc), _p(xj), _p(wj), len(xj), int(integ.max_m),
                        C.c_double(float(integ.eps)), _p(out))
    return dict(j0=out[0], j1=out[1], denom=out[2], vgx=out[3], vgy=out[4])


def disp_eval(desc, integ, w, kx, kz, alpha, want_M=False):
    w = np.ascontiguousarray(np.ravel(w), dtype=np.complex128)
    n = len(w)
    kx = np.ascontiguousarray(np.broadcast_to(np.asarray(kx, dtype=float), (n,)))
    kz = np.ascontiguousarray(np.broadcast_to(np.asarray(kz, dtype=float), (n,)))
    al = np.ascontiguousarray(np.broadcast_to(np.asarray(alpha, dtype=float), (n,)))
    D = np.empty(n, dtype=np.complex128)
    M = np.empty((n, 7), dtype=np.complex128)
    ne = C.c_ulonglong(0)
    xj = np.ascontiguousarray(integ.xj)
    wj = np.ascontiguousarray(integ.wj)
    lib().hs_disp_eval(C.byref(desc), _p(xj), _p(wj), len(xj), int(integ.max_m),
                       C.c_double(float(integ.eps)), C.c_longlong(n), _p(kx), _p(kz), _p(al), _p(w),
                       _p(D), _p(M), C.byref(ne))
    return (D, M, ne.value) if want_M else D


class HostsimContext:
    """Test double with the interface of psitools_public_b200._device.Context, backed by the
    lane-emulation harness.  Lets the CPU suite run the product's host-side state machines
    (lock-step engine, PSIMode, contours, grid refinement) end to end without a GPU."""

    def __init__(self, integ):
        self.integ = integ
        self.descs = []
        self.node_evals = 0
        self.launch_count = 0

    def add_dist(self, desc):
        for i, d in enumerate(self.descs):
            if d.key() == desc.key():
                return i
        self.descs.append(desc)
        return len(self.descs) - 1

    def background(self, dist_id):
        return background(self.descs[dist_id], self.integ)

    def disp_eval(self, dist, kx, kz, alpha, w, want_M=False, out=None):
        w = np.ascontiguousarray(np.ravel(w), dtype=np.complex128)
        n = len(w)
        dist = np.broadcast_to(np.asarray(dist, dtype=np.int32), (n,))
        kx = np.broadcast_to(np.asarray(kx, dtype=float), (n,))
        kz = np.broadcast_to(np.asarray(kz, dtype=float), (n,))
        al = np.broadcast_to(np.asarray(alpha, dtype=float), (n,))
        D = np.empty(n, dtype=np.complex128) if out is None else out
        M = np.empty((n, 7), dtype=np.complex128)
        self.launch_count += 1
        for d in np.unique(dist):
            sel = (dist == d).nonzero()[0]
            Dd, Md, ne = disp_eval(self.descs[int(d)], self.integ, w[sel], kx[sel], kz[sel], al[sel], True)
            D[sel] = Dd
            M[sel] = Md
            self.node_evals += ne
        return (D, M) if want_M else D
