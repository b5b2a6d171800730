"""ctypes binding of libpsi_b200.so (C ABI in include/psi_b200.h) and the host-side translation of
the reference's Python objects (TanhSinh, SizeDensity, PowerBump*) into device descriptors.

There is deliberately NO CPU fallback: if the library is missing or no CUDA device is visible every
entry point raises.
"""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PSI_B200_LIB", os.path.join(_HERE, "libpsi_b200.so"))   # override: experiments only

PSI_MAX_POLES = 4
MODE_MAX_SAMPLES = 512            # PSI_MODE_MAX_SAMPLES of include/psi_b200.h
KIND_POWERLAW, KIND_POWERBUMP, KIND_POWERBUMPTAIL, KIND_TABLE = 0, 1, 2, 3
PSI_MAX_TABLE_SEGS = PSI_MAX_POLES + 1


class PsiError(RuntimeError):
    """An error code returned by libpsi_b200."""


class DistDesc(C.Structure):
    """Mirror of ``psi_dist_desc`` (include/psi_b200.h)."""
    _fields_ = [("kind", C.c_int), ("single_size", C.c_int),
                ("mu", C.c_double), ("tau_min", C.c_double), ("tau_max", C.c_double),
                ("sound_speed", C.c_double), ("rhod", C.c_double), ("q", C.c_double),
                ("beta", C.c_double), ("norm_pl", C.c_double), ("amp", C.c_double),
                ("curv", C.c_double), ("aL", C.c_double), ("aP", C.c_double), ("aBT", C.c_double),
                ("scale", C.c_double), ("n_poles", C.c_int), ("poles", C.c_double*PSI_MAX_POLES)]

    def key(self):
        return bytes(self)


class TableDescStruct(C.Structure):
    """Mirror of ``psi_table_desc`` (include/psi_b200.h)."""
    _fields_ = [("n_seg", C.c_int), ("brk", C.c_double*(PSI_MAX_TABLE_SEGS + 1)),
                ("n_cells", C.c_int*PSI_MAX_TABLE_SEGS), ("log_mode", C.c_int*PSI_MAX_TABLE_SEGS),
                ("coef", C.c_void_p)]


class SizeDensityTable:
    """Host side of a tabulated size density (C ABI psi_dist_add_table): the quadrature weight w(u) = e^u F(e^u),
    u = ln s, of ANY callable the reference accepts (SizeDensity(sigma, size_range), sizedensity.py:25-48), as
    piecewise cubics on uniform grids between the kinks of sigma (``poles``, in units of stopping time).

    Per cell the cubic interpolates w -- ln w wherever w > 0 on the whole piece: exact for power laws, and it keeps
    the RELATIVE error uniform over the eight decades of s -- at the four Chebyshev points of the cell.  The cell
    count starts at ``cells`` per piece and doubles until the interpolant reproduces the callable to ``rtol`` at
    random probe points (sigma is called with arrays, as the reference's quadrature does)."""

    def __init__(self, size_density, tau_min, tau_max, poles, cells=16384, rtol=2e-13, max_cells=1 << 21):
        self.f = size_density
        smin = tau_min/tau_max
        brk = [np.log(smin)] + sorted(np.log(p/tau_max) for p in poles if tau_min < p < tau_max) + [0.0]
        if len(brk) - 1 > PSI_MAX_TABLE_SEGS:
            raise ValueError("at most %d size-density poles are supported" % PSI_MAX_POLES)
        self.brk = np.array(brk)
        self.n_cells, self.log_mode, coef = [], [], []
        self.max_rel_err = 0.0
        rng = np.random.RandomState(12345)
        for a, b in zip(brk[:-1], brk[1:]):
            n = max(16, int(cells*(b - a)/(brk[-1] - brk[0])) + 1)
            while True:
                c, logm = self._fit(a, b, n)
                err = self._probe(a, b, n, c, logm, rng)
                if err <= rtol or 2*n > max_cells:
                    break
                n *= 2
            self.max_rel_err = max(self.max_rel_err, err)
            self.n_cells.append(n)
            self.log_mode.append(int(logm))
            coef.append(c)
        self.coef = np.ascontiguousarray(np.concatenate(coef), dtype=np.float64)

    def weight(self, u):
        """w(u) = e^u F(e^u) from the callable."""
        s = np.exp(u)
        return s*np.asarray(self.f(s), dtype=np.float64)

    _T = 0.5 - 0.5*np.cos((2*np.arange(4) + 1)*np.pi/8)          # Chebyshev points of [0, 1]
    _VINV = np.linalg.inv(np.vander(_T, 4, increasing=True))       # values at _T -> monomial coefficients

    def _fit(self, a, b, n):
        h = (b - a)/n
        u = a + h*(np.arange(n)[:, None] + self._T[None, :])
        w = self.weight(u.ravel()).reshape(n, 4)
        logm = bool(np.all(w > 0) and np.all(np.isfinite(w)))
        v = np.log(w) if logm else w
        return v @ self._VINV.T, logm

    def _probe(self, a, b, n, c, logm, rng):
        t = rng.uniform(0, 1, 4096)
        i = rng.randint(0, n, 4096)
        u = a + (b - a)/n*(i + t)
        v = ((c[i, 3]*t + c[i, 2])*t + c[i, 1])*t + c[i, 0]
        got = np.exp(v) if logm else v
        ref = self.weight(u)
        scale = np.abs(ref) if logm else np.maximum(np.abs(ref), 1e-300 + np.max(np.abs(ref))*1e-6)
        return float(np.max(np.abs(got - ref)/scale))

    def key(self):
        return self.brk.tobytes() + self.coef.tobytes()

    def struct(self):
        t = TableDescStruct()
        t.n_seg = len(self.n_cells)
        for k, v in enumerate(self.brk):
            t.brk[k] = v
        for k, (n, m) in enumerate(zip(self.n_cells, self.log_mode)):
            t.n_cells[k], t.log_mode[k] = n, m
        t.coef = self.coef.ctypes.data
        return t


_lib = None
_lib_lock = threading.Lock()

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p

_SIGNATURES = {
    # name: (restype, argtypes)
    "psi_version": (C.c_char_p, []),
    "psi_last_error": (C.c_char_p, []),
    "psi_device_count": (C.c_int, []),
    "psi_ctx_create": (C.c_int, [C.POINTER(_vp), C.c_int, _vp, _vp, C.c_int, C.c_int, C.c_double]),
    "psi_ctx_destroy": (None, [_vp]),
    "psi_ctx_sm_count": (C.c_int, [_vp]),
    "psi_ctx_set_exit_rule": (C.c_int, [_vp, C.c_int]),
    "psi_ctx_get_exit_rule": (C.c_int, [_vp]),
    "psi_ctx_synchronize": (C.c_int, [_vp]),
    "psi_ctx_launch_count": (C.c_longlong, [_vp]),
    "psi_dist_add": (C.c_int, [_vp, C.POINTER(DistDesc)]),
    "psi_dist_add_table": (C.c_int, [_vp, C.POINTER(DistDesc), C.POINTER(TableDescStruct)]),
    "psi_dist_get_background": (C.c_int, [_vp, C.c_int, _dp]),
    "psi_disp_eval_host": (C.c_int, [_vp, C.c_longlong, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                     C.POINTER(C.c_ulonglong)]),
    "psi_disp_eval_dev": (C.c_int, [_vp, C.c_longlong, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                    _vp, C.c_uint, _vp]),
    "psi_point_class": (C.c_int, [_vp, C.c_int, C.c_double]),
    "psi_measure_fp64_peak": (C.c_int, [_vp, C.c_int, _dp]),
    "psi_engine_bind_lapack": (C.c_int, [C.c_char_p, C.c_char_p]),
    "psi_engine_counters": (None, [_vp]),
    "psi_engine_set_threads": (None, [C.c_int]),
    "psi_mode_batch_host": (C.c_int, [_vp, C.c_int] + [_vp]*10 + [C.c_int] + [_vp]*5),
    "psi_mode_batch_device": (C.c_int, [_vp, C.c_int] + [_vp]*10 + [C.c_int] + [_vp]*5),
    "psi_mode_batch_device_rec": (C.c_int, [_vp, C.c_int] + [_vp]*10 + [C.c_int, C.c_int, C.POINTER(_vp), _vp]),
    "psi_mode_search_detail": (C.c_int, [_vp, C.c_int, C.c_double, C.c_double, C.c_double, _vp, _vp, _vp, _vp,
                                         C.c_int, _vp, C.c_int, C.c_int] + [_vp]*10),
    "psi_contour_batch_host": (C.c_int, [_vp, C.c_int] + [_vp]*13),
    "psi_polish_batch_host": (C.c_int, [_vp, C.c_int] + [_vp]*12),
    "psi_refine_open": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_longlong,
                                  C.c_longlong, C.c_int, _vp, _vp]),
    "psi_refine_sweep": (C.c_int, [_vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "psi_refine_fetch": (C.c_int, [_vp, _vp, _vp, _vp]),
    "psi_refine_update": (C.c_int, [_vp, C.c_longlong, _vp, _vp, _vp, C.c_longlong, _vp, _vp]),
    "psi_refine_close": (None, [_vp]),
    "psi_tv_roots": (C.c_int, [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                               C.c_longlong, _vp, _vp, _vp, _vp]),
}


def exported_symbols():
    """Names every build of the library must export (checked by the CPU test-suite)."""
    return sorted(_SIGNATURES)


def load_library():
    """dlopen libpsi_b200.so and declare the prototypes.  Raises if the library is not built."""
    global _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise PsiError("libpsi_b200.so not built (run `python -c 'import __graft_entry__ as g; "
                               "g.build()'`); there is no CPU fallback")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(rc, what=""):
    if rc < 0:
        msg = load_library().psi_last_error().decode()
        raise PsiError("%s failed (code %d): %s" % (what or "libpsi_b200 call", rc, msg))
    return rc


def default_device():
    if "PSI_B200_DEVICE" in os.environ:
        return int(os.environ["PSI_B200_DEVICE"])
    return int(os.environ.get("LOCAL_RANK", "0"))


def _ptr(a):
    return a.ctypes.data_as(_vp)


def openblas_paths():
    """The OpenBLAS shared objects numpy (ILP64) and scipy (LP64) ship and load themselves."""
    import glob
    import numpy
    import scipy
    out = []
    for mod, pat in ((numpy, "libscipy_openblas64_*.so"), (scipy, "libscipy_openblas-*.so")):
        base = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(mod.__file__))),
                            mod.__name__ + ".libs")
        hits = sorted(glob.glob(os.path.join(base, pat)))
        if not hits:
            raise PsiError("OpenBLAS library %s not found under %s" % (pat, base))
        out.append(hits[0])
    return out


def _checked_seeds(seeds, n):
    """uint32 seeds; out-of-range values raise ValueError like numpy.random.seed (they must not wrap silently:
    the device stream would then differ from the reference's)."""
    s = np.broadcast_to(np.asarray(seeds), (n,))
    if s.size and (np.any(s < 0) or np.any(s > 0xffffffff)):
        raise ValueError("Seed must be between 0 and 2**32 - 1")
    return np.ascontiguousarray(s, dtype=np.uint32)


class DeviceRecords:
    """Packed result records of a K4 batch that are still in device memory (psi_mode_batch_device_rec): ``rows``
    rows of ``width`` = 4 + 2*max_roots doubles {n_roots, status, n_function_call, 0, roots...}.  Exposes
    ``__cuda_array_interface__`` so that torch (or cupy) can wrap the library's buffer without a copy; valid until
    the next call on the same context."""

    def __init__(self, ptr, rows, width, n_items, max_roots):
        self.ptr, self.rows, self.width, self.n_items, self.max_roots = int(ptr), int(rows), int(width), int(n_items), int(max_roots)
        self.__cuda_array_interface__ = {"shape": (self.rows, self.width), "typestr": "<f8",
                                         "data": (self.ptr, False), "version": 2, "strides": None}

    @staticmethod
    def unpack(table, max_roots):
        """(root matrix, root counts, status, n_function_call) of a host copy of record rows."""
        table = np.asarray(table, dtype=np.float64).reshape(-1, 4 + 2*max_roots)
        roots = np.ascontiguousarray(table[:, 4:]).view(np.complex128)
        return roots, table[:, 0].astype(np.int64), table[:, 1].astype(np.int64), table[:, 2].astype(np.int64)


class EngineAPI:
    """Python face of the native lock-step engine (csrc/psi_engine.cu); mixed into Context.  Needs
    ``self.lib`` (a CDLL exporting the psi_* engine entry points) and ``self.handle``."""

    _lapack_bound = False

    @staticmethod
    def default_engine_threads():
        """Host threads per process: all cores, shared fairly between the ranks of one node (torchrun
        sets OMP_NUM_THREADS=1, which would serialise the engine's AAA phase)."""
        if "PSI_B200_THREADS" in os.environ:
            return max(1, int(os.environ["PSI_B200_THREADS"]))
        ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
        return max(1, (os.cpu_count() or 1)//ranks)

    def _bind_lapack(self):
        if not self._lapack_bound:
            self.lib.psi_engine_set_threads(C.c_int(self.default_engine_threads()))
            a, b = openblas_paths()
            fn = self.lib.psi_engine_bind_lapack
            fn.restype, fn.argtypes = C.c_int, [C.c_char_p, C.c_char_p]
            self._check(fn(a.encode(), b.encode()), "psi_engine_bind_lapack")
            type(self)._lapack_bound = True

    def _check(self, rc, what):
        return check(rc, what)

    def mode_batch(self, dist, kx, kz, alpha, domain, n_sample, max_zoom, max_secant, tol, clean_tol,
                   seeds, guesses, max_roots=8, device_aaa=False, dense=False):
        """``_mode_batch`` in chunks: the K4 pipeline keeps every search's samples and work lists resident
        (~50 B per sample slot), so a multi-million-item batch with guess domains is cut into pieces that
        fit ``PSI_B200_BATCH_BYTES`` (default 48 GiB of the 180 GB).  Items are independent: results do not
        depend on the cut."""
        n = len(dist)
        ns = np.broadcast_to(np.asarray(n_sample, dtype=np.int64), (n,))
        mz = np.broadcast_to(np.asarray(max_zoom, dtype=np.int64), (n,))
        if guesses is None:
            goff = np.zeros(n + 1, dtype=np.int64)
        elif isinstance(guesses, tuple):
            goff = np.asarray(guesses[1], dtype=np.int64)
        else:
            goff = np.concatenate(([0], np.cumsum([len(g) for g in guesses]))).astype(np.int64)
        ng = goff[1:] - goff[:-1]
        own = ns*(1 + ng + 4*mz)                       # samples a search can reach
        if device_aaa and n and int(own.max()) > MODE_MAX_SAMPLES:
            return self._mode_batch_mixed(own > MODE_MAX_SAMPLES, dist, kx, kz, alpha, domain, n_sample, max_zoom,
                                          max_secant, tol, clean_tol, seeds, guesses, goff, max_roots, dense)
        cap = int(own.max()) if n else 0
        per_item = 50*cap + 8*int(np.max(ns*np.maximum(1 + ng, 4))) + 16*max_roots + 256 if n else 1
        budget = int(os.environ.get("PSI_B200_BATCH_BYTES", str(48 << 30)))
        chunk = max(1, budget//per_item)
        if not device_aaa or n <= chunk:
            return self._mode_batch(dist, kx, kz, alpha, domain, n_sample, max_zoom, max_secant, tol,
                                    clean_tol, seeds, guesses, max_roots, device_aaa, dense)
        per = lambda a, lo, hi: np.broadcast_to(np.asarray(a), (n,) + np.shape(a)[1:])[lo:hi]
        domain = np.asarray(domain, dtype=float).reshape(n, 4)
        if guesses is not None and not isinstance(guesses, tuple):
            guesses = (np.array([complex(z) for g in guesses for z in g], dtype=np.complex128), goff)
        parts = []
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            g = None if guesses is None else (guesses[0][goff[lo]:goff[hi]], goff[lo:hi + 1] - goff[lo])
            parts.append(self._mode_batch(per(dist, lo, hi), per(kx, lo, hi), per(kz, lo, hi), per(alpha, lo, hi),
                                          domain[lo:hi], per(n_sample, lo, hi), per(max_zoom, lo, hi),
                                          per(max_secant, lo, hi), per(tol, lo, hi), per(clean_tol, lo, hi),
                                          per(seeds, lo, hi), g, max_roots, device_aaa, True))
        width = max(p[0][0].shape[1] for p in parts)
        roots = np.zeros((n, width), dtype=np.complex128)
        at = 0
        for (r, _), _, _ in parts:
            roots[at:at + len(r), :r.shape[1]] = r
            at += len(r)
        nr = np.concatenate([p[0][1] for p in parts])
        nc = np.concatenate([p[1] for p in parts])
        st = dict(parts[0][2])
        for p in parts[1:]:
            for k, v in p[2].items():
                if isinstance(v, dict):
                    st[k] = {kk: st[k][kk] + vv for kk, vv in v.items()}
                elif k != "passes":
                    st[k] = st[k] + v
        st["chunks"] = len(parts)
        if dense:
            return (roots, nr), nc, st
        return [roots[i, :nr[i]].copy() for i in range(n)], nc, st

    def _mode_batch_mixed(self, big, dist, kx, kz, alpha, domain, n_sample, max_zoom, max_secant, tol, clean_tol,
                          seeds, guesses, goff, max_roots, dense):
        """A batch in which SOME searches can exceed the device pipeline's 512 samples (a refinement pixel with many
        neighbour roots as guesses): those go to the host-LAPACK engine, all others stay on the device."""
        n = len(big)
        per = lambda a, sel: np.broadcast_to(np.asarray(a), (n,) + np.shape(a)[1:])[sel]
        domain = np.asarray(domain, dtype=float).reshape(n, 4)
        if guesses is not None and not isinstance(guesses, tuple):
            guesses = (np.array([complex(z) for g in guesses for z in g], dtype=np.complex128), goff)
        out_r, out_n, out_c, stats = {}, np.zeros(n, np.int64), np.zeros(n, np.int64), None
        for sel, dev in ((np.flatnonzero(~big), True), (np.flatnonzero(big), False)):
            if not len(sel):
                continue
            g = None
            if guesses is not None:
                length = (goff[1:] - goff[:-1])[sel]
                off = np.zeros(len(sel) + 1, dtype=np.int64)
                np.cumsum(length, out=off[1:])
                take = np.repeat(goff[:-1][sel] - off[:-1], length) + np.arange(off[-1])
                g = (np.asarray(guesses[0])[take], off)
            (r, nr), nc, st = self.mode_batch(per(dist, sel), per(kx, sel), per(kz, sel), per(alpha, sel), domain[sel],
                                              per(n_sample, sel), per(max_zoom, sel), per(max_secant, sel),
                                              per(tol, sel), per(clean_tol, sel), per(seeds, sel), g, max_roots,
                                              device_aaa=dev, dense=True)
            out_r[dev] = (sel, r)
            out_n[sel], out_c[sel] = nr, nc
            if stats is None:
                stats = dict(st)
            else:
                stats["evals"] = stats.get("evals", 0) + st.get("evals", 0)
        stats["host_engine_items"] = int(np.count_nonzero(big))
        width = max(r.shape[1] for _, r in out_r.values())
        roots = np.zeros((n, width), dtype=np.complex128)
        for sel, r in out_r.values():
            roots[sel, :r.shape[1]] = r
        if dense:
            return (roots, out_n), out_c, stats
        return [roots[i, :out_n[i]].copy() for i in range(n)], out_c, stats

    def _mode_batch(self, dist, kx, kz, alpha, domain, n_sample, max_zoom, max_secant, tol, clean_tol,
                    seeds, guesses, max_roots=8, device_aaa=False, dense=False):
        """Run len(dist) PSIMode.calculate searches.  Per-item arrays; ``guesses`` is a list of
        complex sequences or the CSR pair (flat values, n+1 offsets).  ``device_aaa=False``: native lock-step engine (AAA with LAPACK on the host
        cores, D and the secant polish on the GPU); ``True``: kernel K4, everything on the device.
        Returns (list of root arrays, n_function_call array, stats); with ``dense=True`` the roots come
        back as an (n, max_roots) complex matrix plus the per-item root counts instead of a list."""
        if not device_aaa:
            self._bind_lapack()
        n = len(dist)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        dist, kx, kz, alpha = i32(dist), f64(kx), f64(kz), f64(alpha)
        domain = f64(domain).reshape(n, 4)
        ip = i32(np.stack([np.broadcast_to(n_sample, (n,)), np.broadcast_to(max_zoom, (n,)),
                           np.broadcast_to(max_secant, (n,))], axis=1))
        dp = f64(np.stack([np.broadcast_to(tol, (n,)), np.broadcast_to(clean_tol, (n,))], axis=1))
        seeds = _checked_seeds(seeds, n)
        goff = np.zeros(n + 1, dtype=np.int32)
        if guesses is None:
            flat = []
        elif isinstance(guesses, tuple):          # CSR form: (flat complex array, n+1 offsets)
            flat = np.asarray(guesses[0], dtype=np.complex128)
            goff[:] = np.asarray(guesses[1], dtype=np.int64)
        else:
            goff[1:] = np.cumsum([len(g) for g in guesses])
            flat = [complex(z) for g in guesses for z in g]
        gz = np.ascontiguousarray(np.array(flat, dtype=np.complex128)) if len(flat) else np.zeros(1, np.complex128)
        fn = self.lib.psi_mode_batch_device if device_aaa else self.lib.psi_mode_batch_host
        fn.restype = C.c_int
        while True:
            roots = np.zeros((n, max_roots), dtype=np.complex128)
            nr = np.zeros(n, dtype=np.int32)
            nc = np.zeros(n, dtype=np.int32)
            st = np.zeros(n, dtype=np.int32)
            stats = np.zeros(16, dtype=np.int64)
            self._check(fn(self.handle, C.c_int(n), _ptr(dist), _ptr(kx), _ptr(kz), _ptr(alpha),
                           _ptr(domain), _ptr(ip), _ptr(dp), _ptr(seeds), _ptr(goff), _ptr(gz),
                           C.c_int(max_roots), _ptr(roots), _ptr(nr), _ptr(nc), _ptr(st), _ptr(stats)),
                        "psi_mode_batch")
            if device_aaa:
                stats[5], stats[6], stats[0] = stats[0], 0, 1
                st = np.where(st == 4, 0, st)
                if np.any(st != 0):
                    bad = int(np.flatnonzero(st != 0)[0])
                    raise PsiError("device root search failed for item %d (Kx=%g, Kz=%g): status %d"
                                   % (bad, kx[bad], kz[bad], st[bad]))
            if n and nr.max() > max_roots:           # rare: more roots than slots -> redo JUST those searches with room
                over = np.flatnonzero(nr > max_roots)
                if len(over) == n:
                    max_roots = int(nr.max())
                    continue
                g_over = None
                if len(flat):
                    length = (goff[1:] - goff[:-1])[over].astype(np.int64)
                    off = np.zeros(len(over) + 1, dtype=np.int64)
                    np.cumsum(length, out=off[1:])
                    take = np.repeat(goff[:-1][over].astype(np.int64) - off[:-1], length) + np.arange(off[-1])
                    g_over = (np.asarray(gz)[take], off)
                (r2, n2), c2, _ = self._mode_batch(dist[over], kx[over], kz[over], alpha[over], domain[over],
                                                   ip[over, 0], ip[over, 1], ip[over, 2], dp[over, 0], dp[over, 1],
                                                   seeds[over], g_over, int(nr.max()), device_aaa, True)
                wide = np.zeros((n, r2.shape[1]), dtype=np.complex128)
                wide[:, :max_roots] = roots
                wide[over] = r2
                roots, max_roots = wide, r2.shape[1]
                nr[over], nc[over] = n2, c2
            break
        if np.any(st == 1):
            bad = int(np.flatnonzero(st == 1)[0])
            raise PsiError("root search failed for item %d (Kx=%g, Kz=%g): %s"
                           % (bad, kx[bad], kz[bad], self._last_error()))
        st_out = dict(passes=int(stats[0]), evals=int(stats[1]), logic_s=stats[3]*1e-6,
                      eval_s=stats[4]*1e-6, launches=int(stats[5]), polish_evals=int(stats[6]),
                      stage_s=dict(zip(("sample", "k1", "analyze", "k3", "decide"), (stats[8:13]*1e-6).tolist())))
        if dense:
            return (roots, nr), nc, st_out
        out = [roots[i, :nr[i]].copy() for i in range(n)]
        return out, nc, st_out

    def mode_batch_records(self, dist, kx, kz, alpha, domain, n_sample, max_zoom, max_secant, tol, clean_tol,
                           seeds, guesses, max_roots=8, rec_rows=0):
        """K4 batch whose results stay on the device (``DeviceRecords``) -- for the all-gather of a multi-GPU map.
        Same arguments as ``mode_batch``; ``rec_rows`` >= number of items pads the record table (all ranks of a
        collective need tables of one size)."""
        n = len(dist)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        dist, kx, kz, alpha = i32(dist), f64(kx), f64(kz), f64(alpha)
        domain = f64(domain).reshape(n, 4)
        ip = i32(np.stack([np.broadcast_to(n_sample, (n,)), np.broadcast_to(max_zoom, (n,)),
                           np.broadcast_to(max_secant, (n,))], axis=1))
        dp = f64(np.stack([np.broadcast_to(tol, (n,)), np.broadcast_to(clean_tol, (n,))], axis=1))
        seeds = _checked_seeds(seeds, n)
        goff = np.zeros(n + 1, dtype=np.int32)
        flat = np.zeros(0, np.complex128)
        if guesses is not None:
            if isinstance(guesses, tuple):
                flat = np.asarray(guesses[0], dtype=np.complex128)
                goff[:] = np.asarray(guesses[1], dtype=np.int64)
            else:
                goff[1:] = np.cumsum([len(g) for g in guesses])
                flat = np.array([complex(z) for g in guesses for z in g], dtype=np.complex128)
        gz = np.ascontiguousarray(flat) if len(flat) else np.zeros(1, np.complex128)
        rows = max(int(rec_rows), n, 1)
        ptr = _vp()
        stats = np.zeros(16, dtype=np.int64)
        self._check(self.lib.psi_mode_batch_device_rec(self.handle, C.c_int(n), _ptr(dist), _ptr(kx), _ptr(kz),
                                                       _ptr(alpha), _ptr(domain), _ptr(ip), _ptr(dp), _ptr(seeds),
                                                       _ptr(goff), _ptr(gz), C.c_int(max_roots), C.c_int(rows),
                                                       C.byref(ptr), _ptr(stats)), "psi_mode_batch_device_rec")
        st = dict(passes=1, evals=int(stats[1]), launches=int(stats[0]), logic_s=0.0, eval_s=0.0, polish_evals=0,
                  setup_s=stats[3]*1e-6, call_s=stats[4]*1e-6,
                  stage_s=dict(zip(("sample", "k1", "analyze", "k3", "decide"), (stats[8:13]*1e-6).tolist())))
        return DeviceRecords(ptr.value, rows, 4 + 2*max_roots, n, max_roots), st

    def mode_search_detail(self, dist_id, kx, kz, alpha, domain, n_sample, max_zoom, max_secant, tol, clean_tol,
                           uniforms, guesses, max_roots=16):
        """One PSIMode.calculate search through the K4 pipeline with the caller's own random stream; returns a dict
        with roots, n_function_call, status, z_sample, f_sample, mask (support points of the final rational
        approximation), n_uniforms_used and max_convergence_iterations."""
        g = np.ascontiguousarray(np.asarray(list(guesses), dtype=np.complex128))
        ng = len(g)
        cap = int(n_sample)*(1 + ng + 4*int(max_zoom))
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        dom = np.ascontiguousarray(domain, dtype=np.float64)
        ip = np.array([n_sample, max_zoom, max_secant], dtype=np.int32)
        dp = np.array([tol, clean_tol], dtype=np.float64)
        while True:
            roots = np.zeros(max_roots, dtype=np.complex128)
            ints = np.zeros(6, dtype=np.int32)      # n_roots, n_calls, status, n_samples, n_uni_used, max_conv
            Z, F = np.zeros(cap, np.complex128), np.zeros(cap, np.complex128)
            mask = np.zeros(cap, dtype=np.int32)
            ip_ = lambda k: C.cast(ints.ctypes.data + 4*k, _vp)
            self._check(self.lib.psi_mode_search_detail(
                self.handle, C.c_int(int(dist_id)), C.c_double(kx), C.c_double(kz), C.c_double(alpha), _ptr(dom),
                _ptr(ip), _ptr(dp), _ptr(u), C.c_int(len(u)), _ptr(g) if ng else None, C.c_int(ng),
                C.c_int(max_roots), _ptr(roots), ip_(0), ip_(1), ip_(2), _ptr(Z), _ptr(F), ip_(3), _ptr(mask),
                ip_(4), ip_(5)), "psi_mode_search_detail")
            if ints[0] > max_roots:
                max_roots = int(ints[0])
                continue
            break
        m = int(ints[3])
        return dict(roots=roots[:ints[0]].copy(), n_function_call=int(ints[1]), status=int(ints[2]),
                    z_sample=Z[:m].copy(), f_sample=F[:m].copy(), mask=mask[:m].astype(bool),
                    n_uniforms_used=int(ints[4]), max_convergence_iterations=int(ints[5]))

    def contour_batch(self, dist, kx, kz, alpha, z_lists, tol, max_step, max_iter):
        """Root counts of len(dist) closed contours with ONE launch of kernel K2.  Returns (counts,
        status, n_points, stats); counts are -666 where the count did not converge (status 1/2: the
        reference's RuntimeError; 3: non-finite D, its ValueError; 4: contour buffer exhausted)."""
        n = len(dist)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        dist, kx, kz, alpha = i32(dist), f64(kx), f64(kz), f64(alpha)
        zoff = np.zeros(n + 1, dtype=np.int32)
        zoff[1:] = np.cumsum([len(z) for z in z_lists])
        zz = np.ascontiguousarray(np.array([complex(v) for z in z_lists for v in z], dtype=np.complex128))
        tol = f64(np.broadcast_to(tol, (n,)))
        ms = f64(np.broadcast_to(max_step, (n,)))
        mi = i32(np.broadcast_to(max_iter, (n,)))
        counts = np.zeros(n, dtype=np.int32)
        st = np.zeros(n, dtype=np.int32)
        npts = np.zeros(n, dtype=np.int32)
        stats = np.zeros(2, dtype=np.int64)
        fn = self.lib.psi_contour_batch_host
        fn.restype = C.c_int
        self._check(fn(self.handle, C.c_int(n), _ptr(dist), _ptr(kx), _ptr(kz), _ptr(alpha), _ptr(zoff),
                       _ptr(zz), _ptr(tol), _ptr(ms), _ptr(mi), _ptr(counts), _ptr(st), _ptr(npts),
                       _ptr(stats)), "psi_contour_batch_host")
        return counts, st, npts, dict(passes=int(stats[0]), evals=int(stats[1]))

    def polish_batch(self, dist, kx, kz, alpha, domain, starts, max_iter):
        """Kernel K3: secant polish of every list of start values in ``starts`` (one list per item).
        Returns (list of arrays with the roots or the sentinel -1j, n_calls, max_conv, status)."""
        n = len(dist)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        dist, kx, kz, alpha = i32(dist), f64(kx), f64(kz), f64(alpha)
        domain = f64(domain).reshape(n, 4)
        off = np.zeros(n + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(s) for s in starts])
        flat = [complex(z) for s in starts for z in s]
        w0 = np.ascontiguousarray(np.array(flat, dtype=np.complex128)) if flat else np.zeros(1, np.complex128)
        mi = i32(np.broadcast_to(max_iter, (n,)))
        roots = np.zeros(max(1, int(off[-1])), dtype=np.complex128)
        nc = np.zeros(n, dtype=np.int32)
        mc = np.zeros(n, dtype=np.int32)
        st = np.zeros(n, dtype=np.int32)
        fn = self.lib.psi_polish_batch_host
        fn.restype = C.c_int
        self._check(fn(self.handle, C.c_int(n), _ptr(dist), _ptr(kx), _ptr(kz), _ptr(alpha), _ptr(domain),
                       _ptr(off), _ptr(w0), _ptr(mi), _ptr(roots), _ptr(nc), _ptr(mc), _ptr(st)),
                    "psi_polish_batch_host")
        return [roots[off[i]:off[i + 1]].copy() for i in range(n)], nc, mc, st

    def _last_error(self):
        return load_library().psi_last_error().decode()


def tv_roots(dust_fraction, tau_max, b, tau_min, diffusion, max_it, kx, kz, device=None):
    """psi_tv_roots: terminal-velocity mode frequency for arrays of wave-number pairs (one thread per pair)."""
    lib = load_library()
    kx = np.ascontiguousarray(kx, dtype=np.float64)
    kz = np.ascontiguousarray(kz, dtype=np.float64)
    w = np.zeros(len(kx), dtype=np.complex128)
    st = np.zeros(len(kx), dtype=np.int32)
    check(lib.psi_tv_roots(default_device() if device is None else int(device), float(dust_fraction), float(tau_max),
                           float(b), float(tau_min), float(diffusion), int(max_it), len(kx), _ptr(kx), _ptr(kz),
                           _ptr(w), _ptr(st)), "psi_tv_roots")
    return w, st


class RefineMap:
    """Device mirror of a PSIGridRefiner grid (run ids, nguess, RootStore) and its sweep kernels
    (csrc/psi_refine.cu).  ``sweep()`` returns the pixels that run next with their pruned guesses;
    ``update()`` stores the finished runs."""

    def __init__(self, device, runs, nguess, store_mat, store_cnt, key_capacity):
        self.lib = load_library()
        self.shape = runs.shape
        self.width = int(store_mat.shape[1])
        self.key_capacity = int(key_capacity)
        n_keys = len(store_cnt)
        runs32 = np.ascontiguousarray(runs, dtype=np.int32)
        ng32 = np.ascontiguousarray(nguess, dtype=np.int32)
        cnt32 = np.ascontiguousarray(store_cnt, dtype=np.int32)
        mat = np.ascontiguousarray(store_mat, dtype=np.complex128)
        h = _vp()
        check(self.lib.psi_refine_open(C.byref(h), int(device), int(runs.shape[0]), int(runs.shape[1]),
                                       _ptr(runs32), _ptr(ng32), n_keys, self.key_capacity, self.width,
                                       _ptr(mat), _ptr(cnt32)), "psi_refine_open")
        self.handle = h

    def sweep(self):
        """(iz, ix, guess counts, flat guesses) of the pixels that run, in row-major pixel order; None when a
        pixel collects more neighbour roots than the kernel's buffer holds (nothing changed on the device)."""
        n_run, n_guess = C.c_longlong(0), C.c_longlong(0)
        check(self.lib.psi_refine_sweep(self.handle, C.byref(n_run), C.byref(n_guess)), "psi_refine_sweep")
        if n_run.value < 0:
            return None
        pix = np.zeros(n_run.value, dtype=np.int32)
        nk = np.zeros(n_run.value, dtype=np.int32)
        g = np.zeros(max(1, n_guess.value), dtype=np.complex128)
        check(self.lib.psi_refine_fetch(self.handle, _ptr(pix), _ptr(nk), _ptr(g)), "psi_refine_fetch")
        return pix//self.shape[1], pix % self.shape[1], nk.astype(np.int64), g[:n_guess.value]

    def update(self, keys, cnt, rows, iz, ix, run_ids):
        keys = np.ascontiguousarray(keys, dtype=np.int64)
        cnt = np.ascontiguousarray(cnt, dtype=np.int32)
        rows = np.ascontiguousarray(rows, dtype=np.complex128)
        assert rows.shape == (len(keys), self.width)
        pix = np.ascontiguousarray(np.asarray(iz)*self.shape[1] + np.asarray(ix), dtype=np.int32)
        ids = np.ascontiguousarray(run_ids, dtype=np.int32)
        check(self.lib.psi_refine_update(self.handle, len(keys), _ptr(keys), _ptr(cnt), _ptr(rows), len(pix),
                                         _ptr(pix), _ptr(ids)), "psi_refine_update")

    def close(self):
        if getattr(self, "handle", None):
            self.lib.psi_refine_close(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context(EngineAPI):
    """One device + one node table (``psi_ctx``).  Distributions are registered once and cached by
    the bytes of their descriptor."""

    def __init__(self, integrator, device=None):
        lib = load_library()
        self.lib = lib
        self.device = default_device() if device is None else device
        xj = np.ascontiguousarray(integrator.xj, dtype=np.float64)
        wj = np.ascontiguousarray(integrator.wj, dtype=np.float64)
        h = _vp()
        check(lib.psi_ctx_create(C.byref(h), self.device, _ptr(xj), _ptr(wj), len(xj),
                                 int(integrator.max_m), float(integrator.eps)), "psi_ctx_create")
        self.handle = h
        self._dists = {}
        self.node_evals = 0

    def close(self):
        if getattr(self, "handle", None):
            self.lib.psi_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def sm_count(self):
        return self.lib.psi_ctx_sm_count(self.handle)

    @property
    def exit_rule(self):
        """'predicted' (default): the tanh-sinh level loop stops once the increments have collapsed
        quadratically and the next one is predicted far below the tolerance; 'reference': the literal rule of
        tanhsinh.py:120-126 (increment below 1e-15 absolute), which spends its last levels on rounding noise."""
        return "predicted" if self.lib.psi_ctx_get_exit_rule(self.handle) else "reference"

    @exit_rule.setter
    def exit_rule(self, rule):
        if rule not in ("predicted", "reference"):
            raise ValueError("exit_rule must be 'predicted' or 'reference'")
        check(self.lib.psi_ctx_set_exit_rule(self.handle, 1 if rule == "predicted" else 0), "psi_ctx_set_exit_rule")

    @property
    def launch_count(self):
        return int(self.lib.psi_ctx_launch_count(self.handle))

    def synchronize(self):
        check(self.lib.psi_ctx_synchronize(self.handle), "psi_ctx_synchronize")

    def add_dist(self, desc, table=None):
        k = desc.key() + (table.key() if table is not None else b"")
        if k not in self._dists:
            if table is None:
                self._dists[k] = check(self.lib.psi_dist_add(self.handle, C.byref(desc)), "psi_dist_add")
            else:
                self._dists[k] = check(self.lib.psi_dist_add_table(self.handle, C.byref(desc),
                                                                   C.byref(table.struct())), "psi_dist_add_table")
        return self._dists[k]

    def background(self, dist_id):
        out = (C.c_double*5)()
        check(self.lib.psi_dist_get_background(self.handle, dist_id, out), "psi_dist_get_background")
        return dict(j0=out[0], j1=out[1], denom=out[2], vgx=out[3], vgy=out[4])

    def disp_eval(self, dist, kx, kz, alpha, w, want_M=False, out=None):
        """Batch D(omega).  dist/kx/kz/alpha are scalars (broadcast) or arrays shaped like w.
        ``out`` may be a preallocated (e.g. pinned) complex128 array of the same length."""
        w = np.ascontiguousarray(np.ravel(w), dtype=np.complex128)
        n = len(w)
        bc = int(np.ndim(kx) == 0 and np.ndim(kz) == 0 and np.ndim(alpha) == 0 and np.ndim(dist) == 0)
        m = 1 if bc else n
        da = np.ascontiguousarray(np.broadcast_to(np.asarray(dist, dtype=np.int32), (m,)))
        kxa = np.ascontiguousarray(np.broadcast_to(np.asarray(kx, dtype=np.float64), (m,)))
        kza = np.ascontiguousarray(np.broadcast_to(np.asarray(kz, dtype=np.float64), (m,)))
        ala = np.ascontiguousarray(np.broadcast_to(np.asarray(alpha, dtype=np.float64), (m,)))
        D = np.empty(n, dtype=np.complex128) if out is None else out
        if D.dtype != np.complex128 or D.size != n or not D.flags.c_contiguous:
            raise ValueError("out must be a contiguous complex128 array with one entry per point")
        M = np.empty((n, 7), dtype=np.complex128) if want_M else None
        ne = C.c_ulonglong(0)
        check(self.lib.psi_disp_eval_host(self.handle, n, bc, _ptr(da), _ptr(kxa), _ptr(kza),
                                          _ptr(ala), _ptr(w), _ptr(D),
                                          _ptr(M) if want_M else None, C.byref(ne)),
              "psi_disp_eval_host")
        self.node_evals += ne.value
        return (D, M) if want_M else D

    def point_class(self, dist_id, alpha):
        return check(self.lib.psi_point_class(self.handle, int(dist_id), float(alpha)), "psi_point_class")

    def disp_eval_dev(self, n, broadcast, dist_ptr, kx_ptr, kz_ptr, alpha_ptr, w_ptr, D_ptr, M_ptr=0,
                      node_evals_ptr=0, class_mask=0, stream=0):
        """K1 on DEVICE pointers (ints), asynchronous on ``stream`` (a cudaStream_t as int; 0 = the
        context's own stream).  Used when inputs are already resident in HBM."""
        check(self.lib.psi_disp_eval_dev(self.handle, int(n), int(broadcast), dist_ptr, kx_ptr, kz_ptr,
                                         alpha_ptr, w_ptr, D_ptr, M_ptr or None,
                                         node_evals_ptr or None, int(class_mask), stream or None),
              "psi_disp_eval_dev")

    def measure_fp64_peak(self, reps=5):
        out = C.c_double(0)
        check(self.lib.psi_measure_fp64_peak(self.handle, reps, C.byref(out)), "psi_measure_fp64_peak")
        return out.value


_contexts = {}


def get_context(integrator, device=None):
    """Context cache keyed by (device, node-table identity)."""
    dev = default_device() if device is None else device
    key = (dev, int(integrator.p), int(integrator.max_m), len(integrator.xj),
           hash(np.asarray(integrator.xj).tobytes()))
    if key not in _contexts:
        _contexts[key] = Context(integrator, dev)
    return _contexts[key]


# ------------------------------------------------------------------------------------------------
# Python objects -> device descriptors
# ------------------------------------------------------------------------------------------------
def _closure_vars(fn):
    code = getattr(fn, "__code__", None)
    cells = getattr(fn, "__closure__", None)
    if code is None or not cells:
        return {}
    out = {}
    for name, cell in zip(code.co_freevars, cells):
        try:
            out[name] = cell.cell_contents
        except ValueError:
            pass
    return out


def _find_sigma(size_density):
    """The sigma callable behind a SizeDensity: ours keeps ``.sigma``; the reference's keeps it
    only inside the closure of ``.f`` (reference sizedensity.py:41)."""
    s = getattr(size_density, "sigma", None)
    if s is not None:
        return s
    return _closure_vars(getattr(size_density, "f", None)).get("sigma")


def _bump_constants(pb):
    """(beta, norm_pl, amp, curv) of a PowerBump-like object: attributes on ours, closure cells of
    the ``fnn``/``b`` lambdas on the reference's (power_bump.py:75-80)."""
    if all(hasattr(pb, k) for k in ("beta", "norm_pl", "amp", "curv")):
        return float(pb.beta), float(pb.norm_pl), float(pb.amp), float(pb.curv)
    bv = _closure_vars(pb.b)
    amp, curv = float(bv["fac_b"]), float(bv["fac_b2"])
    if hasattr(pb, "beta"):                       # reference PowerBumpTail
        beta = float(pb.beta)
        norm_pl = float(pb.aP**-(beta))
    else:
        fv = _closure_vars(pb.fnn)
        beta, norm_pl = float(fv["beta"]), float(fv["fac_fnn"])
    return beta, norm_pl, amp, curv


def _analytic_F(d, s):
    """F(s) = taumax*sigma(taumax*s)/rhod of an analytic descriptor, as the device evaluates it (csrc/psi_math.cuh:
    sigma_of, F_of) -- used to CHECK that a recognised callable really is the family its name suggests."""
    a = d.tau_max*np.asarray(s, dtype=float)
    if d.kind == KIND_POWERLAW:
        sig = a**d.q
    else:
        pl = d.norm_pl*a**d.beta
        if d.kind == KIND_POWERBUMPTAIL:
            pl = np.where(a > d.aBT, pl, d.norm_pl*d.aBT**d.beta*np.sqrt(a/d.aBT))
        bb = d.amp*np.exp(d.curv*(a - d.aP)**2)
        v = np.where(a > d.aL, np.maximum(pl, bb), pl)
        v = np.where(a > d.aP, bb, v)
        sig = v*a**3*d.scale
    return d.tau_max*sig/d.rhod


def _matches_callable(d, size_density, n_probe=48):
    """True when the analytic descriptor reproduces the user's callable at n_probe log-spaced sizes (1e-12)."""
    s = np.exp(np.linspace(np.log(d.tau_min/d.tau_max), 0.0, n_probe + 2)[1:-1])
    try:
        ref = np.asarray(size_density(s), dtype=float)
    except Exception:
        return False
    got = _analytic_F(d, s)
    return bool(np.all(np.isfinite(ref)) and np.all(np.abs(got - ref) <= 1e-12*np.abs(ref) + 1e-300))


def describe_size_density(mu, stokes_range, sound_speed, power, single_size, size_density):
    """Build the device description of PSIDispersion's constructor arguments (reference psi_dispersion.py:43-66):
    (``psi_dist_desc``, table or None).

    * no size density: the power law a^(3 - power) with its closed-form rho_d;
    * a SizeDensity wrapping PowerBump.sigma0 / PowerBumpTail.sigma0: the analytic family -- AFTER checking at 48
      probe sizes that the callable really evaluates to the family's formula (a subclass with its own sigma0 does
      not, and is tabulated instead of being silently mistaken for the stock bump);
    * any other callable (the reference takes any ``SizeDensity(sigma, size_range)``, sizedensity.py:25-48): kind
      TABLE -- the quadrature weight tabulated once on the host (SizeDensityTable) and read on the device as
      piecewise cubics, two 16-byte loads per node."""
    d = DistDesc()
    d.single_size = int(bool(single_size))
    d.mu = float(mu)
    d.tau_min = float(stokes_range[0])
    d.tau_max = float(stokes_range[1])
    d.sound_speed = float(sound_speed)
    d.scale = 1.0
    poles = []
    table = None
    if size_density is None:
        d.kind = KIND_POWERLAW
        d.q = 3 - power
        d.rhod = (stokes_range[0]**(4.0 - power) - stokes_range[1]**(4.0 - power))/(power - 4.0)
    else:
        poles = [float(p) for p in size_density.poles]
        if len(poles) > PSI_MAX_POLES:
            raise ValueError("at most %d size-density poles are supported" % PSI_MAX_POLES)
        d.rhod = float(size_density.rhod)
        sigma = _find_sigma(size_density)
        pb = getattr(sigma, "__self__", None)
        is_sigma0 = getattr(sigma, "__name__", "") == "sigma0"
        analytic = (pb is not None and is_sigma0 and all(hasattr(pb, k) for k in ("aP", "aL", "aR", "amin"))
                    and float(size_density.amax) == d.tau_max)
        if analytic:
            try:
                beta, norm_pl, amp, curv = _bump_constants(pb)
                d.kind = KIND_POWERBUMPTAIL if hasattr(pb, "aBT") else KIND_POWERBUMP
                d.beta, d.norm_pl, d.amp, d.curv = beta, norm_pl, amp, curv
                d.aL, d.aP = float(pb.aL), float(pb.aP)
                d.aBT = float(getattr(pb, "aBT", -1.0))
                if getattr(pb, "epstot", None) is not None:
                    d.scale = float(pb.epstot)/float(pb.mnorm)
                analytic = _matches_callable(d, size_density)
            except Exception:
                analytic = False
        if not analytic:
            d.kind = KIND_TABLE
            d.q = d.beta = d.norm_pl = d.amp = d.curv = d.aL = d.aP = d.aBT = 0.0
            d.scale = 1.0
            table = SizeDensityTable(size_density, d.tau_min, d.tau_max, poles)
    d.n_poles = len(poles)
    for i, p in enumerate(poles):
        d.poles[i] = p
    return d, table
