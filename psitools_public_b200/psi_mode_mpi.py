"""Batch drivers with the interface of psitools.psi_mode_mpi (reference psi_mode_mpi.py:33-192).

``MpiScheduler(wall_start, wall_limit_total, verbose).run(arglist)`` keeps its meaning -- a list of
arg-dicts {'__init__': PSIMode kwargs, 'calculate': PSIMode.calculate kwargs, 'random_seed': int}
in, {index: roots} out on rank 0 (None on the other ranks) -- but instead of a master handing
single items to MPI workers, every rank (one per GPU under torchrun) takes the items
i = rank (mod world), advances all their root searches in lock-step on its GPU (one K1 launch per
pass, _engine.py) and the results are exchanged with a single all-gather (_dist.py).  Work items
re-seed their own random stream exactly like the reference's runcompute (:162-167), so results do
not depend on the schedule.
"""
import os
import time
from operator import itemgetter

import numpy as np

from . import _dist
from ._engine import run_lockstep
from .psi_mode import ModeTrace, PSIMode


DEFAULT_ENGINE = "device"
_GET_INIT, _GET_CALC, _GET_SEED = itemgetter('__init__'), itemgetter('calculate'), itemgetter('random_seed')
_GET_KX, _GET_KZ = itemgetter('wave_number_x'), itemgetter('wave_number_z')


def _fingerprint(obj, depth=0):
    """Structural key of a PSIMode '__init__' value.  The reference re-creates its PSIMode whenever the
    '__init__' dict of an item differs from the previous one (psi_mode_mpi.py:168-171); deep copies of an
    arg-dict (psi_grid_refine.py makes one per run) compare unequal there as soon as they hold an object
    without __eq__ (a SizeDensity), although they describe the same PSIMode.  Comparing by content lets
    such campaigns share one PSIMode (and one device distribution)."""
    if obj is None or isinstance(obj, (bool, int, float, complex, str, bytes)):
        return (type(obj).__name__, obj)
    if isinstance(obj, np.generic):
        return (type(obj).__name__, obj.item())
    if isinstance(obj, np.ndarray):
        return ("ndarray", obj.shape, obj.dtype.str, obj.tobytes())
    if isinstance(obj, (list, tuple)):
        return ("seq", tuple(_fingerprint(v, depth + 1) for v in obj))
    if isinstance(obj, dict):
        return ("dict", tuple(sorted((str(k), _fingerprint(v, depth + 1)) for k, v in obj.items())))
    if depth > 6 or isinstance(obj, type):
        return ("id", id(obj))
    bound_to = getattr(obj, "__self__", None)
    if bound_to is not None and callable(obj) and not isinstance(bound_to, type(np)):
        return ("method", getattr(obj, "__name__", ""), _fingerprint(bound_to, depth + 1))   # e.g. PowerBump.sigma0
    if hasattr(obj, "__code__") or callable(obj) and not hasattr(obj, "__dict__"):
        return ("id", id(obj))                           # plain functions and builtins: identity
    state = getattr(obj, "__dict__", None)
    if isinstance(state, dict):
        if "xj" in state and "wj" in state:               # quadrature tables: shape and level are enough
            return ("table", type(obj).__name__, len(state["xj"]), getattr(obj, "max_m", None))
        return ("obj", type(obj).__qualname__, _fingerprint(state, depth + 1))
    return ("id", id(obj))


def _init_key(init):
    try:
        key = _fingerprint(init)
        hash(key)
    except Exception:
        key = ("id", id(init))
    return key


def _same_init(a, b):
    return a is b or _init_key(a) == _init_key(b)


class _ModeCache:
    """PSIMode objects keyed by the content of the '__init__' dict (reference psi_mode_mpi.py:168-171)."""

    def __init__(self):
        self.entries = {}

    def get(self, init):
        key = _init_key(init)
        pm = self.entries.get(key)
        if pm is None:
            pm = PSIMode(**init)
            self.entries[key] = (pm, init)               # keep `init` alive: ids inside the key stay unique
            return pm
        return pm[0]


class MpiScheduler:
    """Execute a set of PSIMode runs given as a list of arg-dicts."""

    def __init__(self, wall_start, wall_limit_total, verbose=True, engine=None):
        self.verbose = verbose
        self.exitFlag = False
        self.rank = _dist.rank()
        self.nranks = _dist.world_size()
        self.wall_start = wall_start
        self.wall_limit_total = wall_limit_total
        self._cache = _ModeCache()
        self.last_stats = {}
        # "device": kernel K4 -- every search runs entirely on the GPU (AAA by Jacobi SVD + Aberth);
        # "native": the C++ lock-step engine (csrc/psi_engine.cu): AAA with LAPACK on the host cores,
        #           D (K1) and the secant polish (K3) on the GPU;
        # "python": the Python generators of psi_mode.py (same logic and LAPACK as "native", slow).
        self.engine = engine or os.environ.get("PSI_B200_ENGINE", DEFAULT_ENGINE)
        if self.engine not in ("device", "native", "python"):
            raise ValueError("engine must be 'device', 'native' or 'python'")

    # -- per-item body ------------------------------------------------------------------------
    def _job(self, args):
        pm = self._cache.get(args['__init__'])
        rng = np.random.RandomState(args.get('random_seed', 0))
        trace = ModeTrace()
        job = pm.make_job(rng=rng, trace=trace, **args['calculate'])
        return job, trace, pm

    def runcompute(self, args):
        """One work item on its own (kept for callers that used the reference's method)."""
        job, _, pm = self._job(args)
        run_lockstep(pm.psi_dispersion.ctx, [job])
        return job.result.copy()

    def _encode(self, result):
        return np.asarray(result, dtype=np.complex128).view(np.float64)

    def _decode(self, rec):
        return np.asarray(rec, dtype=np.float64).view(np.complex128).copy()

    def _run_native(self, arglist, mine, rec_rows=None, max_roots=8):
        """All of this rank's items through the batched engine (K4 pipeline or C++ lock-step engine);
        None if an item needs the general (Python) path.  Returns (root matrix, root counts, stats) -- or, with
        ``rec_rows`` (device engine, multi-GPU), (DeviceRecords, None, stats): the results stay on the device for
        the all-gather."""
        if not mine and rec_rows is None:
            return (np.zeros((0, 1), dtype=np.complex128), np.zeros(0, dtype=np.int64)), {}
        if not mine:
            sub0 = arglist[0]
            ctx = self._cache.get(sub0['__init__']).psi_dispersion.ctx
            z = np.zeros(0)
            rec, st = ctx.mode_batch_records(z.astype(np.int32), z, z, z, np.zeros((0, 4)), z, z, z, z, z, z, None,
                                             max_roots=max_roots, rec_rows=rec_rows)
            return (rec, None), st
        # One C-speed pass per field (map + itemgetter) instead of Python loops over the items: on a 256^2 map
        # the arg-dict handling was 30-50 ms of a 0.3 s batch.
        sub = [arglist[i] for i in mine] if len(mine) != len(arglist) else arglist
        n = len(sub)
        # PSIMode per distinct '__init__' (by identity first: map scripts share one dict between items)
        inits = list(map(_GET_INIT, sub))
        ids = np.fromiter(map(id, inits), dtype=np.int64, count=n)
        uniq, first, inverse = np.unique(ids, return_index=True, return_inverse=True)
        pms, slot = [], np.empty(len(uniq), dtype=np.int64)
        for u, k in enumerate(first.tolist()):
            pm = self._cache.get(inits[k])
            for j, known in enumerate(pms):
                if known is pm:
                    break
            else:
                pms.append(pm)
                j = len(pms) - 1
            slot[u] = j
        which = slot[inverse]
        calc = list(map(_GET_CALC, sub))
        # items that carry nothing but the two wave numbers (the usual map) need no further passes
        plain = max(map(len, calc)) == 2
        if any(pm.verbose_flag for pm in pms) or \
                (not plain and any(c.get('count_roots', False) for c in calc)):
            return None
        ctx = pms[0].psi_dispersion.ctx
        if any(pm.psi_dispersion.ctx is not ctx for pm in pms) or not hasattr(ctx, "mode_batch"):
            return None
        for pm in pms:
            pm._sync_cfg()
        per_pm = lambda f, dt=float: np.array([f(pm) for pm in pms], dtype=dt)[which]
        dom = np.array([[pm.domain.xmin, pm.domain.xmax, pm.domain.ymin, pm.domain.ymax] for pm in pms],
                       dtype=float)[which]
        n_sample = per_pm(lambda pm: pm.cfg.n_sample, int)
        max_zoom = per_pm(lambda pm: pm.cfg.max_zoom_level, int)
        if plain:
            guesses, n_guess = None, np.zeros(n, dtype=np.int64)
            alpha = np.zeros(n)
        else:
            guesses = [c.get('guess_roots', ()) for c in calc]
            n_guess = np.fromiter(map(len, guesses), dtype=np.int64, count=n)
            alpha = np.fromiter((c.get('viscous_alpha', 0) for c in calc), dtype=float, count=n)
        try:
            seeds = np.fromiter(map(_GET_SEED, sub), dtype=np.int64, count=n)
        except KeyError:
            seeds = np.fromiter((a.get('random_seed', 0) for a in sub), dtype=np.int64, count=n)
        device = self.engine == "device"      # (searches beyond the pipeline's 512 samples: Context.mode_batch routes
        batch_args = (                          # just those to the host engine)
            per_pm(lambda pm: pm.psi_dispersion.dist_id, int),
            np.fromiter(map(_GET_KX, calc), dtype=float, count=n),
            np.fromiter(map(_GET_KZ, calc), dtype=float, count=n), alpha,
            dom, n_sample, max_zoom, per_pm(lambda pm: pm.cfg.max_secant_iterations, int),
            per_pm(lambda pm: pm.cfg.tol), per_pm(lambda pm: pm.cfg.clean_tol), seeds,
            guesses if n_guess.any() else None)
        if rec_rows is not None:
            rec, st = ctx.mode_batch_records(*batch_args, max_roots=max_roots, rec_rows=rec_rows)
            return (rec, None), st
        (roots, nr), ncalls, st = ctx.mode_batch(*batch_args, device_aaa=device, dense=True)
        st["function_calls"] = int(np.sum(ncalls))
        return (roots, nr), st

    def run_arrays(self, init, kx, kz, seeds, guesses=None, alpha=0.0):
        """Array form of ``run_all`` for campaigns whose items share one '__init__' dict (grid refinement):
        per-item wave numbers, seeds and -- as the CSR pair (flat values, n+1 offsets) -- guess roots in,
        (root matrix (n, R), root counts (n,)) of ALL items out on every rank.  Same sharding, engine and
        exchange as ``run_all``; no per-item Python objects.  Returns None when the batch needs the
        general path (verbose PSIMode, python engine)."""
        t0 = time.time()
        if self.engine == "python":
            return None
        pm = self._cache.get(init)
        ctx = pm.psi_dispersion.ctx
        if pm.verbose_flag or not hasattr(ctx, "mode_batch"):
            return None
        pm._sync_cfg()
        n = len(kx)
        lo, step = _dist.rank(), _dist.world_size()
        mine = np.arange(lo, n, step)
        kx_in, kz_in, seeds_in = np.asarray(kx, dtype=float), np.asarray(kz, dtype=float), seeds
        al_in = np.broadcast_to(np.asarray(alpha, dtype=float), (n,))
        kx, kz = kx_in[mine], kz_in[mine]
        seeds = np.broadcast_to(np.asarray(seeds, dtype=np.int64), (n,))[mine]
        al = al_in[mine]
        gl = None
        if guesses is not None and len(guesses[0]):
            flat, off = np.asarray(guesses[0], dtype=np.complex128), np.asarray(guesses[1], dtype=np.int64)
            length = (off[1:] - off[:-1])[mine]
            myoff = np.zeros(len(mine) + 1, dtype=np.int64)
            np.cumsum(length, out=myoff[1:])
            take = np.repeat(off[:-1][mine] - myoff[:-1], length) + np.arange(myoff[-1])
            gl = (flat[take], myoff)
        device = self.engine == "device"
        cfg = pm.cfg
        dom = np.tile(np.array([pm.domain.xmin, pm.domain.xmax, pm.domain.ymin, pm.domain.ymax], dtype=float),
                      (len(mine), 1))
        t1 = time.time()
        dist_ids = np.full(len(mine), pm.psi_dispersion.dist_id, dtype=np.int32)
        n_guess_max = int(np.max(gl[1][1:] - gl[1][:-1])) if gl is not None and len(mine) else 0
        cap = cfg.n_sample*(1 + n_guess_max + 4*cfg.max_zoom_level)
        fits = cap <= 512 and 60*cap*max(len(mine), 1) <= int(os.environ.get("PSI_B200_BATCH_BYTES", str(48 << 30)))
        use_records = device and _dist.nccl_active() and _dist.all_reduce_max(0 if fits else 1) == 0
        t1b = time.time()                        # (the all-reduce doubles as the point where ranks wait for each other)
        if use_records:
            # results stay on the device: one ncclAllGather of packed records out of the library's buffer
            rows, max_roots, redone = (n + step - 1)//step, 4, 0
            while True:
                rec, st = ctx.mode_batch_records(dist_ids, kx, kz, al, dom, cfg.n_sample, cfg.max_zoom_level,
                                                 cfg.max_secant_iterations, cfg.tol, cfg.clean_tol, seeds, gl,
                                                 max_roots=max_roots, rec_rows=rows)
                t2 = time.time()
                got = _dist.all_gather_device_records(rec, n, to_host=True)
                if got["worst"] not in (0, 4):
                    raise RuntimeError("device root search failed on some rank (status %d)" % got["worst"])
                break
            top, cnt_all = got["top"], got["cnt"]
            roots_all = np.zeros((n, max_roots), dtype=np.complex128)
            roots_all[got["idx"]] = got["roots"]
            over = np.flatnonzero(cnt_all > max_roots)
            if len(over):
                # rare: a search found more roots than the records have slots.  Every rank holds the whole table, so
                # every rank redoes just those searches itself (deterministic: identical everywhere), no second exchange
                redone = len(over)
                seeds_all = np.broadcast_to(np.asarray(seeds_in, dtype=np.int64), (n,))
                g_over = None
                if guesses is not None and len(guesses[0]):
                    flat, off = np.asarray(guesses[0], dtype=np.complex128), np.asarray(guesses[1], dtype=np.int64)
                    length = (off[1:] - off[:-1])[over]
                    o2 = np.zeros(len(over) + 1, dtype=np.int64)
                    np.cumsum(length, out=o2[1:])
                    g_over = (flat[np.repeat(off[:-1][over] - o2[:-1], length) + np.arange(o2[-1])], o2)
                (r2, n2), _, _ = ctx.mode_batch(np.full(len(over), pm.psi_dispersion.dist_id, dtype=np.int32),
                                                kx_in[over], kz_in[over], al_in[over],
                                                np.tile(np.array([[pm.domain.xmin, pm.domain.xmax, pm.domain.ymin,
                                                                   pm.domain.ymax]], dtype=float), (len(over), 1)),
                                                cfg.n_sample, cfg.max_zoom_level, cfg.max_secant_iterations, cfg.tol,
                                                cfg.clean_tol, seeds_all[over], g_over, max_roots=top,
                                                device_aaa=True, dense=True)
                wide = np.zeros((n, r2.shape[1]), dtype=np.complex128)
                wide[:, :roots_all.shape[1]] = roots_all
                wide[over] = r2
                roots_all = wide
                cnt_all[over] = n2
            width = max(1, int(cnt_all.max()) if n else 1)
            self.last_stats = dict(items=len(mine), passes=1, function_calls=got["calls"],
                                   evals=st.get("evals", 0), seconds=time.time() - t0, prep_seconds=t1 - t0,
                                   batch_seconds=t2 - t1b, wait_seconds=t1b - t1, launches=st.get("launches"),
                                   lib_setup_seconds=st.get("setup_s"), lib_call_seconds=st.get("call_s"),
                                   redone_for_root_slots=redone, max_roots=max_roots,
                                   stage_s=st.get("stage_s"), engine=self.engine, exchange_seconds=time.time() - t2,
                                   exchange="nccl all_gather_into_tensor (device records)")
            return np.ascontiguousarray(roots_all[:, :width]), cnt_all
        if len(mine):
            (roots, nr), ncalls, st = ctx.mode_batch(
                dist_ids, kx, kz, al, dom, cfg.n_sample,
                cfg.max_zoom_level, cfg.max_secant_iterations, cfg.tol, cfg.clean_tol, seeds, gl,
                device_aaa=device, dense=True)
        else:
            roots, nr, ncalls, st = np.zeros((0, 1), np.complex128), np.zeros(0, np.int64), np.zeros(0), {}
        t2 = time.time()
        self.last_stats = dict(items=len(mine), passes=st.get("passes", 0), function_calls=int(np.sum(ncalls)),
                               evals=st.get("evals", 0), seconds=time.time() - t0, prep_seconds=t1 - t0,
                               batch_seconds=t2 - t1,
                               host_logic_seconds=st.get("logic_s"), device_seconds=st.get("eval_s"),
                               launches=st.get("launches"), polish_evals=st.get("polish_evals"),
                               stage_s=st.get("stage_s"), engine=self.engine)
        if os.environ.get("PSI_B200_CALL_STATS") and len(mine):      # diagnostic: spread of samples per search
            c = np.asarray(ncalls)
            self.last_stats["calls_pct"] = [int(v) for v in np.percentile(c, [50, 90, 99, 99.9, 100])] + \
                [int(np.sum(c > 160)), int(np.sum(c > 320))]
        # exchange: per item a count, and the roots of the items that have any (rank q owns items q::world)
        nr = np.asarray(nr, dtype=np.int64)
        valid = np.arange(roots.shape[1])[None, :] < nr[:, None]
        cnts, flats = _dist.all_gather_ragged(nr, roots[valid])
        width = max([1] + [int(c.max()) for c in cnts if len(c)])
        out = np.zeros((n, width), dtype=np.complex128)
        out_n = np.zeros(n, dtype=np.int64)
        for q, (c, f) in enumerate(zip(cnts, flats)):
            out_n[q::step] = c
            rows = np.repeat(np.arange(q, n, step), c)
            cols = np.arange(len(f)) - np.repeat(np.cumsum(c) - c, c)
            out[rows, cols] = f
        self.last_stats["exchange_seconds"] = time.time() - t2
        return out, out_n

    def _records_ok(self, arglist):
        """Device-to-device exchange: device engine, NCCL, no search beyond the pipeline's 512 samples, batch within
        the device budget (one piece)."""
        if self.engine != "device" or not _dist.nccl_active() or not arglist:
            return False
        if len(arglist) > int(os.environ.get("PSI_B200_RECORD_ITEMS", "4000000")):
            return False
        return True

    def _run_records(self, arglist, to_host):
        """Multi-GPU map with the results gathered device to device: every rank solves its items i = rank (mod
        world), the packed records go through ONE ncclAllGather out of the library's own buffer, and only ranks
        that need the values (``to_host``) copy the gathered table to the host.  Returns the table rows in global
        item order as (root matrix, root counts), or None (general path needed)."""
        n, world = len(arglist), _dist.world_size()
        mine = _dist.my_items(n)
        rows = (n + world - 1)//world
        max_roots = 8
        tick = [time.time()]
        lap = lambda: tick.append(time.time())
        while True:
            try:
                done = self._run_native(arglist, mine, rec_rows=rows, max_roots=max_roots)
            except Exception as exc:             # (capacity etc.): every rank must still reach the collective
                done, err = "error", exc
            lap()
            flag = _dist.all_reduce_max(0 if isinstance(done, tuple) else (2 if done == "error" else 1))
            lap()
            if flag == 2:
                raise err if done == "error" else RuntimeError("root search failed on another rank")
            if flag == 1:
                return None
            (rec, _), st = done
            got = _dist.all_gather_device_records(rec, n, to_host=to_host)
            lap()
            if got["worst"] not in (0, 4):
                raise RuntimeError("device root search failed on some rank (status %d)" % got["worst"])
            if got["top"] > max_roots:            # rare: more roots than slots somewhere -> every rank redoes it
                max_roots = got["top"]
                continue
            break
        self.last_stats = dict(items=len(mine), passes=st.get("passes", 0), function_calls=got["calls"],
                               evals=st.get("evals", 0), launches=st.get("launches"), stage_s=st.get("stage_s"),
                               engine=self.engine, exchange="nccl all_gather_into_tensor (device records)",
                               host_split=dict(solve=tick[-3] - tick[0], lib_call=st.get("call_s"),
                                               flag_allreduce=tick[-2] - tick[-3], gather=tick[-1] - tick[-2]))
        if got["cnt"] is None:
            return ()
        roots = np.zeros((n, max_roots), dtype=np.complex128)
        roots[got["idx"]] = got["roots"]
        return roots, got["cnt"]

    @staticmethod
    def _as_dict(n, mat, cnt):
        """{index: roots}: root-less items (the majority on a map) share one read-only empty array."""
        none = np.zeros(0, dtype=np.complex128)
        none.flags.writeable = False
        res = dict.fromkeys(range(n), none)
        for k in np.flatnonzero(cnt > 0).tolist():
            res[k] = mat[k, :cnt[k]].copy()
        return res

    def run_all(self, arglist, root_only=False):
        """Results of every item, on every rank (``root_only``: on rank 0, None elsewhere)."""
        t0 = time.time()
        if self._records_ok(arglist):
            got = self._run_records(arglist, to_host=(not root_only) or self.rank == 0)
            if got is not None:
                self.last_stats["seconds"] = time.time() - t0
                if got == ():
                    return None
                res = self._as_dict(len(arglist), *got)
                if self.verbose:
                    for i in _dist.my_items(len(arglist)):
                        print(i, 'result ', res[i])
                return res
        mine = _dist.my_items(len(arglist))
        if self.engine != "python":
            try:
                done, err = self._run_native(arglist, mine), None
            except Exception as exc:             # a failing search on one rank: every rank still reaches the
                done, err = None, exc            # collective, and every rank raises
            if _dist.all_reduce_max(0 if err is None else 1):
                raise err if err is not None else RuntimeError("root search failed on another rank")
            if _dist.all_reduce_max(0 if done is not None else 1):
                done = None
            if done is not None:
                (roots, nr), st = done
                self.last_stats = dict(items=len(mine), passes=st.get("passes", 0),
                                       function_calls=st.get("function_calls", 0),
                                       evals=st.get("evals", 0), seconds=time.time() - t0,
                                       host_logic_seconds=st.get("logic_s"),
                                       device_seconds=st.get("eval_s"), launches=st.get("launches"),
                                       polish_evals=st.get("polish_evals"), stage_s=st.get("stage_s"),
                                       engine=self.engine)
                idx, mat, cnt = _dist.all_gather_dense(mine, roots, nr)
                # {index: roots}: root-less items (the majority on a map) share one read-only empty array
                none = np.zeros(0, dtype=np.complex128)
                none.flags.writeable = False
                res = dict.fromkeys(range(len(arglist)), none)
                for k in np.flatnonzero(cnt > 0).tolist():
                    res[int(idx[k])] = mat[k, :cnt[k]].copy()
                if self.verbose:
                    for i in mine:
                        print(i, 'result ', res[i])
                return res
        jobs, ctx = [], None
        for i in mine:
            job, trace, pm = self._job(arglist[i])
            jobs.append((i, job, trace))
            ctx = pm.psi_dispersion.ctx
        stats = {}
        if jobs:
            run_lockstep(ctx, [j for _, j, _ in jobs], stats=stats)
        local = {i: self._encode(j.result) for i, j, _ in jobs}
        if self.verbose:
            for i, j, _ in jobs:
                print(i, 'result ', j.result)
        self.last_stats = dict(items=len(jobs), passes=stats.get("passes", 0),
                               function_calls=sum(t.n_function_call for _, _, t in jobs),
                               seconds=time.time() - t0, engine="python")
        merged = _dist.all_gather_records(local)
        return {i: self._decode(merged[i]) for i in sorted(merged)}

    def run(self, arglist):
        res = self.run_all(arglist, root_only=True)
        return res if self.rank == 0 else None


class DispersionRelationMpiScheduler(MpiScheduler):
    """Values of the dispersion relation for items {'__init__': ..., 'dispersion': {'w', 'wave_number_x',
    'wave_number_z', 'viscous_alpha'}} (reference psi_mode_mpi.py:181-192): all items of a rank go
    through ONE K1 launch."""

    def runcompute(self, args):
        pm = self._cache.get(args['__init__'])
        return pm.dispersion(**args['dispersion'])

    def run_all(self, arglist, root_only=False):
        """{index: D}: every item's 'w' may be a scalar or an array of any shape (the reference hands it straight to
        PSIMode.dispersion, psi_mode_mpi.py:185-192); all frequencies of a rank's items go through ONE K1 launch and
        come back in the shape they were given."""
        mine = _dist.my_items(len(arglist))
        local = {}
        if mine:
            pms = [self._cache.get(arglist[i]['__init__']) for i in mine]
            d = [arglist[i]['dispersion'] for i in mine]
            ctx = pms[0].psi_dispersion.ctx
            ws = [np.asarray(x['w'], dtype=np.complex128) for x in d]
            size = np.array([w.size for w in ws], dtype=np.int64)
            per_point = lambda vals, dt: np.repeat(np.asarray(vals, dtype=dt), size)
            vals = ctx.disp_eval(per_point([p.psi_dispersion.dist_id for p in pms], np.int32),
                                 per_point([float(x['wave_number_x']) for x in d], float),
                                 per_point([float(x['wave_number_z']) for x in d], float),
                                 per_point([float(x.get('viscous_alpha', 0)) for x in d], float),
                                 np.concatenate([w.ravel() for w in ws]) if ws else np.zeros(0, complex))
            at = np.concatenate(([0], np.cumsum(size)))
            self._shapes = {i: w.shape for i, w in zip(mine, ws)}
            local = {i: np.concatenate(([float(w.ndim)], np.asarray(w.shape, dtype=float),
                                        vals[at[k]:at[k + 1]].view(np.float64)))
                     for k, (i, w) in enumerate(zip(mine, ws))}
        merged = _dist.all_gather_records(local)
        out = {}
        for i in sorted(merged):
            rec = merged[i]
            nd = int(rec[0])
            shape = tuple(int(v) for v in rec[1:1 + nd])
            val = np.ascontiguousarray(rec[1 + nd:]).view(np.complex128).reshape(shape)
            out[i] = np.squeeze(val) if nd == 0 else val
        return out
