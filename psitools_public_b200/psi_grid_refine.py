"""Adaptive growth-rate maps -- drop-in for psitools.psi_grid_refine.PSIGridRefiner (reference
psi_grid_refine.py:36-386).

A base (Kx, Kz) log grid is solved, pixels without roots are re-run with the roots of their eight
neighbours as guesses until nothing new is found, then the grid is refined to 2n-1 points per axis
(log mid-points) and the sweeps repeat.  Every batch of runs goes through ``MpiScheduler.run_all``:
the runs of a sweep are interleaved over the GPUs, advanced in lock-step on each, and the new roots
are all-gathered so that every rank keeps the same replicated root map for the next sweep (this is
the single exchange step per sweep that replaces the reference's bcast, :325-326).

Kept bug-for-bug (results must match the reference's): the upper pad cell is Kaxis[1]+dK (:155,159),
and the first run of every sweep re-uses the largest existing run id (offset = max(results), :226,
:276, :321), overwriting that run's roots.
Out of scope: prune_kmeans (dead code in the reference, :87-98, :272-273).
"""
import copy
import time
from collections.abc import MutableMapping

import numpy as np

from . import _dist
from . import psi_mode_mpi
from .psi_mode import guess_domain_size  # noqa: F401  (re-exported like the reference)


def spreadgrid(old, const_axis=0):
    """Stretch a meshgrid array to 2n-1 points per axis, inserting log10 mid-points along the
    varying axis.  For ``A, B = meshgrid(a, b, indexing='ij')``: A has const_axis=1, B const_axis=0."""
    new = np.zeros([2*n - 1 for n in old.shape])
    new[::2, ::2] = old
    if const_axis == 0:
        new[1:-1:2, ::2] = new[:-1:2, ::2]
        new[:, 1:-1:2] = 10**(0.5*(np.log10(new[:, :-2:2]) + np.log10(new[:, 2::2])))
    elif const_axis == 1:
        new[::2, 1:-1:2] = new[::2, :-1:2]
        new[1:-1:2, :] = 10**(0.5*(np.log10(new[:-2:2, :]) + np.log10(new[2::2, :])))
    else:
        raise ValueError('Bad const_axis')
    return new


def get_if(finishedruns, runkey):
    """Roots of run ``runkey`` as a list; the key -1 (no run yet) gives an empty list."""
    return list(finishedruns[runkey]) if runkey >= 0 else []


def prune_eps(values, epsilon=1e-5):
    """Crude pruning of root guesses: drop values within ``epsilon`` of an earlier kept one."""
    if values is None or len(values) <= 1:
        return values
    kept = [values[0]]
    for v in values[1:]:
        if not any(np.abs(v - u) < epsilon for u in kept):
            kept.append(v)
    return np.array(kept)


_NEIGHBOURS = ((-1, -1), (0, -1), (1, -1), (-1, 0), (1, 0), (-1, 1), (0, 1), (1, 1))   # (diz, dix)


class RootStore(MutableMapping):
    """The reference's ``results`` dict {run id: complex array of roots} held as two dense arrays (a root
    matrix and a count per run id, -1 = no such run), so that sweeps over million-pixel maps need no
    per-run Python objects.  Behaves like the dict for readers (indexing, iteration in ascending id order,
    ``max``, ``len``, ``in``)."""

    def __init__(self, source=None, width=4):
        self.mat = np.zeros((0, width), dtype=np.complex128)
        self.cnt = np.zeros(0, dtype=np.int64)
        if source is not None:
            for key in sorted(source):
                self[key] = source[key]

    def _reserve(self, n_keys, width):
        if n_keys > len(self.cnt) or width > self.mat.shape[1]:
            rows = max(n_keys, 2*len(self.cnt)) if n_keys > len(self.cnt) else len(self.cnt)
            # slots grow in steps of four: a wider row means a copy of the whole matrix (0.9 GB on a 1089^2 map) and a
            # re-opened device mirror, and the widest run of a campaign keeps creeping up by one
            width = -(-width//4)*4 if width > self.mat.shape[1] else self.mat.shape[1]
            mat = np.zeros((rows, max(width, self.mat.shape[1])), dtype=np.complex128)
            mat[:len(self.cnt), :self.mat.shape[1]] = self.mat
            cnt = np.full(rows, -1, dtype=np.int64)
            cnt[:len(self.cnt)] = self.cnt
            self.mat, self.cnt = mat, cnt

    @property
    def max_key(self):
        return int(np.flatnonzero(self.cnt >= 0)[-1])

    def copy(self):
        new = RootStore(width=self.mat.shape[1])
        new.mat, new.cnt = self.mat.copy(), self.cnt.copy()
        return new

    def counts(self, keys):
        """Root counts for an integer array of run ids (-1 = no run: 0 roots); KeyError on unknown ids."""
        keys = np.asarray(keys)
        safe = np.where(keys >= 0, keys, 0)
        if keys.size and (safe.max() >= len(self.cnt) or np.any((self.cnt[safe] < 0) & (keys >= 0))):
            raise KeyError('run id without stored result')
        return np.where(keys >= 0, self.cnt[safe], 0)

    def __getitem__(self, key):
        key = int(key)
        if key < 0 or key >= len(self.cnt) or self.cnt[key] < 0:
            raise KeyError(key)
        return self.mat[key, :self.cnt[key]].copy()

    def __setitem__(self, key, roots):
        key = int(key)
        roots = np.asarray(roots, dtype=np.complex128).ravel()
        self._reserve(key + 1, len(roots))
        self.mat[key, :len(roots)] = roots
        self.cnt[key] = len(roots)

    def __delitem__(self, key):
        self[key]
        self.cnt[int(key)] = -1

    def __iter__(self):
        return iter(np.flatnonzero(self.cnt >= 0).tolist())

    def __len__(self):
        return int(np.count_nonzero(self.cnt >= 0))


class PSIGridRefiner:
    def __init__(self, batchname, baseargs=None, nbase=(16, 16), reruns=3, verbose=False,
                 krange=(-1.0, 3.0), scheduler=None, tv_guesses=False):
        self.tv_guesses = tv_guesses
        if baseargs:
            self.baseargs = baseargs
        else:
            # defaults for testing (the reference's default Stokes range reads (1-8, 1e-1), an
            # obvious typo for 1e-8 that would make every run fail)
            self.baseargs = {'__init__': {'stokes_range': (1e-8, 1e-1), 'dust_to_gas_ratio': 10.0,
                                          'size_distribution_power': 3.5,
                                          'real_range': [-2.0, 2.0], 'imag_range': [1e-8, 1.0],
                                          'n_sample': 30, 'max_zoom_domains': 1, 'tol': 1.0e-13,
                                          'clean_tol': 1e-4},
                             'calculate': {'wave_number_x': None, 'wave_number_z': None,
                                           'guess_roots': []},
                             'random_seed': 2}
        self.nbase = nbase
        self.krange = krange
        self.reruns = reruns
        self.batchname = batchname
        self.datafile_hdf5 = batchname + '.hdf5'
        self.verbose = verbose
        self.wall_start = time.time()
        self.wall_limit_total = 70*60*60
        self.rank = _dist.rank()
        self.root = self.rank == 0
        self.ms = scheduler if scheduler is not None else \
            psi_mode_mpi.MpiScheduler(self.wall_start, self.wall_limit_total, verbose=self.verbose)
        self.grids = []
        self.sweep_log = []          # (grid index, runs launched, new roots) per sweep

    # -- helpers -------------------------------------------------------------------------------
    def _run(self, arglist):
        """All results on all ranks (scheduler shards, solves and all-gathers)."""
        if hasattr(self.ms, "run_all"):
            return self.ms.run_all(arglist)
        return self.ms.run(arglist)

    def _padded_axis(self, n):
        axis = np.logspace(self.krange[0], self.krange[1], n)
        dk = axis[1] - axis[0]
        return np.hstack(([axis[0] - dk], axis, [axis[1] + dk]))

    def _sweep_until_quiet(self, max_sweep=32):
        for _ in range(max_sweep):
            if self.sweep_last_grid() == 0:
                break

    def _array_path(self):
        """True when sweeps can use the array form end to end: the stock scheduler (not wrapped or
        subclassed with its own run_all), a batched engine, and a 'calculate' template that carries
        nothing but the wave numbers, the guesses and optionally viscous_alpha."""
        ms = self.ms
        if type(ms) is not psi_mode_mpi.MpiScheduler or 'run_all' in ms.__dict__ or ms.engine == "python":
            return False
        if self.verbose or self.baseargs['__init__'].get('verbose_flag', False):
            return False
        extra = set(self.baseargs['calculate']) - {'wave_number_x', 'wave_number_z', 'guess_roots',
                                                   'viscous_alpha'}
        return not extra and set(self.baseargs) <= {'__init__', 'calculate', 'random_seed'}

    def _shared_init(self, max_zoom):
        """One '__init__' dict per zoom setting, shared by every run of the campaign (the reference deep
        copies the whole arg-dict per run, :216-219; the copies are equal, and the scheduler keys its
        PSIMode cache on their content, so sharing changes nothing but the cost)."""
        cache = self.__dict__.setdefault('_init_cache', {})
        if max_zoom not in cache:
            init = copy.deepcopy(self.baseargs['__init__'])
            init['max_zoom_domains'] = max_zoom
            cache[max_zoom] = init
        return cache[max_zoom]

    def _item(self, max_zoom, kx, kz, guesses=None, seed_shift=0):
        calc = {k: copy.copy(v) for k, v in self.baseargs['calculate'].items()}
        calc['wave_number_x'] = kx
        calc['wave_number_z'] = kz
        if guesses is not None:
            calc['guess_roots'] = guesses
        args = {k: copy.deepcopy(v) for k, v in self.baseargs.items() if k not in ('__init__', 'calculate')}
        args['__init__'] = self._shared_init(max_zoom)
        args['calculate'] = calc
        if seed_shift:
            args['random_seed'] += seed_shift
        return args

    def _terminal_velocity_guesses(self, kx, kz):
        """Optional guess roots for the base map (``tv_guesses=True``): the mode of the
        terminal-velocity approximation at every pixel, computed for the whole map in one kernel
        (``PSIModeTV.find_roots_map``), kept where it exists, grows and lies inside the search domain.  Needs a
        power-law size distribution.  Returns the CSR pair (flat guesses, n+1 offsets)."""
        from .terminalvelocitysolver import PSIModeTV
        init = self.baseargs['__init__']
        if init.get('size_density') is not None or init.get('single_size_flag'):
            raise ValueError('tv_guesses needs a power-law size distribution')
        tv = PSIModeTV(init['dust_to_gas_ratio'], init['stokes_range'][1],
                       power_law_exponent_size_distribution=-init.get('size_distribution_power', 3.5))
        w, status = tv.find_roots_map(init['stokes_range'][0], kx, kz,
                                      viscous_alpha=self.baseargs['calculate'].get('viscous_alpha', 0),
                                      c_over_eta=init.get('sound_speed_over_eta', 1/0.05))
        re, im = init['real_range'], init['imag_range']
        keep = (status == 0) & (w.real > re[0]) & (w.real < re[1]) & (w.imag > max(im[0], 0.0)) & (w.imag < im[1])
        off = np.zeros(len(kx) + 1, dtype=np.int64)
        np.cumsum(keep, out=off[1:])
        return w[keep], off

    # -- public API ----------------------------------------------------------------------------
    def run_basegrid(self):
        kx_axis = self._padded_axis(self.nbase[1])
        kz_axis = self._padded_axis(self.nbase[0])
        kz_grid, kx_grid = np.meshgrid(kz_axis, kx_axis, indexing='ij')
        runs = np.arange(kx_grid.size, dtype=int).reshape(kx_grid.shape)
        results = None
        guesses = self._terminal_velocity_guesses(kx_grid.ravel(), kz_grid.ravel()) if self.tv_guesses else None
        if self._array_path():
            got = self.ms.run_arrays(self._shared_init(1), kx_grid.ravel(), kz_grid.ravel(),
                                     self.baseargs.get('random_seed', 0), guesses=guesses,
                                     alpha=self.baseargs['calculate'].get('viscous_alpha', 0))
            if got is not None:
                results = RootStore(width=got[0].shape[1])
                results._reserve(kx_grid.size, got[0].shape[1])
                results.mat[:kx_grid.size, :got[0].shape[1]] = got[0]
                results.cnt[:kx_grid.size] = got[1]
        if results is None:
            per = [None]*kx_grid.size if guesses is None else \
                [list(guesses[0][guesses[1][i]:guesses[1][i + 1]]) for i in range(kx_grid.size)]
            arglist = [self._item(1, kx, kz, g) for kx, kz, g in zip(kx_grid.ravel(), kz_grid.ravel(), per)]
            results = self._run(arglist)
        self.grids.append({'Kx': kx_grid, 'Kz': kz_grid, 'runs': runs, 'results': results,
                           'nguess': np.zeros(kx_grid.shape, dtype=int)})
        self._sweep_until_quiet()

    def fill_in_grid(self):
        old = self.grids[-1]
        kx_grid = spreadgrid(old['Kx'], const_axis=0)
        kz_grid = spreadgrid(old['Kz'], const_axis=1)
        runs = -1*np.ones(kx_grid.shape, dtype=int)
        runs[::2, ::2] = old['runs']
        # root arrays are replaced, never modified in place: a shallow copy of the map is a snapshot
        self.grids.append({'Kx': kx_grid, 'Kz': kz_grid, 'runs': runs,
                           'results': old['results'].copy(),
                           'nguess': np.zeros(kx_grid.shape, dtype=int)})
        if self.root:
            print(' running grid sweeps')
        self._sweep_until_quiet()

    @staticmethod
    def _root_counts(runs, results):
        """Number of roots per pixel (0 where no run exists); raises KeyError on a dangling run id like
        the reference's validity loop (:228-229)."""
        keys, inverse = np.unique(runs, return_inverse=True)
        per_key = np.array([len(results[k]) if k >= 0 else 0 for k in keys], dtype=int)
        return per_key[inverse].reshape(runs.shape)

    @staticmethod
    def _candidates(has):
        """Pixels without roots that touch (8-neighbourhood) a pixel with roots, in row-major order."""
        nz, nx = has.shape
        touched = np.zeros_like(has)
        for diz, dix in _NEIGHBOURS:
            src = has[max(0, diz):nz + min(0, diz), max(0, dix):nx + min(0, dix)]
            touched[max(0, -diz):nz + min(0, -diz), max(0, -dix):nx + min(0, -dix)] |= src
        return np.argwhere(touched & ~has)

    def _queue_numpy(self, grid):
        """Host form of the sweep's cell queue (the device form is csrc/psi_refine.cuh: refine_pixel): the
        pixels that run this sweep (row-major), their guess counts and the guesses, flat; updates nguess."""
        runs, nguess, store = grid['runs'], grid['nguess'], grid['results']
        nz, nx = runs.shape
        has = store.counts(runs) > 0
        cand = self._candidates(has)
        cz, cx = cand[:, 0], cand[:, 1]
        nc, width = len(cand), store.mat.shape[1]
        # neighbour roots in the reference's order: neighbour by neighbour, each run's roots in order
        near = np.full((nc, 8*width), np.nan + 0j)
        slot = np.arange(width)[None, :]
        for q, (diz, dix) in enumerate(_NEIGHBOURS):
            jz, jx = cz + diz, cx + dix
            ok = (jz >= 0) & (jz < nz) & (jx >= 0) & (jx < nx)
            key = np.where(ok, runs[np.clip(jz, 0, nz - 1), np.clip(jx, 0, nx - 1)], -1)
            ok &= key >= 0
            key = np.where(ok, key, 0)
            n_here = np.where(ok, store.cnt[key], 0)
            near[:, q*width:(q + 1)*width] = np.where(slot < n_here[:, None], store.mat[key], np.nan)
        valid = ~np.isnan(near.real)
        n_near = valid.sum(axis=1)
        order = np.argsort(~valid, axis=1, kind='stable')            # valid entries first, order kept
        longest = int(n_near.max()) if nc else 0
        near = np.take_along_axis(near, order[:, :longest], axis=1)
        # epsilon = mean(imag)*1e-4 over exactly the pixel's values (same summation as the list version)
        eps = np.zeros(nc)
        for length in np.unique(n_near):
            rows = np.flatnonzero(n_near == length)
            eps[rows] = np.ascontiguousarray(near[rows, :length]).imag.mean(axis=1)*1e-4
        kept = np.zeros((nc, max(longest, 1)), dtype=np.complex128)
        n_kept = np.zeros(nc, dtype=np.int64)
        rows_all = np.arange(nc)
        for col in range(longest):
            v = near[:, col]
            keep = col < n_near
            if col > 0:
                top = int(n_kept.max())
                close = np.abs(v[:, None] - kept[:, :top]) < eps[:, None]
                close &= np.arange(top)[None, :] < n_kept[:, None]
                keep &= ~(close.any(axis=1) & (n_near > 1))           # a single value is never pruned
            r = rows_all[keep]
            kept[r, n_kept[r]] = v[r]
            n_kept[r] += 1
        go = n_kept > nguess[cz, cx]
        cz, cx, kept, n_kept = cz[go], cx[go], kept[go], n_kept[go]
        nguess[cz, cx] = n_kept
        flat = kept[np.arange(kept.shape[1])[None, :] < n_kept[:, None]]          # row-major: pixel by pixel
        return cz, cx, n_kept, flat

    def _close_device_map(self):
        dev = self.__dict__.pop('_dev', None)
        if dev is not None:
            dev[1].close()

    def _device_map(self, grid):
        """RefineMap mirror of the last grid on this rank's GPU, opened lazily and re-opened when the
        RootStore outgrew it; None when there is no device library behind the scheduler (CPU harness) or
        ``device_sweep`` is off."""
        if not getattr(self, 'device_sweep', True):
            return None
        from . import _device
        ctx = self.ms._cache.get(self._shared_init(0)).psi_dispersion.ctx
        store = grid['results']
        if not isinstance(ctx, _device.Context):
            return None
        held = self.__dict__.get('_dev')
        if held is not None and held[0] is grid and held[1].width == store.mat.shape[1]:
            return held[1]
        self._close_device_map()
        n_keys = store.max_key + 1
        capacity = n_keys + 2*(1 + self.reruns)*grid['runs'].size
        dev = _device.RefineMap(ctx.device, grid['runs'], grid['nguess'], store.mat[:n_keys], store.cnt[:n_keys],
                                capacity)
        self._dev = (grid, dev)
        return dev

    def _sweep_arrays(self, grid):
        """sweep_last_grid on arrays: neighbour roots are gathered, pruned (prune_eps, same greedy order)
        and compared with ``nguess`` for all candidate pixels at once, the runs go to the scheduler as
        arrays (run_arrays) and their roots are merged into the RootStore by run id exactly as the
        dict-based path does (first run of a pixel assigns, its reruns append)."""
        kx_grid, kz_grid = grid['Kx'], grid['Kz']
        runs, nguess, store = grid['runs'], grid['nguess'], grid['results']
        offset = store.max_key
        nz, nx = runs.shape
        dev = self._device_map(grid)
        queue = dev.sweep() if dev is not None else None
        if queue is not None:                     # the cell queue comes from the device mirror of the map
            cz, cx, n_kept, flat_one = queue
            nguess[cz, cx] = n_kept
        else:
            if dev is not None:                   # a pixel overflowed the kernel's neighbour buffer: this
                self._close_device_map()          # sweep runs on the host, the mirror is rebuilt afterwards
                dev = None
            cz, cx, n_kept, flat_one = self._queue_numpy(grid)
        n_run, per = len(cz), 1 + self.reruns
        if self.root:
            print('Will run ', n_run*per, ' new points')
        newroots = 0
        if n_run:
            first = np.arange(n_run)*per
            bump = np.zeros_like(runs)
            bump[cz, cx] = offset + first - runs[cz, cx]
            starts = np.zeros(n_run + 1, dtype=np.int64)
            np.cumsum(n_kept, out=starts[1:])
            # every pixel's guess list once per (re)run
            lengths = np.repeat(n_kept, per)
            off = np.zeros(n_run*per + 1, dtype=np.int64)
            np.cumsum(lengths, out=off[1:])
            take = np.repeat(np.repeat(starts[:-1], per) - off[:-1], lengths) + np.arange(off[-1])
            seeds = self.baseargs.get('random_seed', 0) + np.tile(np.arange(per), n_run)
            got = self.ms.run_arrays(self._shared_init(0), np.repeat(kx_grid[cz, cx], per),
                                     np.repeat(kz_grid[cz, cx], per), seeds, guesses=(flat_one[take], off),
                                     alpha=self.baseargs['calculate'].get('viscous_alpha', 0))
            if got is None:
                raise RuntimeError('scheduler refused the array batch')
            mat, cnt = got
            newroots = int(cnt.sum())
            total = cnt.reshape(n_run, per).sum(axis=1)
            keys = offset + first
            store._reserve(int(keys[-1]) + 1, int(total.max()))
            store.cnt[keys] = 0
            pos = np.zeros(n_run, dtype=np.int64)
            for q in range(per):
                c = cnt[first + q]
                for t in range(int(c.max()) if n_run else 0):
                    sel = np.flatnonzero(c > t)
                    store.mat[keys[sel], pos[sel] + t] = mat[first[sel] + q, t]
                pos += c
            store.cnt[keys] = pos
            runs += bump
            store.counts(runs)                        # every pixel still maps to a stored run
            if dev is not None:
                if store.mat.shape[1] == dev.width and int(keys[-1]) < dev.key_capacity:
                    dev.update(keys, store.cnt[keys], store.mat[keys], cz, cx, runs[cz, cx])
                else:                                 # the map outgrew its mirror: re-opened next sweep
                    self._close_device_map()
        self.sweep_log.append((len(self.grids) - 1, n_run*per, newroots))
        return newroots

    def sweep_last_grid(self):
        """Re-run every root-less pixel that has gained neighbour roots; returns #new roots.

        Same visiting order, guess lists and run ids as the reference (:221-334); only pixels without
        roots that touch a pixel with roots can launch a run, and those are found with array shifts
        instead of a Python loop over the whole map."""
        grid = self.grids[-1]
        if self._array_path():
            if not isinstance(grid['results'], RootStore):
                grid['results'] = RootStore(grid['results'])
            done = self._sweep_arrays(grid)
            if done is not None:
                return done
        kx_grid, kz_grid = grid['Kx'], grid['Kz']
        runs, nguess, results = grid['runs'], grid['nguess'], grid['results']
        offset = max(results)
        nz, nx = kz_grid.shape
        count = self._root_counts(runs, results)
        has = count > 0
        candidates = self._candidates(has)                # row-major: the reference's iz, ix order

        arglist = []
        bump = np.zeros_like(runs)
        for iz, ix in candidates:
            near = []
            for diz, dix in _NEIGHBOURS:
                jz, jx = iz + diz, ix + dix
                if 0 <= jz < nz and 0 <= jx < nx and has[jz, jx]:
                    near += list(results[runs[jz, jx]])
            near = prune_eps(near, epsilon=np.array(near).imag.mean()*1e-4)
            if len(near) > nguess[iz, ix]:            # only when more guesses than last time
                nguess[iz, ix] = len(near)
                first = len(arglist)
                bump[iz, ix] = offset + first - runs[iz, ix]
                arglist.append(self._item(0, kx_grid[iz, ix], kz_grid[iz, ix], near))
                for k in range(self.reruns):
                    args = self._item(0, kx_grid[iz, ix], kz_grid[iz, ix], near, seed_shift=k + 1)
                    args['is_rerun'] = first          # its roots are appended to the first run's
                    arglist.append(args)
        if self.root:
            print('Will run ', len(arglist), ' new points')

        newroots = 0
        if arglist:
            fresh = self._run(arglist)
            for irun in sorted(fresh):
                if 'is_rerun' in arglist[irun]:
                    key = offset + arglist[irun]['is_rerun']
                    if len(results[key]) == 0 and len(fresh[irun]) > 0 and self.root and self.verbose:
                        print('combined ', results[key], fresh[irun])
                    if len(fresh[irun]):
                        results[key] = np.append(results[key], fresh[irun])
                else:
                    results[offset + irun] = fresh[irun]
                newroots += len(fresh[irun])
            runs += bump
            self._root_counts(runs, results)          # every pixel still maps to a stored run
            grid['runs'], grid['results'] = runs, results
        self.sweep_log.append((len(self.grids) - 1, len(arglist), newroots))
        return newroots

    # -- output ("next" row: HDF5 writer, reference psi_grid_refine.py:336-386) -----------------
    def grid_arrays(self, grid):
        """Dense arrays of one grid: Kxgrid, Kzgrid, root_real, root_imag (depth = largest root
        multiplicity) and error (= number of roots per pixel)."""
        store, runs = grid['results'], grid['runs']
        if isinstance(store, RootStore):               # array form: no per-pixel Python
            n = store.counts(runs)
            depth = max(1, int(n.max()))
            roots = np.where(np.arange(depth)[None, None, :] < n[:, :, None],
                             store.mat[np.where(runs >= 0, runs, 0)][:, :, :depth], 0.0)
            return {'Kxgrid': np.array(grid['Kx']), 'Kzgrid': np.array(grid['Kz']),
                    'root_real': roots.real.copy(), 'root_imag': roots.imag.copy(),
                    'error': np.where(runs >= 0, n, 0).astype(int)}
        depth = max([1] + [len(grid['results'][r]) for r in grid['runs'].ravel() if r >= 0])
        shape = list(grid['runs'].shape) + [depth]
        re, im = np.zeros(shape), np.zeros(shape)
        err = -1*np.ones_like(grid['runs'], dtype=int)
        for (iz, ix), _ in np.ndenumerate(grid['Kz']):
            roots = get_if(grid['results'], grid['runs'][iz, ix])
            err[iz, ix] = len(roots)
            for k, r in enumerate(roots):
                re[iz, ix, k], im[iz, ix, k] = r.real, r.imag
        return {'Kxgrid': np.array(grid['Kx']), 'Kzgrid': np.array(grid['Kz']),
                'root_real': re, 'root_imag': im, 'error': err}

    def to_hdf5(self):
        """Write all grids to ``batchname.hdf5`` (groups batchname/<igrid> with the datasets of
        grid_arrays and the run attributes).  Without h5py the same arrays go to ``batchname.npz``
        under the keys '<igrid>/<dataset>'."""
        if not self.root:
            return None
        print('Writing ', len(self.grids), 'grids to hdf5')
        init = self.baseargs['__init__']
        try:
            import h5py
        except ImportError:
            flat = {}
            for i, g in enumerate(self.grids):
                for name, arr in self.grid_arrays(g).items():
                    flat['%d/%s' % (i, name)] = arr
            path = self.batchname + '.npz'
            np.savez(path, run_walltime=time.time() - self.wall_start, **flat)
            return path
        with h5py.File(self.datafile_hdf5, 'w') as h5f:
            top = h5f.create_group(self.batchname)
            top.attrs['run_walltime'] = time.time() - self.wall_start
            for i, g in enumerate(self.grids):
                self.write_grid(i, g, h5f)
        return self.datafile_hdf5

    def write_grid(self, igrid, grid, h5f):
        """One grid into an open HDF5 file (reference psi_grid_refine.py:347-386): group ``batchname/igrid`` with the
        run attributes and the datasets Kxgrid, Kzgrid, root_real, root_imag (depth = largest root multiplicity)
        and error (= number of roots per pixel, -1 where no run exists)."""
        init = self.baseargs['__init__']
        grp = h5f.create_group(self.batchname + '/' + str(igrid))
        grp.attrs['stokes_range'] = init['stokes_range']
        grp.attrs['dust_to_gas_ratio'] = init['dust_to_gas_ratio']
        if 'viscous_alpha' in self.baseargs['calculate']:
            grp.attrs['viscous_alpha'] = self.baseargs['calculate']['viscous_alpha']
        grp.attrs['real_range'] = init['real_range']
        grp.attrs['imag_range'] = init['imag_range']
        arrays = self.grid_arrays(grid)
        for name in ('Kxgrid', 'Kzgrid', 'root_real', 'root_imag', 'error'):
            grp.create_dataset(name, data=arrays[name])
