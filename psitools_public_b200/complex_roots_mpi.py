"""Root-count maps with the interface of psitools.complex_roots_mpi.MpiScheduler (reference
complex_roots_mpi.py:32-175): items {'PSIDispersion.__init__': ..., 'dispersion': {wave numbers},
'z_list': contour, 'count_roots': kwargs} -> number of roots inside the contour, or the sentinel
-666 when the count did not converge (:171-174).  All contours of a rank are refined in lock-step,
one K1 launch per refinement pass."""
import numpy as np

from . import _dist
from ._engine import Job, run_lockstep
from .complex_roots import contour_machine
from .psi_dispersion import PSIDispersion
from .psi_mode_mpi import _init_key

FAILED_COUNT = -666


class MpiScheduler:
    def __init__(self, wall_start, wall_limit_total, verbose=True, engine="device"):
        self.engine = engine          # "device": kernel K2; "python": contour_machine generators + K1
        self.last_stats = {}
        self.verbose = verbose
        self.rank = _dist.rank()
        self.nranks = _dist.world_size()
        self.wall_start = wall_start
        self.wall_limit_total = wall_limit_total
        self._disp = {}

    def _dispersion(self, init):
        key = _init_key(init)
        hit = self._disp.get(key)
        if hit is None:
            hit = self._disp[key] = (PSIDispersion(**init), init)
        return hit[0]

    def _job(self, args):
        pd = self._dispersion(args['PSIDispersion.__init__'])
        kw = dict(args.get('count_roots', {}))
        kw.pop('verbose', None)
        d = args['dispersion']
        job = Job(contour_machine(args['z_list'], **kw), pd.dist_id, d['wave_number_x'],
                  d['wave_number_z'], d.get('viscous_alpha', 0))
        return job, pd

    def runcompute(self, args):
        job, pd = self._job(args)
        run_lockstep(pd.ctx, [job], capture_errors=True)
        return self._value(job)

    @staticmethod
    def _value(job):
        if job.error is None:
            return int(job.result)
        if isinstance(job.error, RuntimeError):
            print('complex_roots_mpi caught RuntimeError in count_roots', job.error)
            print('Likely means the max iterations was exceeded.')
            return FAILED_COUNT
        raise job.error

    def _run_native(self, arglist, mine):
        pds = [self._dispersion(arglist[i]['PSIDispersion.__init__']) for i in mine]
        if not pds:
            return {}
        ctx = pds[0].ctx
        if any(pd.ctx is not ctx for pd in pds) or not hasattr(ctx, "contour_batch"):
            return None
        kw = [dict(arglist[i].get('count_roots', {})) for i in mine]
        d = [arglist[i]['dispersion'] for i in mine]
        counts, status, npts, st = ctx.contour_batch(
            [pd.dist_id for pd in pds], [float(x['wave_number_x']) for x in d],
            [float(x['wave_number_z']) for x in d], [float(x.get('viscous_alpha', 0)) for x in d],
            [arglist[i]['z_list'] for i in mine], [k.get('tol', 0.1) for k in kw],
            [k.get('max_step', 0.1) for k in kw], [k.get('max_iter', 100) for k in kw])
        if np.any(status == 3):
            raise ValueError("Infinite value in f_add")
        if np.any(status == 4):
            raise RuntimeError("contour needs more than 4096 points")
        for s_ in status:
            if s_ in (1, 2):
                print('complex_roots_mpi caught RuntimeError in count_roots')
                print('Likely means the max iterations was exceeded.')
        self.last_stats = dict(st, items=len(mine), points=int(np.sum(npts)), engine="device")
        return {i: np.array([float(c)]) for i, c in zip(mine, counts)}

    def run_all(self, arglist):
        mine = _dist.my_items(len(arglist))
        if self.engine != "python":
            local = self._run_native(arglist, mine)
            if local is not None:
                merged = _dist.all_gather_records(local)
                return {i: int(merged[i][0]) for i in sorted(merged)}
        jobs, ctx = [], None
        for i in mine:
            job, pd = self._job(arglist[i])
            jobs.append((i, job))
            ctx = pd.ctx
        if jobs:
            run_lockstep(ctx, [j for _, j in jobs], capture_errors=True)
        local = {i: np.array([float(self._value(j))]) for i, j in jobs}
        merged = _dist.all_gather_records(local)
        return {i: int(merged[i][0]) for i in sorted(merged)}

    def run(self, arglist):
        res = self.run_all(arglist)
        return res if self.rank == 0 else None
