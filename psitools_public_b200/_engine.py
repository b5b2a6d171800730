"""Lock-step execution of many root-search / contour state machines over ONE batched GPU call per
pass.

Every machine is a generator that yields a 1-D complex array of frequencies at which it needs the
dispersion relation and receives the values; its distribution id and (Kx, Kz, alpha) are fixed.
``run_lockstep`` concatenates the requests of all live machines, evaluates them with a single K1
launch (``Context.disp_eval`` with per-point parameter arrays), scatters the values back and repeats
until every machine has returned.  This replaces the reference's one-process-per-work-item loop
(psi_mode_mpi.py:130-178): there is no per-point host round trip, only one per pass.
"""
import contextlib

import numpy as np

try:                                    # tiny LAPACK calls (AAA) crawl when BLAS spins up all its threads
    from threadpoolctl import threadpool_limits
except ImportError:                     # pragma: no cover
    threadpool_limits = None


def single_threaded_blas():
    if threadpool_limits is None:
        return contextlib.nullcontext()
    return threadpool_limits(limits=1, user_api="blas")


class Job:
    """One state machine plus the parameters of its dispersion relation."""
    __slots__ = ("machine", "dist_id", "kx", "kz", "alpha", "result", "error", "n_calls")

    def __init__(self, machine, dist_id, kx, kz, alpha):
        self.machine = machine
        self.dist_id = int(dist_id)
        self.kx = float(kx)
        self.kz = float(kz)
        self.alpha = float(alpha)
        self.result = None
        self.error = None
        self.n_calls = 0


def _advance(job, values, capture_errors):
    """Send values (None = start) into the machine; returns the next request or None when done."""
    try:
        req = next(job.machine) if values is None else job.machine.send(values)
        return np.asarray(req, dtype=np.complex128).ravel()
    except StopIteration as stop:
        job.result = stop.value
    except Exception as exc:  # the reference lets exceptions kill the work item
        if not capture_errors:
            raise
        job.error = exc
    return None


def run_lockstep(ctx, jobs, capture_errors=False, stats=None):
    """Run all jobs to completion.  Results end up in ``job.result`` (or ``job.error``)."""
    with single_threaded_blas():
        return _run_lockstep(ctx, jobs, capture_errors, stats)


def _run_lockstep(ctx, jobs, capture_errors, stats):
    pending = []
    for job in jobs:
        req = _advance(job, None, capture_errors)
        if req is not None:
            pending.append((job, req))
    passes = 0
    while pending:
        sizes = [len(r) for _, r in pending]
        w = np.concatenate([r for _, r in pending])
        if len(w):
            dist = np.repeat(np.array([j.dist_id for j, _ in pending], dtype=np.int32), sizes)
            kx = np.repeat(np.array([j.kx for j, _ in pending]), sizes)
            kz = np.repeat(np.array([j.kz for j, _ in pending]), sizes)
            al = np.repeat(np.array([j.alpha for j, _ in pending]), sizes)
            vals = ctx.disp_eval(dist, kx, kz, al, w)
        else:
            vals = np.zeros(0, dtype=np.complex128)
        passes += 1
        nxt = []
        off = 0
        for (job, req), n in zip(pending, sizes):
            job.n_calls += n
            new = _advance(job, vals[off:off + n], capture_errors)
            off += n
            if new is not None:
                nxt.append((job, new))
        pending = nxt
    if stats is not None:
        stats["passes"] = stats.get("passes", 0) + passes
    return jobs
