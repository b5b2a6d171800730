"""Multi-GPU plumbing: one process per GPU (torchrun), work items interleaved across ranks, results
exchanged with ONE all-gather of fixed-slot float64 records per phase (NCCL over NVLink on GPUs,
gloo in the CPU test-suite).  There is no data-path collective inside the kernels: every (Kx, Kz,
omega) item is independent; this replaces the reference's pickled MPI
point-to-point traffic (psi_mode_mpi.py:57-160) and the per-sweep bcast of psi_grid_refine.py:325-326.
"""
import numpy as np


def _td():
    try:
        import torch.distributed as td
    except Exception:
        return None
    return td if td.is_available() and td.is_initialized() else None


def world_size():
    td = _td()
    return td.get_world_size() if td else 1


def rank():
    td = _td()
    return td.get_rank() if td else 0


def my_items(n_items):
    """Indices owned by this rank: interleaved (i mod world == rank) so that the strongly
    K-dependent cost of neighbouring pixels is spread over all GPUs."""
    return list(range(rank(), n_items, world_size()))


def _device_for_collectives():
    import torch
    td = _td()
    if td is not None and td.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def all_gather_records(local):
    """``local``: dict {global index: 1-D float64 array}.  Returns the union over all ranks.

    Records are packed as rows [index, length, payload..., padding] of a common width (one
    all-reduce MAX for the width and the row count, one all-gather for the data)."""
    td = _td()
    if td is None or td.get_world_size() == 1:
        return dict(local)
    import torch
    dev = _device_for_collectives()
    width = max([len(v) for v in local.values()] + [0])
    shape = torch.tensor([len(local), width], dtype=torch.int64, device=dev)
    td.all_reduce(shape, op=td.ReduceOp.MAX)
    rows, width = int(shape[0]), int(shape[1])
    buf = np.full((rows, width + 2), -1.0)
    for r, (idx, v) in enumerate(sorted(local.items())):
        buf[r, 0] = idx
        buf[r, 1] = len(v)
        buf[r, 2:2 + len(v)] = v
    mine = torch.from_numpy(buf).to(dev)
    parts = [torch.empty_like(mine) for _ in range(td.get_world_size())]
    td.all_gather(parts, mine)
    out = {}
    for p in parts:
        a = p.cpu().numpy()
        for row in a:
            if row[0] >= 0:
                out[int(row[0])] = row[2:2 + int(row[1])].copy()
    return out


def all_gather_dense(idx, mat, cnt):
    """Dense variant for big maps: ``idx`` (n,) global indices, ``mat`` (n, R) complex128 root slots,
    ``cnt`` (n,) number of valid slots per row.  Returns the concatenation over all ranks (rows in rank
    order); one all-reduce for the common shape, one all-gather for the data, no per-item Python."""
    td = _td()
    idx = np.asarray(idx, dtype=np.int64)
    cnt = np.asarray(cnt, dtype=np.int64)
    mat = np.asarray(mat, dtype=np.complex128)
    mat = mat.reshape(len(idx), -1) if len(idx) else np.zeros((0, max(1, mat.shape[-1] if mat.ndim == 2 else 1)), np.complex128)
    if td is None or td.get_world_size() == 1:
        return idx, mat, cnt
    import torch
    dev = _device_for_collectives()
    shape = torch.tensor([len(idx), mat.shape[1]], dtype=torch.int64, device=dev)
    td.all_reduce(shape, op=td.ReduceOp.MAX)
    rows, width = int(shape[0]), int(shape[1])
    buf = np.zeros((rows, 2 + 2*width))
    buf[:, 0] = -1.0
    buf[:len(idx), 0] = idx
    buf[:len(idx), 1] = cnt
    buf[:len(idx), 2:2 + 2*mat.shape[1]] = mat.view(np.float64)
    mine = torch.from_numpy(buf).to(dev)
    parts = [torch.empty_like(mine) for _ in range(td.get_world_size())]
    td.all_gather(parts, mine)
    allbuf = torch.cat(parts).cpu().numpy()
    allbuf = allbuf[allbuf[:, 0] >= 0]
    return (allbuf[:, 0].astype(np.int64), np.ascontiguousarray(allbuf[:, 2:]).view(np.complex128),
            allbuf[:, 1].astype(np.int64))


def all_gather_ragged(cnt, flat):
    """Per-rank ragged results: ``cnt`` (n_local,) values per local item, ``flat`` their concatenation
    (complex128).  Returns the lists [cnt of rank 0, ...], [flat of rank 0, ...].  Three collectives: the two
    lengths, the counts, the values -- only items that HAVE values cost bandwidth (on a growth-rate map most
    searches return nothing)."""
    td = _td()
    cnt = np.ascontiguousarray(cnt, dtype=np.int64)
    flat = np.ascontiguousarray(flat, dtype=np.complex128)
    if td is None or td.get_world_size() == 1:
        return [cnt], [flat]
    import torch
    dev = _device_for_collectives()
    world = td.get_world_size()
    sizes = torch.tensor([len(cnt), len(flat)], dtype=torch.int64, device=dev)
    all_sizes = [torch.empty_like(sizes) for _ in range(world)]
    td.all_gather(all_sizes, sizes)
    all_sizes = [[int(v) for v in t.tolist()] for t in all_sizes]
    n_max = max(a for a, _ in all_sizes)
    f_max = max(b for _, b in all_sizes)
    cbuf = np.zeros(max(n_max, 1), dtype=np.int64)
    cbuf[:len(cnt)] = cnt
    fbuf = np.zeros(2*max(f_max, 1))
    fbuf[:2*len(flat)] = flat.view(np.float64)
    mine_c, mine_f = torch.from_numpy(cbuf).to(dev), torch.from_numpy(fbuf).to(dev)
    parts_c = [torch.empty_like(mine_c) for _ in range(world)]
    parts_f = [torch.empty_like(mine_f) for _ in range(world)]
    td.all_gather(parts_c, mine_c)
    td.all_gather(parts_f, mine_f)
    cnts = [p.cpu().numpy()[:a] for p, (a, _) in zip(parts_c, all_sizes)]
    flats = [np.ascontiguousarray(p.cpu().numpy()[:2*b]).view(np.complex128) for p, (_, b) in zip(parts_f, all_sizes)]
    return cnts, flats


_NCCL_CACHE = {}


def all_reduce_max(value):
    """Maximum of a small non-negative integer over all ranks (error / fallback flags ahead of a collective)."""
    td = _td()
    if td is None or td.get_world_size() == 1:
        return int(value)
    import torch
    t = torch.tensor([int(value)], dtype=torch.int64, device=_device_for_collectives())
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return int(t.item())


def nccl_active():
    """True when results can be exchanged device to device (torch.distributed with the NCCL backend, > 1 rank)."""
    td = _td()
    if td is None:
        return False
    key = id(td.group.WORLD)
    if _NCCL_CACHE.get("key") != key:                       # (get_backend costs milliseconds: asked once per group)
        _NCCL_CACHE.update(key=key, on=td.get_world_size() > 1 and td.get_backend() == "nccl")
    return _NCCL_CACHE["on"]


def all_gather_device_records(rec, n_items, to_host=True):
    """ONE ncclAllGather of fixed-slot result records straight out of the library's device buffer (``rec``: a
    _device.DeviceRecords with the same number of rows on every rank; rank q owns items q::world, so row j of rank
    q is item q + j*world).  What comes to the host is compacted on the device first -- on a growth-rate map most
    searches return nothing: the root count of every item (int32) and the root slots of the items that have any.
    Returns a dict: ``cnt`` (n_items,), ``idx`` (indices of the items with roots), ``roots`` (len(idx), max_roots)
    complex -- all three None when ``to_host`` is false -- and, reduced on the device so that EVERY rank can act on
    them, ``top`` (largest root count), ``worst`` (largest status) and ``calls`` (sum of n_function_call)."""
    import torch
    td = _td()
    world = td.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device())
    mine = torch.as_tensor(rec, device=dev)                       # zero-copy view of the library's buffer
    out = torch.empty((world, rec.rows, rec.width), dtype=torch.float64, device=dev)
    td.all_gather_into_tensor(out.view(-1), mine.reshape(-1))
    head = torch.stack((out[:, :, 0].max(), out[:, :, 1].max(), out[:, :, 2].sum())).cpu()
    res = dict(top=int(head[0]), worst=int(head[1]), calls=int(head[2]), cnt=None, idx=None, roots=None)
    if to_host:
        table = out.permute(1, 0, 2).reshape(world*rec.rows, rec.width)[:n_items]      # global item order
        cnt = table[:, 0].to(torch.int32)
        has = cnt > 0
        res["cnt"] = cnt.cpu().numpy().astype(np.int64)
        res["idx"] = np.flatnonzero(res["cnt"] > 0)
        res["roots"] = np.ascontiguousarray(table[has][:, 4:].contiguous().cpu().numpy()).view(np.complex128)
    return res


def barrier():
    td = _td()
    if td is not None:
        td.barrier()
