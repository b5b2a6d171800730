"""N > 1 path on the CPU: world_size 2, gloo.  Items are interleaved over the ranks, each rank
solves its share (lane-emulation backend) and one all-gather of fixed-slot records gives every rank
the full result -- the same code path that runs with NCCL on the GPU box."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from common import carr, load
from test_host_logic import same_root_sets

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_sharding(tmp_path, built):
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29731",
           os.path.join(ROOT, "tests", "_rank_worker.py"), str(tmp_path)]
    subprocess.run(cmd, check=True, timeout=600, env=env, cwd=ROOT)
    outs = [json.load(open(tmp_path / ("rank%d.json" % r))) for r in range(2)]
    assert [o["world"] for o in outs] == [2, 2]
    assert outs[0]["mine"] == [0, 2, 4, 6, 8] and outs[1]["mine"] == [1, 3, 5, 7]
    assert outs[0]["mode_items_local"] == 5 and outs[1]["mode_items_local"] == 4
    assert all(o["run_is_none_on_workers"] for o in outs)
    gold = load("modegrid.json")["grid3"]["cases"]
    for o in outs:                                   # every rank holds the full, identical result
        assert sorted(map(int, o["mode"])) == list(range(9))
        for i, case in enumerate(gold):
            got = [complex(a, b) for a, b in o["mode"][str(i)]]
            assert same_root_sets(got, list(carr(case["roots"]))), (i, got)
        assert [o["counts"][str(i)] for i in range(4)] == [0, 0, 0, 1]
        np.testing.assert_allclose(o["disp"]["0"], [-1.72733464670814e+08, 5.32484293672187e+07], rtol=1e-7)
    assert outs[0]["mode"] == outs[1]["mode"]
    # the refinement campaign: identical on both ranks and identical to a single-rank run
    assert outs[0]["refine"] == outs[1]["refine"]
    assert outs[0]["refine"]["log"] == [[0, 12, 0], [1, 24, 22], [1, 24, 0]]
    single = _single_rank_refine()
    assert outs[0]["refine"]["runs"] == single["runs"]
    assert outs[0]["refine"]["roots"] == single["roots"]


def _single_rank_refine():
    import time
    import hostsim_util as hs
    from psitools_public_b200 import _device, PSIGridRefiner, TanhSinhNoDeepCopy, psi_mode_mpi
    cache, saved = {}, _device.get_context
    _device.get_context = lambda integ, device=None: cache.setdefault("c", hs.HostsimContext(integ))
    try:
        base = {'__init__': {'stokes_range': [1e-8, 1.0], 'dust_to_gas_ratio': 3.0, 'real_range': [-2.0, 2.0],
                             'imag_range': [1e-8, 1.0], 'n_sample': 20, 'max_zoom_domains': 1, 'tol': 1.0e-13,
                             'clean_tol': 1e-4, 'tanhsinh_integrator': TanhSinhNoDeepCopy()},
                'calculate': {'wave_number_x': None, 'wave_number_z': None, 'guess_roots': []},
                'random_seed': 2}
        gr = PSIGridRefiner('x', baseargs=base, nbase=(3, 3), reruns=1, krange=(1.0, 1.6),
                            scheduler=psi_mode_mpi.MpiScheduler(time.time(), 3600, verbose=False))
        gr.run_basegrid()
        gr.fill_in_grid()
        g = gr.grids[-1]
        return {"runs": g['runs'].tolist(),
                "roots": {str(k): [[z.real, z.imag] for z in g['results'][k]] for k in g['results']}}
    finally:
        _device.get_context = saved


@pytest.mark.gpu
def test_two_gpu_device_records(tmp_path, built):
    """world_size 2 over NCCL on two GPUs (skipped on a one-GPU box): the map and the refinement campaign, whose
    results are exchanged device to device (psi_mode_batch_device_rec + all_gather_into_tensor on the library's
    buffer), equal the single-GPU run bit for bit on both ranks."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    worker = os.path.join(ROOT, "tests", "_rank_worker_nccl.py")
    subprocess.run([sys.executable, worker, str(tmp_path)], check=True, timeout=900, cwd=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29741", worker, str(tmp_path)]
    subprocess.run(cmd, check=True, timeout=900, cwd=ROOT)
    one = json.load(open(tmp_path / "rank0_of1.json"))
    two = [json.load(open(tmp_path / ("rank%d_of2.json" % r))) for r in range(2)]
    assert one["exchange"] is None and one["map_with_roots"] > 50
    for o in two:
        assert o["world"] == 2 and o["exchange"] and "device records" in o["exchange"]
        assert o["refine_exchange"] and "device records" in o["refine_exchange"]
        assert o["map_hash"] == one["map_hash"] and o["map_with_roots"] == one["map_with_roots"]
        assert o["refine_log"] == one["refine_log"] and o["refine_runs"] == one["refine_runs"]
        assert o["refine_hash"] == one["refine_hash"]
        assert o["run_is_none"] == (o["rank"] != 0)
