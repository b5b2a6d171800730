"""N > 1 path on the CPU: world_size 2, gloo.  Items are interleaved over the ranks, each rank
solves its share (lane-emulation backend) and one all-gather of fixed-slot records gives every rank
the full result -- the same code path that runs with NCCL on the GPU box."""
import json
import os
import subprocess
import sys

import numpy as np

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
