"""cProfile of rank 0 of a multi-GPU PSIGridRefiner campaign (torchrun): where the host time goes.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/profile_refine_ranks.py [fills]"""
import cProfile
import contextlib
import io
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench_extra  # noqa: E402

fills = int(sys.argv[1]) if len(sys.argv) > 1 else 6
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if "RANK" in os.environ:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
with contextlib.redirect_stdout(io.StringIO()):
    bench_extra.refine(16, 1, 3)          # warm-up: contexts, communicator
pr = cProfile.Profile()
pr.enable()
with contextlib.redirect_stdout(io.StringIO()) as buf:
    bench_extra.refine(16, fills, 3)
pr.disable()
if local == 0:
    print(buf.getvalue().splitlines()[-1][:900])
    pstats.Stats(pr).sort_stats('tottime').print_stats(18)
if "RANK" in os.environ:
    dist.barrier()
    dist.destroy_process_group()
