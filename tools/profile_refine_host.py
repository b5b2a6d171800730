"""cProfile of a full PSIGridRefiner campaign (host side): where the Python time goes."""
import cProfile
import io
import contextlib
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_extra  # noqa: E402

fills = int(sys.argv[1]) if len(sys.argv) > 1 else 5
pr = cProfile.Profile()
pr.enable()
with contextlib.redirect_stdout(io.StringIO()) as buf:
    bench_extra.refine(16, fills, 3)
pr.disable()
print(buf.getvalue().splitlines()[-1][:600])
pstats.Stats(pr).sort_stats('tottime').print_stats(22)
