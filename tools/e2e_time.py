"""End-to-end timing of ba_core.run_ba_optimization on the bench workload (host buffers in and out), repeated."""
import os
import sys
import time


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sat_bundleadjust_b200 import ba_core  # noqa: E402

p = bench.build_problem(sys.argv[1] if len(sys.argv) > 1 else "cfg2", 1)
ls = dict(bench.LS, max_iter=300)
for rep in range(4):
    t = time.perf_counter()
    out = ba_core.run_ba_optimization(p, ls, False, False, return_info=True)
    wall = time.perf_counter() - t
    info = out[-1]
    print("rep %d wall %.1f ms  %s  its %d" % (rep, wall * 1e3, {k: round(v * 1e3, 2) for k, v in info["wall_s"].items()}, info["iterations"]), flush=True)
