"""Runs a few trust-region iterations of a bench workload (for ncu launch lists / captures)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from sat_bundleadjust_b200.solver import DeviceProblem, initial_vars  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
p = bench.build_problem(workload, 1)
prob = DeviceProblem(p)
x = torch.from_numpy(initial_vars(p)).cuda()
out = torch.empty_like(x)
info = prob.solve_device(x.data_ptr(), out.data_ptr(), None, ftol=0.0, xtol=0.0, gtol=0.0, max_nfev=10 ** 6,
                         max_iterations=iters, timed_from=1, **bench.LS)
torch.cuda.synchronize()
print({k: v for k, v in info.items() if k != "phase_ms"})
print({k: round(v / max(1, info["timed_iterations"]), 4) for k, v in info["phase_ms"].items()})
