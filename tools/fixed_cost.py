"""Fixed cost of the pattern-engine passes: times each pass with the unit walk disabled (SBA_PT_SKIP=1)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from sat_bundleadjust_b200.solver import DeviceProblem, initial_vars
p = bench.build_problem(sys.argv[1] if len(sys.argv) > 1 else "1m", 1)
prob = DeviceProblem(p)
x = torch.from_numpy(initial_vars(p)).cuda()
out = torch.empty_like(x)
for it in (6,):
    info = prob.solve_device(x.data_ptr(), out.data_ptr(), None, ftol=0.0, xtol=0.0, gtol=0.0, max_nfev=10 ** 6, max_iterations=it, timed_from=1, **bench.LS)
    print({k: round(v / max(1, info["timed_iterations"]), 4) for k, v in info["phase_ms"].items()}, info["iter_ms"] / max(1, info["timed_iterations"]))
