import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, util
from oracle import ba_oracle
from sat_bundleadjust_b200.solver import DeviceProblem, initial_vars
G = util.load_ba_golden()
p = util.params_from_golden(G, "persp_RT_huber")
x0 = initial_vars(p)
with DeviceProblem(p) as prob:
    x, r, info = prob.solve(x0, loss="huber", f_scale=2.0, ftol=1e-14, xtol=1e-14, max_nfev=400, verbose=2)
    print({k: v for k, v in info.items() if k != "phase_ms"})
    print("oracle cost at x:", ba_oracle.robust_cost(ba_oracle.residuals(x.copy(), p), "huber", 2.0), "conv", float(G["persp_RT_huber/conv_cost"]))
