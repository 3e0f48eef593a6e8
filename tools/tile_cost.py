"""
Calibration of the pattern engine's cost model (csrc/sba_pattern.h pattern_tile_cost): per-tile device time of the four kernels
as a function of the track length L, from scenes in which every track has exactly L observations (10 views, ~1e6 observations).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from sat_bundleadjust_b200 import synth  # noqa: E402
from sat_bundleadjust_b200.solver import DeviceProblem, initial_vars  # noqa: E402

M = 10
print("L  T  tracks  tiles | us per iteration: K2 jvp, K3 schur, K4 backsub, K1 assemble | ns per tile (x 148 CTAs): K2 K3 K4 K1")
for L in [int(a) for a in sys.argv[1:]] or [2, 3, 4, 5, 6, 7, 8, 10]:
    n_tracks = 1000000 // L

    def vis(rng, n, m, L=L):
        order = np.argsort(rng.random((n, m)), axis=1)[:, :L]
        v = np.zeros((n, m), dtype=bool)
        np.put_along_axis(v, order, True, axis=1)
        return v
    sc = synth.make_scene(n_cam=M, n_tracks=n_tracks, cam_model="perspective", seed=L, visibility=vis)
    p = synth.SparseParams(sc, ["R", "T"])
    T = min(32 // L, 16)
    with DeviceProblem(p, engine="pattern") as prob:
        x = torch.from_numpy(initial_vars(p)).cuda()
        out = torch.empty_like(x)
        info = prob.solve_device(x.data_ptr(), out.data_ptr(), None, ftol=0.0, xtol=0.0, gtol=0.0, max_nfev=10 ** 6, max_iterations=14,
                                 timed_from=2, **bench.LS)
        torch.cuda.synchronize()
    n = max(1, info["timed_iterations"])
    ph = {k: 1e3 * v / n for k, v in info["phase_ms"].items()}
    tiles = (p.n_pts + T - 1) // T
    sel = [ph["scale_jvp"], ph["schur"], ph["backsub"], ph["step_eval"]]
    print("%2d %2d %7d %6d | %6.1f %6.1f %6.1f %6.1f | %s" % (L, T, p.n_pts, tiles, *sel, " ".join("%7.1f" % (1e3 * v * 148 / tiles) for v in sel)), flush=True)
