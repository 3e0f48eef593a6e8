"""Prints the outcome of the full-size fixed-camera solve of tests/test_gpu_ba.py::test_full_size_properties."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sat_bundleadjust_b200 import ba_core, synth  # noqa: E402

sc = synth.make_scene(n_cam=10, n_tracks=100000, p_vis=0.5, cam_model="perspective", seed=0)
for fix in ((1, 100), (0, 0)):
    p = synth.scene_to_params(sc, ["R", "T"], n_cam_fix=fix[0], n_pts_fix=fix[1])
    for ftol in (1e-4, 1e-10):
        ls = {"loss": "soft_l1", "f_scale": 1.0, "max_iter": 300, "ftol": ftol, "xtol": 1e-12, "verbose": 0}
        v0, v1, e0, e1, nfev, info = ba_core.run_ba_optimization(p, ls, False, False, return_info=True)
        print(fix, ftol, {k: info[k] for k in ("status", "nfev", "iterations", "cost_init", "cost", "optimality", "chol_retries")},
              "rmse %.4f -> %.4f" % (np.sqrt(np.mean(e0 ** 2)), np.sqrt(np.mean(e1 ** 2))),
              "median %.4f -> %.4f" % (np.median(e0), np.median(e1)))
