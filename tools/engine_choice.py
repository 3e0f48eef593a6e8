"""Per-iteration device time of both engines on scenes whose tracks share their camera sets less and less (tile fill of the pattern layout)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from sat_bundleadjust_b200 import synth  # noqa: E402
from sat_bundleadjust_b200.solver import DeviceProblem, initial_vars  # noqa: E402

for n_cam, n_tracks, p_vis in ((10, 100000, 0.5), (14, 100000, 0.4), (18, 100000, 0.3), (22, 100000, 0.25)):
    sc = synth.make_scene(n_cam=n_cam, n_tracks=n_tracks, p_vis=p_vis, cam_model="perspective", seed=0)
    p = synth.scene_to_params(sc, ["R", "T"])
    res = {}
    for eng in ("auto", "pattern", "generic"):
        if eng == "auto":
            os.environ.pop("SBA_ENGINE", None)
        else:
            os.environ["SBA_ENGINE"] = eng
        with DeviceProblem(p) as prob:
            x = torch.from_numpy(initial_vars(p)).cuda()
            out = torch.empty_like(x)
            info = prob.solve_device(x.data_ptr(), out.data_ptr(), None, ftol=0.0, xtol=0.0, gtol=0.0, max_nfev=10 ** 6, max_iterations=12,
                                     timed_from=2, no_phase_timing=True, **bench.LS)
            torch.cuda.synchronize()
            res[eng] = (prob.engine, info["iter_ms"] / max(1, info["timed_iterations"]), info["cost"])
    print("M=%d obs=%d p_vis=%.2f: " % (n_cam, p.n_obs, p_vis) + " | ".join("%s -> %s %.3f ms/it cost %.6e" % (k, v[0], v[1], v[2]) for k, v in res.items()), flush=True)
