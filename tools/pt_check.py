"""Pattern engine against the generic engine on the same problems (GPU box): blocks, reduced system, solves, timing."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from sat_bundleadjust_b200 import synth  # noqa: E402
from sat_bundleadjust_b200.solver import DeviceProblem, initial_vars  # noqa: E402


def with_engine(name, fn):
    os.environ["SBA_ENGINE"] = name
    try:
        return fn()
    finally:
        os.environ.pop("SBA_ENGINE", None)


def check(n_cam, n_tracks, model, corr, loss, ncf=0, npf=0, seed=3, p_vis=0.5):
    sc = synth.make_scene(n_cam=n_cam, n_tracks=n_tracks, p_vis=p_vis, cam_model=model, seed=seed)
    p = synth.scene_to_params(sc, corr, n_cam_fix=ncf, n_pts_fix=npf)
    x0 = initial_vars(p)
    out = {}
    for eng in ("generic", "pattern"):
        def run():
            with DeviceProblem(p) as prob:
                r, c = prob.residuals(x0, loss, 1.0)
                U, V, g = prob.normal_blocks(x0, loss, 1.0)
                S, rhs = prob.reduced_system(x0, loss, 1.0, 0.37)
                Jc, Jp = prob.jacobian_blocks(x0)
                t0 = time.perf_counter()
                x, rr, info = prob.solve(x0, loss=loss, ftol=1e-10, xtol=0.0, max_nfev=200)
                dt = time.perf_counter() - t0
                return dict(r=r, c=c, U=U, V=V, g=g, S=S, rhs=rhs, Jc=Jc, Jp=Jp, x=x, rr=rr, info=info, dt=dt)
        out[eng] = with_engine(eng, run)
    a, b = out["generic"], out["pattern"]

    def rel(u, v):
        return float(np.abs(u - v).max() / max(1e-300, np.abs(u).max()))
    print("%s M=%d N=%d c=%s %s fix=(%d,%d): r %.1e U %.1e V %.1e g %.1e S %.1e rhs %.1e Jc %.1e Jp %.1e | cost %.12e / %.12e nfev %d/%d it %d/%d status %d/%d ms %.2f/%.2f launches %d/%d"
          % (model, n_cam, p.n_pts, corr, loss, ncf, npf, rel(a["r"], b["r"]), rel(a["U"], b["U"]), rel(a["V"], b["V"]), rel(a["g"], b["g"]),
             rel(a["S"], b["S"]), rel(a["rhs"], b["rhs"]), rel(a["Jc"], b["Jc"]), rel(a["Jp"], b["Jp"]),
             a["info"]["cost"], b["info"]["cost"], a["info"]["nfev"], b["info"]["nfev"], a["info"]["iterations"], b["info"]["iterations"],
             a["info"]["status"], b["info"]["status"], a["info"]["solve_ms"], b["info"]["solve_ms"],
             a["info"]["gpu_launches"], b["info"]["gpu_launches"]), flush=True)
    assert rel(a["U"], b["U"]) < 1e-8 and rel(a["V"], b["V"]) < 1e-8 and rel(a["g"], b["g"]) < 1e-8
    assert rel(a["S"], b["S"]) < 1e-8 and rel(a["rhs"], b["rhs"]) < 1e-7
    if a["info"]["status"] > 0 and b["info"]["status"] > 0:        # both converged (ftol 1e-10): same minimum
        assert abs(a["info"]["cost"] - b["info"]["cost"]) < 1e-6 * a["info"]["cost"]


if __name__ == "__main__":
    check(4, 80, "perspective", ["R", "T"], "linear")
    check(6, 2000, "perspective", ["R", "T"], "soft_l1")
    check(6, 2000, "perspective", ["R", "T"], "soft_l1", ncf=1, npf=50)
    check(5, 1500, "affine", ["R", "T"], "huber")
    check(5, 1500, "perspective", ["R"], "linear")
    check(12, 3000, "perspective", ["R", "T"], "soft_l1", p_vis=0.7)      # tracks of 8+ observations: multi-pass Schur
    check(3, 3000, "affine", ["R"], "cauchy", p_vis=0.8)
    check(10, 100000, "perspective", ["R", "T"], "soft_l1", seed=0)
    print("pattern engine checks passed")
