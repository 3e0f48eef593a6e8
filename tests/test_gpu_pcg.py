"""
G5 (run with -m gpu): block-Jacobi PCG on the reduced camera system, matrix-free (csrc/sba_pcg.cuh), against the dense
Schur + Cholesky path of the same library on the same problems, and against the oracle's cost function.
The reference has no counterpart of either (it hands the whole sparse Jacobian to LSMR, scipy/optimize/_lsq/trf.py:485-500);
what must hold is that both paths produce the same trust-region steps (to the CG tolerance) and reach the same minimum.
"""
import os

import numpy as np
import pytest

from oracle import ba_oracle
from sat_bundleadjust_b200 import synth
from sat_bundleadjust_b200.solver import DeviceProblem, initial_vars

pytestmark = pytest.mark.gpu


def solve_with(solver, p, x0, **kw):
    old = os.environ.get("SBA_SOLVER")
    os.environ["SBA_SOLVER"] = solver
    os.environ["SBA_PCG_TOL"] = "1e-11"
    try:
        with DeviceProblem(p) as prob:
            assert prob.engine == "generic"
            return prob.solve(x0, **kw)
    finally:
        os.environ.pop("SBA_PCG_TOL", None)
        if old is None:
            os.environ.pop("SBA_SOLVER", None)
        else:
            os.environ["SBA_SOLVER"] = old


@pytest.mark.parametrize("model,loss,ncf", [("perspective", "soft_l1", 0), ("affine", "linear", 2)])
def test_pcg_matches_dense_at_n300(built, model, loss, ncf):
    """50 views x 6 unknowns (BASELINE config 3 camera count): first steps agree to the CG tolerance, same minimum."""
    sc = synth.make_scene(n_cam=50, n_tracks=8000, p_vis=0.1, cam_model=model, seed=5)
    p = synth.scene_to_params(sc, ["R", "T"], n_cam_fix=ncf)
    assert p.n_cam * p.n_params == 300 if model == "perspective" else True
    x0 = initial_vars(p)
    # a fixed number of evaluations: the iterates must agree (the steps solve the same linear systems; what is left is the CG
    # tolerance amplified along the weakly determined gauge directions of S over ~20 steps: measured 1e-5 of the step)
    xa, _, ia = solve_with("dense", p, x0, loss=loss, max_nfev=30, ftol=0.0, xtol=0.0, gtol=0.0)
    xb, _, ib = solve_with("pcg", p, x0, loss=loss, max_nfev=30, ftol=0.0, xtol=0.0, gtol=0.0)
    assert ia["pcg_solves"] == 0 and ib["pcg_solves"] >= 1 and ib["pcg_iterations"] > 0
    step = np.abs(xa - x0).max()
    assert step > 0 and np.abs(xa - xb).max() <= 1e-4 * step, (np.abs(xa - xb).max(), step)
    assert abs(ia["cost"] - ib["cost"]) <= 1e-5 * ia["cost"]
    # tight solves reach the same minimum of the oracle's cost function
    xa, _, ia = solve_with("dense", p, x0, loss=loss, ftol=1e-13, xtol=0.0, max_nfev=600)
    xb, _, ib = solve_with("pcg", p, x0, loss=loss, ftol=1e-13, xtol=0.0, max_nfev=600)
    ca = ba_oracle.robust_cost(ba_oracle.residuals(xa.copy(), p), loss, 1.0)
    cb = ba_oracle.robust_cost(ba_oracle.residuals(xb.copy(), p), loss, 1.0)
    assert ia["status"] > 0 and ib["status"] > 0
    assert abs(ca - cb) <= 1e-8 * ca, (ca, cb)
    assert abs(cb - ib["cost"]) <= 1e-9 * cb


def test_pcg_is_the_default_for_time_series_scale(built):
    """300 views (BASELINE config 4 camera count, 1800 unknowns): PCG is selected automatically, no (camera, track) table and
    no pair lists are built, and the solve reaches a stationary point of the oracle's cost."""
    sc = synth.make_scene(n_cam=300, n_tracks=20000, p_vis=0.02, cam_model="perspective", seed=9)
    p = synth.scene_to_params(sc, ["R", "T"])
    x0 = initial_vars(p)
    with DeviceProblem(p) as prob:
        x, r, info = prob.solve(x0, loss="soft_l1", ftol=1e-10, xtol=0.0, max_nfev=400)
    assert info["pcg_solves"] >= 1 and info["status"] > 0
    c = ba_oracle.robust_cost(ba_oracle.residuals(x.copy(), p), "soft_l1", 1.0)
    assert abs(c - info["cost"]) <= 1e-9 * c and c < 0.2 * info["cost_init"]
    # same problem through the dense path
    old = os.environ.get("SBA_SOLVER")
    os.environ["SBA_SOLVER"] = "dense"
    try:
        with DeviceProblem(p) as prob:
            xd, _, infod = prob.solve(x0, loss="soft_l1", ftol=1e-10, xtol=0.0, max_nfev=400)
    finally:
        os.environ.pop("SBA_SOLVER", None) if old is None else os.environ.__setitem__("SBA_SOLVER", old)
    cd = ba_oracle.robust_cost(ba_oracle.residuals(xd.copy(), p), "soft_l1", 1.0)
    assert infod["pcg_solves"] == 0
    assert abs(c - cd) <= 1e-6 * cd, (c, cd)
