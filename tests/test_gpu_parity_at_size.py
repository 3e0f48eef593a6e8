"""
Convergence parity at the benchmark sizes (run with -m gpu on the B200 box), all through the C ABI.

north_star: ">= 1M-observation BA converging to the reference's cost within 1e-6 relative".  The reference's own stopping
point is path dependent at the 1e-4 level (ftol 1e-4, forward differences, inexact LSMR steps -- SURVEY.md H1), so the
claim is made at a stationary point:
  * the GPU solver is run to tight tolerances at BASELINE config 2 size (~5e5 observations) and at the metric's size
    (~1e6 observations);
  * the ORACLE (numpy port of ba_core.fun, pinned bit-exactly to the reference) then judges that point with the
    reference's own machinery -- scipy's 2-point finite differences over the reference's sparsity pattern, scipy's robust
    rescale -- : the gradient J^T f must be at the finite-difference noise floor (measured ~1.4e-7 of the initial
    gradient on a converged small case; bar 1e-5), and scipy's TRF (the reference's solver call, ba_core.py:284-297)
    restarted from the GPU solution for a few evaluations must not lower the oracle's cost by more than 1e-6 relative;
  * the same for cam_model='rpc' against the oracle's scipy run on project_rpc (float32-rounded residual);
  * and for the multi-GPU solve (every visible GPU, >= 2) against the single-GPU solve.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import util
from oracle import ba_oracle
from sat_bundleadjust_b200 import ba_core, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def oracle_gradient(v, p, loss, f_scale, A, groups):
    """J^T f of the reference's cost at v: scipy's sparse 2-point differences of the oracle's residual + scipy's robust rescale."""
    from scipy.optimize._lsq.common import compute_grad, scale_for_robust_loss_function
    from scipy.optimize._lsq.least_squares import construct_loss_function
    from scipy.optimize._numdiff import approx_derivative
    f = ba_oracle.residuals(v.copy(), p)
    J = approx_derivative(ba_oracle.residuals, v.copy(), method="2-point", sparsity=(A, groups), args=(p,))
    if loss != "linear":
        rho = construct_loss_function(f.size, loss, f_scale)(f.copy())
        J, f = scale_for_robust_loss_function(J, f.copy(), rho)
    return compute_grad(J, f)


@pytest.mark.parametrize("n_tracks", [100000, 200000])
def test_convergence_parity_at_bench_size(built, n_tracks):
    from scipy.optimize import least_squares
    from scipy.optimize._numdiff import group_columns
    sc = synth.make_scene(n_cam=10, n_tracks=n_tracks, p_vis=0.5, cam_model="perspective", seed=0)
    p = synth.scene_to_params(sc, ["R", "T"])
    assert p.n_obs > 4.8 * n_tracks
    loss, fs = "soft_l1", 1.0
    tight = {"loss": loss, "f_scale": fs, "ftol": 1e-14, "xtol": 0.0, "max_iter": 3000, "verbose": 0}
    v0, v1, e0, e1, nfev, info = ba_core.run_ba_optimization(p, tight, False, False, return_info=True)
    assert info["status"] > 0 and nfev < 3000
    cost_gpu = ba_oracle.robust_cost(ba_oracle.residuals(v1.copy(), p), loss, fs)       # judged by the oracle
    assert abs(cost_gpu - info["cost"]) <= 1e-9 * cost_gpu
    A = ba_oracle.jacobian_sparsity(p)
    groups = group_columns(A)
    g0 = np.abs(oracle_gradient(v0, p, loss, fs, A, groups)).max()
    g1 = np.abs(oracle_gradient(v1, p, loss, fs, A, groups)).max()
    assert g1 <= 1e-5 * g0, (g1, g0)
    # the reference's solver, started from the GPU solution, finds nothing to gain
    res = least_squares(ba_oracle.residuals, v1.copy(), jac_sparsity=A, verbose=0, x_scale="jac", method="trf", ftol=1e-15,
                        xtol=1e-15, gtol=1e-15, loss=loss, f_scale=fs, max_nfev=4, args=(p,))
    assert cost_gpu - res.cost <= 1e-6 * cost_gpu, (cost_gpu, res.cost)
    # reprojection RMSE at that point, device vs oracle
    err_oracle = ba_oracle.reprojection_error(ba_oracle.residuals(v1.copy(), p), p.pts2d_w)
    assert abs(np.sqrt(np.mean(e1 ** 2)) - np.sqrt(np.mean(err_oracle ** 2))) <= 1e-6 * np.sqrt(np.mean(err_oracle ** 2))
    # default tolerances (what the pipeline runs; ftol 1e-4 stops 1e-4 .. 5e-3 above the stationary cost): never below it
    dflt = {"loss": loss, "f_scale": fs, "max_iter": 300, "verbose": 0}
    _, v2, _, _, _ = ba_core.run_ba_optimization(p, dflt, False, False)
    cost_dflt = ba_oracle.robust_cost(ba_oracle.residuals(v2.copy(), p), loss, fs)
    assert cost_gpu * (1 - 1e-9) <= cost_dflt <= cost_gpu * (1 + 1e-2), (cost_dflt, cost_gpu)


@pytest.mark.parametrize("corr", [["R"], ["R", "T"]])
def test_rpc_solve_parity_vs_oracle(built, corr):
    """
    cam_model='rpc' (the pipeline default, ba_pipeline.py:83).  The reference's residual is rounded to float32
    (ba_core.py:150): at ~3000 px that is a 2.4e-4 px quantum, which perturbs the cost of a ~0.5 px RMS problem by
    ~1e-7 relative, and makes the reference's finite-difference Jacobian noisy (it stalls above the minimum).  Bars:
    the oracle-evaluated cost at the GPU solution is not above the cost the reference's own solve reaches (+1e-5, the
    float32 noise allowance), and the reference's solver restarted from the GPU solution gains less than 1e-5 relative.
    """
    from scipy.optimize import least_squares
    R = util.load_rpc_golden()
    p = util.rpc_ba_params_from_golden(R, corr)
    ls = {"loss": "linear", "ftol": 1e-12, "xtol": 1e-14, "max_iter": 400, "verbose": 0}
    _, x_ref, _, err_ref, _ = ba_oracle.solve(p, ls)
    v0, x_gpu, e0, e1, nfev, info = ba_core.run_ba_optimization(p, ls, False, False, return_info=True)
    c_ref = ba_oracle.robust_cost(ba_oracle.residuals(x_ref.copy(), p))
    c_gpu = ba_oracle.robust_cost(ba_oracle.residuals(x_gpu.copy(), p))
    assert c_gpu <= c_ref * (1 + 1e-5), (c_gpu, c_ref)
    res = least_squares(ba_oracle.residuals, x_gpu.copy(), jac_sparsity=ba_oracle.jacobian_sparsity(p), verbose=0, x_scale="jac",
                        method="trf", ftol=1e-15, xtol=1e-15, gtol=1e-15, max_nfev=6, args=(p,))
    assert c_gpu - res.cost <= 1e-5 * c_gpu, (c_gpu, res.cost)
    rmse_gpu = np.sqrt(np.mean(ba_oracle.reprojection_error(ba_oracle.residuals(x_gpu.copy(), p), p.pts2d_w) ** 2))
    rmse_ref = np.sqrt(np.mean(err_ref ** 2))
    assert rmse_gpu <= rmse_ref * (1 + 1e-5)
    assert abs(np.sqrt(np.mean(e1 ** 2)) - rmse_gpu) <= 3e-4        # device errors carry the float32 rounding of `fun`


def test_multi_gpu_solve_matches_single_gpu(built):
    """Every visible GPU (>= 2): the sharded tight solve reaches the single-GPU minimum (spawns torchrun)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 8 if n >= 8 else (4 if n >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tools", "dist_check.py"), "--tight"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    sys.stdout.write(out.stdout[-4000:])
    assert out.returncode == 0, out.stderr[-4000:]
    assert "all checks: True" in out.stdout


def test_rpc_solve_at_size_judged_by_the_oracle(built):
    """
    cam_model='rpc' at 2.4e5 observations (synth.make_rpc_scene: the golden fixture's four RPC cameras, synthetic tracks).  The
    oracle (numpy RPC projection pinned to the compiled reference, float32-rounded residual like ba_core.py:150) evaluates the
    GPU solution: its cost equals the device's own FP64 cost to the float32 noise of the oracle's residual (bar 1e-5 relative, as in
    test_rpc_solve_parity_vs_oracle; measured 3.5e-6), the solve converged
    (status 2), the cost fell by more than half and the median reprojection error is at the 0.5 px noise level.
    """
    G = util.load_rpc_golden()
    sc = synth.make_rpc_scene(G["rpc_cams"], G["rpcba/camera_centers"], n_tracks=80000, p_vis=0.8, seed=1)
    p = synth.SparseParams(sc, ["R", "T"])
    assert p.n_obs > 2.3e5
    ls = {"loss": "soft_l1", "f_scale": 1.0, "ftol": 1e-10, "xtol": 0.0, "max_iter": 600, "verbose": 0}
    v0, v1, e0, e1, nfev, info = ba_core.run_ba_optimization(p, ls, False, False, return_info=True)
    assert info["status"] > 0 and info["cost"] < 0.5 * info["cost_init"]
    q = synth.SparseParams(sc, ["R", "T"])
    q.cameras = [util.rpc_from_array(a) for a in G["rpc_cams"]]          # the oracle's own RPC objects (numpy), same coefficients
    c_oracle = ba_oracle.robust_cost(ba_oracle.residuals(v1.copy(), q), "soft_l1", 1.0)
    assert abs(c_oracle - info["cost"]) <= 1e-5 * c_oracle, (c_oracle, info["cost"])
    err_oracle = ba_oracle.reprojection_error(ba_oracle.residuals(v1.copy(), q), q.pts2d_w)
    assert abs(np.median(err_oracle) - np.median(e1)) < 1e-3 and 0.3 < np.median(e1) < 0.9
