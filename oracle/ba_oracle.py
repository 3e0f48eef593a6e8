"""
TEST INFRASTRUCTURE ONLY.  CPU restatement (numpy + scipy) of the reference's bundle-adjustment hot
path, used as the checker in tests/, in __graft_entry__.smoke() and as bench.py's CPU baseline.
Nothing under sat_bundleadjust_b200/ imports this module.

Parity pin: this restatement is checked against the UNMODIFIED reference code (imported through
oracle/ref_loader.py in the build container) and against the golden fixtures generated from it
(tests/golden/make_golden.py -> tests/golden/*.npz).  The reference's own test-suite has no
known-answer vectors for this path, so the pin is "outputs of the reference
itself run here".

What is restated, and where it lives in the reference:
  rotate_euler            bundle_adjust/ba_core.py:36-56     R = Rz Ry Rx applied as x-, y-, z-rotation
  project_perspective     bundle_adjust/ba_core.py:84-107
  project_affine          bundle_adjust/ba_core.py:59-81
  adjust_pts3d            bundle_adjust/ba_core.py:110-130    X' = R (X - T - C) + C
  project_rpc             bundle_adjust/ba_core.py:133-154    float32 output buffer (:150)
  residuals (`fun`)       bundle_adjust/ba_core.py:157-183
  unpack_variables        bundle_adjust/ba_params.py:221-257  (get_vars_ready_for_fun)
  jacobian_sparsity       bundle_adjust/ba_core.py:186-219
  solve                   bundle_adjust/ba_core.py:244-332    scipy.optimize.least_squares(trf, x_scale='jac',
                                                             jac_sparsity, 2-point differences, LSMR)
  reprojection_error      bundle_adjust/ba_core.py:335-349
The solver itself is third-party: scipy (unpinned in the reference's requirements.txt:10; 1.18.1 in
this image), call site ba_core.py:284-297.
"""
import numpy as np

from . import rpc_oracle


# ------------------------------------------------------------------------------------------------
# projection models
# ------------------------------------------------------------------------------------------------
def rotate_euler(X, ang):
    """Rows of X rotated by Rz(ang[:,2]) Ry(ang[:,1]) Rx(ang[:,0]); same operation order as the reference."""
    ca, sa = np.cos(ang[:, 0]), np.sin(ang[:, 0])
    cb, sb = np.cos(ang[:, 1]), np.sin(ang[:, 1])
    cg, sg = np.cos(ang[:, 2]), np.sin(ang[:, 2])
    x0, y0, z0 = X[:, 0], X[:, 1], X[:, 2]
    x1, y1, z1 = x0, ca * y0 - sa * z0, sa * y0 + ca * z0
    x2, y2, z2 = cb * x1 + sb * z1, y1, -sb * x1 + cb * z1
    x3, y3, z3 = cg * x2 - sg * y2, sg * x2 + cg * y2, z2
    return np.stack((x3, y3, z3), axis=1)


def project_perspective(pts3d, cam_params, pts_ind, cam_ind):
    c = cam_params[cam_ind]
    q = rotate_euler(pts3d[pts_ind], c[:, 0:3])
    q += c[:, 3:6]
    fx, fy, sk, cx, cy = c[:, 6], c[:, 7], c[:, 8], c[:, 9], c[:, 10]
    u = fx * q[:, 0] + sk * q[:, 1] + cx * q[:, 2]
    v = fy * q[:, 1] + cy * q[:, 2]
    return np.stack((u, v), axis=1) / q[:, 2, np.newaxis]


def project_affine(pts3d, cam_params, pts_ind, cam_ind):
    c = cam_params[cam_ind]
    q = rotate_euler(pts3d[pts_ind], c[:, 0:3])[:, :2]
    q += c[:, 3:5]
    fx, fy, sk = c[:, 5], c[:, 6], c[:, 7]
    u = fx * q[:, 0] + sk * q[:, 1]
    v = fy * q[:, 1]
    return np.stack((u, v), axis=1)


def adjust_pts3d(pts3d, Rt_vec):
    q = pts3d - Rt_vec[:, 3:6]
    q -= Rt_vec[:, 6:9]
    q = rotate_euler(q, Rt_vec[:, 0:3])
    q += Rt_vec[:, 6:9]
    return q


def project_rpc(pts3d, rpcs, cam_params, pts_ind, cam_ind):
    """rpcs: list of objects with `.projection(lon, lat, alt)` (e.g. oracle.rpc_oracle.RPCModel)."""
    q = adjust_pts3d(pts3d[pts_ind], cam_params[cam_ind])
    out = np.zeros((pts_ind.shape[0], 2), dtype=np.float32)   # float32 on purpose: ba_core.py:150
    for j in np.unique(cam_ind).tolist():
        sel = cam_ind == j
        lat, lon, alt = rpc_oracle.ecef_to_latlon(q[sel, 0], q[sel, 1], q[sel, 2])
        col, row = rpcs[j].projection(lon, lat, alt)
        out[sel] = np.stack((col, row), axis=1)
    return out


# ------------------------------------------------------------------------------------------------
# variable vector <-> model parameters
# ------------------------------------------------------------------------------------------------
def _nK(cam_model):
    return 3 if cam_model == "affine" else 5


def _common_K(p):
    return "K" in p.cam_params_to_optimize and "COMMON_K" in p.cam_params_to_optimize


def unpack_variables(v, p):
    """(pts3d (N,3), cam_params (M,P)) from the variable vector; frozen cameras are written into v."""
    n_params = p.n_params
    K = None
    if _common_K(p):
        nK = _nK(p.cam_model)
        K, v = v[:nK], v[nK:]
        n_params -= nK
    n_c = p.n_cam * n_params
    pts3d = v[n_c:].reshape((p.n_pts, 3)).copy()
    if p.n_pts_fix > 0:
        pts3d[: p.n_pts_fix] = p.pts3d[: p.n_pts_fix]
    block = v[:n_c].reshape((p.n_cam, n_params))
    if p.n_cam_fix > 0:
        block[: p.n_cam_fix] = p.cam_params[: p.n_cam_fix, :n_params]
    cam_params = np.hstack((block, p.cam_params[:, n_params:]))
    if K is not None:
        cam_params[:, -nK:] = K
    return pts3d, cam_params


def residuals(v, p):
    """Weighted reprojection residuals, interleaved (x0, y0, x1, y1, ...), shape (2K,)."""
    pts3d, cam_params = unpack_variables(v, p)
    if p.cam_model == "perspective":
        proj = project_perspective(pts3d, cam_params, p.pts_ind, p.cam_ind)
    elif p.cam_model == "affine":
        proj = project_affine(pts3d, cam_params, p.pts_ind, p.cam_ind)
    else:
        proj = project_rpc(pts3d, p.cameras, cam_params, p.pts_ind, p.cam_ind)
    return np.repeat(p.pts2d_w, 2, axis=0) * (proj - p.pts2d).ravel()


fun = residuals


def jacobian_sparsity(p):
    """
    0/1 structure of the (2K x n) Jacobian as scipy CSR (the reference returns the same matrix as LIL).
    Row 2k and 2k+1 carry ones at the camera columns  off + cam_ind[k]*c + s  and the point columns
    off + M*c + 3*pts_ind[k] + s ; with COMMON_K the first nK columns are dense.
    """
    from scipy.sparse import csr_matrix

    c = p.n_params
    K = p.pts_ind.size
    nK = _nK(p.cam_model)
    common = _common_K(p)
    if common:
        c -= nK
    off = nK if common else 0
    n = off + p.n_cam * c + p.n_pts * 3
    cols = [np.repeat(np.arange(off)[None, :], K, axis=0)] if common else []
    cols.append(off + p.cam_ind[:, None] * c + np.arange(c)[None, :])
    cols.append(off + p.n_cam * c + p.pts_ind[:, None] * 3 + np.arange(3)[None, :])
    cols = np.hstack(cols)                       # (K, width)
    width = cols.shape[1]
    cols2 = np.repeat(cols, 2, axis=0).ravel()   # both rows of an observation share the pattern
    indptr = np.arange(0, 2 * K * width + 1, width)
    A = csr_matrix((np.ones(cols2.size, dtype=int), cols2, indptr), shape=(2 * K, n))
    A.sum_duplicates()
    return A


def reprojection_error(res, pts2d_w=None):
    w = np.ones(res.size, dtype=np.float32) if pts2d_w is None else np.repeat(pts2d_w, 2, axis=0)
    return np.linalg.norm(np.abs(res / w).reshape(res.size // 2, 2), axis=1)


DEFAULT_LS = {"loss": "linear", "ftol": 1e-4, "xtol": 1e-10, "f_scale": 1.0, "max_iter": 300, "verbose": 1}


def solve(p, ls_params=None, return_result=False):
    """
    The reference solve: scipy TRF with a finite-difference Jacobian over the sparsity pattern.
    Returns (vars_init, vars_ba, err_init, err_ba, nfev) like ba_core.run_ba_optimization.
    """
    from scipy.optimize import least_squares

    cfg = dict(DEFAULT_LS)
    if ls_params:
        cfg.update({k: ls_params[k] for k in DEFAULT_LS if k in ls_params})
    x0 = p.params_opt.copy()
    r0 = residuals(x0, p)
    A = jacobian_sparsity(p)
    res = least_squares(residuals, x0, jac_sparsity=A, verbose=cfg["verbose"], x_scale="jac", method="trf",
                        ftol=cfg["ftol"], xtol=cfg["xtol"], loss=cfg["loss"], f_scale=cfg["f_scale"],
                        max_nfev=cfg["max_iter"], args=(p,))
    out = (x0, res.x, reprojection_error(r0, p.pts2d_w), reprojection_error(res.fun, p.pts2d_w), res.nfev)
    return out + (res,) if return_result else out


# ------------------------------------------------------------------------------------------------
# checker-only helpers (no reference counterpart)
# ------------------------------------------------------------------------------------------------
def robust_cost(f, loss="linear", f_scale=1.0):
    """0.5 * f_scale^2 * sum(rho((f/f_scale)^2)) -- scipy/optimize/_lsq/least_squares.py:183-240."""
    z = (f / f_scale) ** 2
    if loss == "linear":
        rho = z
    elif loss == "soft_l1":
        rho = 2 * (np.sqrt(1 + z) - 1)
    elif loss == "huber":
        rho = np.where(z <= 1, z, 2 * np.sqrt(z) - 1)
    elif loss == "cauchy":
        rho = np.log1p(z)
    elif loss == "arctan":
        rho = np.arctan(z)
    else:
        raise ValueError(loss)
    return 0.5 * f_scale ** 2 * np.sum(rho)


def dense_jacobian_fd(v, p, rel_step=1e-6):
    """Central-difference Jacobian, column by column (small problems only)."""
    v = np.asarray(v, dtype=np.float64)
    n = v.size
    r0 = residuals(v.copy(), p)
    J = np.zeros((r0.size, n))
    for j in range(n):
        h = rel_step * max(1.0, abs(v[j]))
        a, b = v.copy(), v.copy()
        a[j] += h
        b[j] -= h
        J[:, j] = (residuals(a, p) - residuals(b, p)) / (a[j] - b[j])
    return J


def solve_converged(p, loss="linear", f_scale=1.0, x_start=None, max_nfev=300):
    """
    Checker only: the reference's cost function driven to a true local minimum, so that a cost-parity
    claim at 1e-6 relative is meaningful (the reference's own stopping point -- ftol=1e-4,
    forward-difference Jacobian, inexact LSMR steps -- is path dependent at the 1e-4..1e-3 level, and with
    LSMR it does not reach the minimum even after thousands of evaluations).
    Same scipy TRF and x_scale='jac', but 3-point differences, the dense `exact` trust-region solver
    and tight tolerances; small problems only (dense SVD of the Jacobian).
    Returns (x, cost, scipy result).
    """
    from scipy.optimize import least_squares

    x0 = p.params_opt.copy() if x_start is None else np.array(x_start, dtype=np.float64)
    res = least_squares(residuals, x0, jac="3-point", x_scale="jac", method="trf", tr_solver="exact",
                        ftol=1e-15, xtol=1e-15, gtol=1e-15, loss=loss, f_scale=f_scale, max_nfev=max_nfev, args=(p,))
    return res.x, robust_cost(residuals(res.x.copy(), p), loss, f_scale), res
