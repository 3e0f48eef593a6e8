"""
RPC refit after bundle adjustment -- drop-in mirror of the reference's bundle_adjust/ba_rpcfit.py on the GPU.

    weighted_lsq(target, input_locs, h, tol, max_iter)                       ba_rpcfit.py:88-153
    fit_Rt_corrected_rpc(Rt_vec, global_transform, original_rpc, crop_offset, pts3d_ba, n_samples)     :270-345
    fit_rpc_from_projection_matrix(P, global_transform, original_rpc, crop_offset, pts3d_ba, n_samples) :201-267
    check_errors(rpc_calib, input_locs, target)                              :359-370
    fit_Rt_corrected_rpcs(...)   batched over cameras (no reference counterpart: the reference loops in Python)

The sampling grid, the localisation of the grid with the original RPC, the corrective mapping
X' = R (X - T - C) + C, the projection of the corrected points and the regularised IRLS fit all run in sm_100a
kernels (csrc/sba_rpc.cu, csrc/sba_rpcfit.cu).  The reference's "does the fit cover the whole image" test uses shapely
(intersection area of the image rectangle with the convex hull of the reprojected samples == the rectangle's area,
:348-356); a rectangle lies inside a convex hull exactly when its four corners do, which is what is tested here.
"""
import ctypes

import numpy as np

from . import _lib, cam_utils, geo_utils
from .rpc_model import RPCModel


def _model_from_table(t):
    r = RPCModel()
    (r.row_offset, r.col_offset, r.lat_offset, r.lon_offset, r.alt_offset,
     r.row_scale, r.col_scale, r.lat_scale, r.lon_scale, r.alt_scale) = [float(v) for v in t[:10]]
    r.row_num, r.row_den, r.col_num, r.col_den = [t[10 + 20 * i: 30 + 20 * i].copy() for i in range(4)]
    return r


def weighted_lsq_batch(targets, input_locs, h=1e-3, tol=1e-2, max_iter=20):
    """targets (B,N,2), input_locs (B,N,3) -> list of B RPCModel, iterations (B,), rmse (B,)"""
    t = np.ascontiguousarray(targets, dtype=np.float64)
    x = np.ascontiguousarray(input_locs, dtype=np.float64)
    B, N = t.shape[0], t.shape[1]
    out = np.empty((B, 90))
    it = np.empty(B, dtype=np.int32)
    rmse = np.empty(B)
    lib = _lib.load()
    _lib.check(lib.sba_rpcfit_weighted_lsq(_lib.dptr(t), _lib.dptr(x), B, N, h, tol, max_iter, _lib.dptr(out),
                                           it.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _lib.dptr(rmse)))
    return [_model_from_table(out[b]) for b in range(B)], it, rmse


def weighted_lsq(target, input_locs, h=1e-3, tol=1e-2, max_iter=20):
    """Regularised iterative weighted least squares of one RPC (same arguments and result as the reference)."""
    models, _, _ = weighted_lsq_batch(target[None], input_locs[None], h, tol, max_iter)
    return models[0]


def check_errors(rpc_calib, input_locs, target, plot=False):
    col, row = rpc_calib.projection(input_locs[:, 0], input_locs[:, 1], input_locs[:, 2])
    return np.linalg.norm(np.hstack([col.reshape(-1, 1), row.reshape(-1, 1)]) - target, axis=1)


def check_correspondences_are_good(target, image_corners):
    """True when the image rectangle lies inside the convex hull of the (N,2) pixel coordinates `target`."""
    from scipy.spatial import ConvexHull
    hull = ConvexHull(target)
    A, b = hull.equations[:, :2], hull.equations[:, 2]
    return bool(np.all(image_corners @ A.T + b <= 1e-9 * max(1.0, np.abs(image_corners).max())))


def adjust_pts3d(pts3d, Rt_vec):
    """X' = R (X - T - C) + C (ba_core.py:110-130), host form used by the sampling driver"""
    Rt = np.asarray(Rt_vec, dtype=np.float64).reshape(-1, 9)
    q = pts3d - Rt[:, 3:6] - Rt[:, 6:9]
    ca, sa, cb, sb, cg, sg = [f(Rt[:, k]) for k in range(3) for f in (np.cos, np.sin)]
    y1, z1 = ca * q[:, 1] - sa * q[:, 2], sa * q[:, 1] + ca * q[:, 2]
    x2, z2 = cb * q[:, 0] + sb * z1, -sb * q[:, 0] + cb * z1
    x3, y3 = cg * x2 - sg * y1, sg * x2 + cg * y1
    return np.stack((x3, y3, z2), axis=1) + Rt[:, 6:9]


def _fit_loop(project_fn, original_rpc, crop_offset, alt_range, global_transform, n_samples):
    x0, y0, w, h = crop_offset["col0"], crop_offset["row0"], crop_offset["width"], crop_offset["height"]
    corners = np.array([[x0, y0], [x0, y0 + h], [x0 + w, y0 + h], [x0 + w, y0]], dtype=np.float64)
    margin = 10
    while True:
        cols, lins, alts = cam_utils.generate_point_mesh([x0 - margin, x0 + w + margin, n_samples],
                                                         [y0 - margin, y0 + h + margin, n_samples], alt_range)
        lons, lats = original_rpc.localization(cols, lins, alts)
        x, y, z = geo_utils.latlon_to_ecef_custom(lats, lons, alts)
        pts3d = np.vstack([x, y, z]).T
        if global_transform is not None:
            pts3d = pts3d + global_transform
        target = project_fn(pts3d)
        input_locs = np.vstack([lons, lats, alts]).T
        rpc_calib = weighted_lsq(target, input_locs)
        err = check_errors(rpc_calib, input_locs, target)
        reproj = cam_utils.apply_rpc_projection(rpc_calib, np.vstack([x, y, z]).T)
        if margin > 1000 or check_correspondences_are_good(reproj, corners):
            return rpc_calib, err, margin
        margin *= 2


def fit_Rt_corrected_rpc(Rt_vec, global_transform, original_rpc, crop_offset, pts3d_ba, n_samples=10):
    """New RPC reproducing  x = P_rpc( R (X - T - C) + C )  (same arguments / returns as the reference)."""
    pts = pts3d_ba - global_transform if global_transform is not None else pts3d_ba
    _, _, alts = geo_utils.ecef_to_latlon_custom(pts[:, 0], pts[:, 1], pts[:, 2])
    dev = abs(original_rpc.alt_offset - np.median(alts))
    if dev > 5:
        print("warning: median altitude of bundle adjustment points is {:.2f} meters deviated from the original rpc "
              "alt_offset".format(dev))
    alt_range = [original_rpc.alt_offset - original_rpc.alt_scale, original_rpc.alt_offset + original_rpc.alt_scale, n_samples]
    Rt = np.asarray(Rt_vec, dtype=np.float64).reshape(1, 9)
    return _fit_loop(lambda X: cam_utils.apply_rpc_projection(original_rpc, adjust_pts3d(X, Rt)), original_rpc, crop_offset,
                     alt_range, global_transform, n_samples)


def fit_rpc_from_projection_matrix(P, global_transform, original_rpc, crop_offset, pts3d_ba, n_samples=10):
    """New RPC reproducing a 3x4 projection matrix P (same arguments / returns as the reference)."""
    pts = pts3d_ba - global_transform if global_transform is not None else pts3d_ba
    _, _, alts = geo_utils.ecef_to_latlon_custom(pts[:, 0], pts[:, 1], pts[:, 2])
    alt_offset, alt_scale = np.median(alts), max(8000, original_rpc.alt_scale)
    x0, y0 = crop_offset["col0"], crop_offset["row0"]
    return _fit_loop(lambda X: cam_utils.apply_projection_matrix(P, X) + np.array([x0, y0]), original_rpc, crop_offset,
                     [alt_offset - alt_scale, alt_offset + alt_scale, n_samples], global_transform, n_samples)
