"""
TEST INFRASTRUCTURE ONLY.  CPU restatement (numpy) of the reference's RPC refit,
bundle_adjust/ba_rpcfit.py: `poly_vect` (:17-44), normalisation (:47-74), `calculate_RMSE_row_col` (:77-85),
`weighted_lsq` (:88-153), `scaling_params` / `initialize_rpc` (:156-198) and the sampling driver
`fit_Rt_corrected_rpc` (:270-345) without its shapely coverage test.

Parity pin: `weighted_lsq` here is checked against the UNMODIFIED reference function, which runs in the build
container once the stubbed `rpcm` module hands out oracle.rpc_oracle.RPCModel (oracle/ref_loader.py); golden
vectors from that run are stored in tests/golden/rpcfit_golden.npz (tests/golden/make_golden.py).
"""
import numpy as np

from . import rpc_oracle


def poly_terms(lon, lat, alt):
    """the 19 non-constant monomials in RPC00B order, shape (19, N)"""
    return rpc_oracle.monomials(lon, lat, alt)[1:]


def scaling_params(v):
    lo, hi = min(v), max(v)
    scale = (hi - lo) / 2
    return scale, lo + scale


def rmse_row_col(rpc, input_locs, target):
    col, row = rpc.projection(input_locs[:, 0], input_locs[:, 1], input_locs[:, 2])
    mse_col, mse_row = np.mean((np.hstack([col.reshape(-1, 1), row.reshape(-1, 1)]) - target) ** 2, axis=0)
    return np.sqrt(np.mean([mse_col, mse_row]))


def weighted_lsq(target, input_locs, h=1e-3, tol=1e-2, max_iter=20, return_iters=False):
    rpc = rpc_oracle.RPCModel()
    rpc.row_scale, rpc.row_offset = scaling_params(target[:, 1])
    rpc.col_scale, rpc.col_offset = scaling_params(target[:, 0])
    rpc.lat_scale, rpc.lat_offset = scaling_params(input_locs[:, 1])
    rpc.lon_scale, rpc.lon_offset = scaling_params(input_locs[:, 0])
    rpc.alt_scale, rpc.alt_offset = scaling_params(input_locs[:, 2])
    reg = (h ** 2) * np.eye(39)
    C = ((target[:, 0] - rpc.col_offset) / rpc.col_scale)[:, None]
    R = ((target[:, 1] - rpc.row_offset) / rpc.row_scale)[:, None]
    lon = (input_locs[:, 0] - rpc.lon_offset) / rpc.lon_scale
    lat = (input_locs[:, 1] - rpc.lat_offset) / rpc.lat_scale
    alt = (input_locs[:, 2] - rpc.alt_offset) / rpc.alt_scale
    pv = poly_terms(lon, lat, alt).T
    one = np.ones((lon.shape[0], 1))
    MC = np.hstack([one, pv, -C * pv])
    MR = np.hstack([one, pv, -R * pv])

    def set_coefs(JR, JC):
        coefs = np.vstack([JR[:20], 1, JR[20:], JC[:20], 1, JC[20:]]).reshape(-1)
        rpc.row_num, rpc.row_den = coefs[:20], coefs[20:40]
        rpc.col_num, rpc.col_den = coefs[40:60], coefs[60:]
        return coefs

    JR = np.linalg.inv(MR.T @ MR) @ (MR.T @ R)
    JC = np.linalg.inv(MC.T @ MC) @ (MC.T @ C)
    coefs = set_coefs(JR, JC)
    rmse = rmse_row_col(rpc, input_locs, target)
    n_iter = 0
    for n_iter in range(1, max_iter + 1):
        wr = 1 / ((MR[:, :20] @ coefs[20:40]) ** 2)
        wc = 1 / ((MC[:, :20] @ coefs[60:80]) ** 2)
        JR = np.linalg.inv((MR.T * wr) @ MR + reg) @ ((MR.T * wr) @ R)
        JC = np.linalg.inv((MC.T * wc) @ MC + reg) @ ((MC.T * wc) @ C)
        coefs = set_coefs(JR, JC)
        prev, rmse = rmse, rmse_row_col(rpc, input_locs, target)
        if np.abs(prev - rmse) < tol:
            break
    return (rpc, n_iter, rmse) if return_iters else rpc


def point_mesh(col_range, row_range, alt_range):
    c, r, a = [np.linspace(v[0], v[1], v[2]) for v in (col_range, row_range, alt_range)]
    A, R, C = np.meshgrid(a, r, c, indexing="ij")
    return C.ravel(), R.ravel(), A.ravel()


def rt_corrected_samples(Rt_vec, original_rpc, crop_offset, margin=10, n_samples=10, global_transform=None):
    """The (target, input_locs) correspondences of ba_rpcfit.py:312-332 for one margin value."""
    from . import ba_oracle
    x0, y0, w, h = crop_offset["col0"], crop_offset["row0"], crop_offset["width"], crop_offset["height"]
    a0, a1 = original_rpc.alt_offset - original_rpc.alt_scale, original_rpc.alt_offset + original_rpc.alt_scale
    cols, lins, alts = point_mesh([x0 - margin, x0 + w + margin, n_samples], [y0 - margin, y0 + h + margin, n_samples],
                                  [a0, a1, n_samples])
    lons, lats = original_rpc.localization(cols, lins, alts)
    pts3d = np.stack(rpc_oracle.latlon_to_ecef(lats, lons, alts), axis=1)
    if global_transform is not None:
        pts3d = pts3d + global_transform
    adj = ba_oracle.adjust_pts3d(pts3d, np.tile(np.asarray(Rt_vec).reshape(1, 9), (pts3d.shape[0], 1)))
    target = original_rpc.project_ecef(adj)
    return target, np.stack([lons, lats, alts], axis=1), pts3d


def weighted_lsq_accurate(target, input_locs, n_passes, h=1e-3):
    """
    Checker only (no reference counterpart): the algorithm of ba_rpcfit.weighted_lsq (bundle_adjust/ba_rpcfit.py:88-153) with
    numerically accurate linear algebra -- QR least squares for the first, unregularised solve, np.linalg.solve for the
    re-weighted ridge systems -- and exactly `n_passes` re-weighting passes.  The reference multiplies by np.linalg.inv of
    matrices with condition numbers 4e16 / 6e11; the difference between the two is the reference's own numerical noise
    (up to ~5e-2 px on a held-out grid, tests/test_rpcfit.py), which bounds what "same RPC out" can mean for any other
    implementation.
    """
    rpc = rpc_oracle.RPCModel()
    rpc.row_scale, rpc.row_offset = scaling_params(target[:, 1])
    rpc.col_scale, rpc.col_offset = scaling_params(target[:, 0])
    rpc.lat_scale, rpc.lat_offset = scaling_params(input_locs[:, 1])
    rpc.lon_scale, rpc.lon_offset = scaling_params(input_locs[:, 0])
    rpc.alt_scale, rpc.alt_offset = scaling_params(input_locs[:, 2])
    reg = (h ** 2) * np.eye(39)
    C = ((target[:, 0] - rpc.col_offset) / rpc.col_scale)[:, None]
    R = ((target[:, 1] - rpc.row_offset) / rpc.row_scale)[:, None]
    lon = (input_locs[:, 0] - rpc.lon_offset) / rpc.lon_scale
    lat = (input_locs[:, 1] - rpc.lat_offset) / rpc.lat_scale
    alt = (input_locs[:, 2] - rpc.alt_offset) / rpc.alt_scale
    pv = poly_terms(lon, lat, alt).T
    one = np.ones((lon.shape[0], 1))
    MC = np.hstack([one, pv, -C * pv])
    MR = np.hstack([one, pv, -R * pv])

    def set_coefs(JR, JC):
        coefs = np.vstack([JR[:20], 1, JR[20:], JC[:20], 1, JC[20:]]).reshape(-1)
        rpc.row_num, rpc.row_den = coefs[:20], coefs[20:40]
        rpc.col_num, rpc.col_den = coefs[40:60], coefs[60:]
        return coefs

    coefs = set_coefs(np.linalg.lstsq(MR, R, rcond=None)[0], np.linalg.lstsq(MC, C, rcond=None)[0])
    for _ in range(n_passes):
        wr = 1 / ((MR[:, :20] @ coefs[20:40]) ** 2)
        wc = 1 / ((MC[:, :20] @ coefs[60:80]) ** 2)
        JR = np.linalg.solve((MR.T * wr) @ MR + reg, (MR.T * wr) @ R)
        JC = np.linalg.solve((MC.T * wc) @ MC + reg, (MC.T * wc) @ C)
        coefs = set_coefs(JR, JC)
    return rpc


def held_out_grid(input_locs, n=(17, 17, 7), inset=0.05):
    """Dense lon/lat/alt grid strictly inside the bounding box of the fit samples (never one of them)."""
    lo, hi = input_locs.min(0), input_locs.max(0)
    axes = [np.linspace(lo[d] + inset * (hi[d] - lo[d]), hi[d] - inset * (hi[d] - lo[d]), n[d]) for d in range(3)]
    return np.stack(np.meshgrid(*axes, indexing="ij"), -1).reshape(-1, 3)
