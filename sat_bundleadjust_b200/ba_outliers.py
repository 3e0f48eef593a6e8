"""
Suppression of outlier feature-track observations between the two bundle-adjustment passes -- mirror of the
reference's bundle_adjust/ba_outliers.py, same function names, arguments and return values.

The per-camera work (sorting every camera's reprojection errors and locating the elbow of the sorted curve,
ba_outliers.py:14-58, and thresholding every observation, :140-146) runs on the GPU for all cameras at once
(csrc/sba_outliers.cu: one stable radix sort by (camera, error) + one CTA per camera).  What is left on the host is the
O(n_cam) scalar logic, kept in numpy so that it rounds exactly like the reference: np.percentile's interpolation,
max(elbow, min_thr), np.round(thr, 2).  Thresholds and the set of removed observations are bit-identical to the
reference's (tests/test_outliers.py).
"""
import ctypes

import numpy as np

from . import _lib


def _percentile_positions(n, q):
    """Order statistics and weight np.percentile(., q) (method 'linear') interpolates between, for a sample of size n."""
    quantile = np.true_divide(q, 100)
    virtual = (n - 1) * quantile                        # numpy: lambda n, quantiles: (n - 1) * quantiles
    prev = np.floor(virtual)
    nxt = prev + 1
    if virtual >= n - 1:
        prev, nxt = -1.0, -1.0                          # numpy takes the last element
    gamma = np.float64(virtual - prev)
    lo, hi = int(prev) % n, int(nxt) % n
    return lo, hi, gamma


def _lerp(a, b, t):
    """numpy's _lerp (lib/_function_base_impl.py): a + (b - a) t, or b - (b - a)(1 - t) when t >= 0.5."""
    a, b, t = np.float64(a), np.float64(b), np.float64(t)
    d = b - a
    return b - d * (1 - t) if t >= 0.5 else a + d * t


def _elbow_stats(err, cam_ind, n_cam, max_outliers_percent):
    """GPU part: per camera (count, elbow value, percentile, maximum)."""
    err = np.ascontiguousarray(err, dtype=np.float64)
    cam = np.ascontiguousarray(cam_ind, dtype=np.int32)
    counts = np.bincount(cam, minlength=n_cam).astype(np.int64)
    if np.any(counts == 0):
        # the reference fails on a camera without observations (IndexError in get_elbow_value)
        raise IndexError("camera(s) %s have no observations" % np.where(counts == 0)[0].tolist())
    pos = [_percentile_positions(int(n), 100 - max_outliers_percent) for n in counts]
    q_lo = np.array([p[0] for p in pos], dtype=np.int64)
    q_hi = np.array([p[1] for p in pos], dtype=np.int64)
    stats = np.empty((n_cam, 5), dtype=np.float64)
    got = np.empty(n_cam, dtype=np.int64)
    lib = _lib.load()
    _lib.check(lib.sba_outlier_elbow(_lib.dptr(err), cam.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(err.size),
                                     ctypes.c_int32(n_cam), q_lo.ctypes.data_as(ctypes.c_void_p),
                                     q_hi.ctypes.data_as(ctypes.c_void_p), _lib.dptr(stats),
                                     got.ctypes.data_as(ctypes.c_void_p)))
    assert np.array_equal(got, counts)
    perc = np.array([_lerp(stats[c, 1], stats[c, 2], pos[c][2]) for c in range(n_cam)])
    perc[np.isnan(stats[:, 3])] = np.nan                # a NaN in the sample makes np.percentile return NaN
    return counts, stats[:, 0].copy(), perc, stats[:, 3].copy()


def get_elbow_value(err, max_outliers_percent=20, verbose=False):
    """
    Elbow value of a function expected to follow an L shape (ba_outliers.py:14-58): the sorted sample furthest from
    the segment joining the smallest and the largest one.  Returns (elbow_value, success); success is False when the
    elbow falls below the (100 - max_outliers_percent)-th percentile.
    """
    err = np.asarray(err, dtype=np.float64).ravel()
    _, elbow, perc, _ = _elbow_stats(err, np.zeros(err.size, dtype=np.int32), 1, max_outliers_percent)
    elbow_value = float(elbow[0])
    success = False if (elbow_value < perc[0]) else True
    return elbow_value, success


def compute_obs_to_remove(err, p, predef_thr=None, min_thr=1.0):
    """
    Per-camera reprojection-error thresholds and the correspondence matrix without the observations above them
    (ba_outliers.py:112-153).  Returns (C_new, cam_thr, n_detected_outliers).
    """
    err = np.ascontiguousarray(err, dtype=np.float64)
    cam = np.ascontiguousarray(p.cam_ind, dtype=np.int32)
    if predef_thr is None:
        _, elbow, perc, vmax = _elbow_stats(err, cam, p.n_cam, 20)
        cam_thr = []
        for c in range(p.n_cam):
            success = False if (elbow[c] < perc[c]) else True
            thr = max(elbow[c], min_thr) if success else vmax[c]
            cam_thr.append(np.round(thr, 2))
    else:
        cam_thr = [np.round(float(predef_thr), 2) for _ in range(p.n_cam)]
    remove = np.empty(err.size, dtype=np.uint8)
    lib = _lib.load()
    thr = np.ascontiguousarray(cam_thr, dtype=np.float64)
    _lib.check(lib.sba_outlier_mark(_lib.dptr(err), cam.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(err.size),
                                    ctypes.c_int32(p.n_cam), _lib.dptr(thr), remove.ctypes.data_as(ctypes.c_void_p)))
    rm = remove.astype(bool)
    C_new = p.C.copy()
    if rm.any():
        ci, pi = np.asarray(p.cam_ind)[rm], np.asarray(p.pts_ind)[rm]
        C_new[ci * 2, pi] = np.nan
        C_new[ci * 2 + 1, pi] = np.nan
    n_detected_outliers = np.sum(~np.isnan(p.C[::2]).ravel()) - np.sum(~np.isnan(C_new[::2]).ravel())
    return C_new, cam_thr, n_detected_outliers


def filter_C_using_pairs_to_triangulate(C, pairs_to_triangulate):
    """Columns of C holding at least one pair of pairs_to_triangulate (feature_tracks/ft_utils.py:38-62), vectorised."""
    mask = ~np.isnan(C[::2])
    keep = np.zeros(C.shape[1], dtype=bool)
    n_cam = mask.shape[0]
    for (i, j) in set(pairs_to_triangulate):
        if i < j and 0 <= i < n_cam and 0 <= j < n_cam:
            keep |= mask[i] & mask[j]
    return np.where(keep)[0]


def reset_ba_params_after_outlier_removal(C_new, p, verbose=True):
    """
    Bundle-adjustment parameters coherent with the filtered correspondence matrix (ba_outliers.py:61-109): tracks left
    with fewer than two observations or without a pair suitable for triangulation are dropped, the surviving tracks are
    re-triangulated from the observations they kept (frozen points keep their coordinates) and a fresh
    BundleAdjustmentParameters object is built with the options of `p`.
    """
    from .ba_params import BundleAdjustmentParameters
    from .ft_triangulate import init_pts3d

    # an observation occupies two rows of C, hence ">= 4 finite entries" == ">= 2 observations"
    long_enough = np.flatnonzero(np.count_nonzero(~np.isnan(C_new), axis=0) >= 4)
    C_kept = C_new[:, long_enough]
    triangulable = filter_C_using_pairs_to_triangulate(C_kept, p.pairs_to_triangulate)
    C_kept = C_kept[:, triangulable]
    survivors = long_enough[triangulable]                      # indices into the tracks of p
    frozen = survivors[survivors < p.n_pts_fix]                # frozen tracks are the first n_pts_fix of p
    pts3d = init_pts3d(C_kept, p.cameras, p.cam_model, p.pairs_to_triangulate, verbose=verbose)
    if frozen.size > 0:
        pts3d[: frozen.size, :] = p.pts3d[frozen, :]
    options = {"n_cam_fix": p.n_cam_fix, "n_pts_fix": np.sum(1 * (survivors < p.n_pts_fix)), "reduce": False, "verbose": verbose,
               "correction_params": p.cam_params_to_optimize, "ref_cam_weight": p.ref_cam_weight}
    new_p = BundleAdjustmentParameters(C_kept, pts3d, p.cameras, p.cam_model, p.pairs_to_triangulate, p.camera_centers, options)
    new_p.pts_prev_indices = p.pts_prev_indices[survivors]
    return new_p


def rm_outliers(err, p, predef_thr=None, min_thr=1.0, verbose=False):
    """
    Remove outlier observations according to their reprojection error (ba_outliers.py:156-186).  Returns `p` itself
    when nothing is removed, else the rebuilt parameters.
    """
    C_new, cam_thr, n_removed = compute_obs_to_remove(err, p, predef_thr=predef_thr, min_thr=min_thr)
    new_p = p if n_removed == 0 else reset_ba_params_after_outlier_removal(C_new, p, verbose=verbose)
    if verbose:
        n_obs, n_tracks = len(p.cam_ind), p.C.shape[1]
        n_tracks_removed = n_tracks - new_p.C.shape[1]
        print("Reprojection error threshold per camera: {} px".format(cam_thr))
        print("Deleted {} observations ({:.2f}%) and {} tracks ({:.2f}%)".format(
            n_removed, 100.0 * n_removed / n_obs, n_tracks_removed, 100.0 * n_tracks_removed / n_tracks))
    return new_p
