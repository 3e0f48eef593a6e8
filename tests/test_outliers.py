"""
Outlier removal between the two bundle-adjustment passes against golden vectors produced by the
unmodified reference (tests/golden/make_outliers_golden.py).  Bar: thresholds, elbow values and the set of removed
observations bit-exact; the re-triangulated float32 points within one float32 ulp (the reference triangulates with
cv2.triangulatePoints, we with a one-sided Jacobi DLT on the device that agrees with it to < 1e-7 m).
"""
import os

import numpy as np
import pytest

from sat_bundleadjust_b200 import ba_outliers
from sat_bundleadjust_b200.ba_params import BundleAdjustmentParameters

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "outliers_golden.npz"))


def test_percentile_positions_reproduce_numpy():
    rng = np.random.default_rng(0)
    for n in list(range(1, 40)) + [100, 101, 4999, 50000]:
        v = np.sort(rng.normal(0, 3, n) ** 2)
        for q in (80, 50, 99.5, 0, 100):
            lo, hi, gamma = ba_outliers._percentile_positions(n, q)
            assert ba_outliers._lerp(v[lo], v[hi], gamma) == np.percentile(v, q), (n, q)


def test_filter_pairs_matches_the_loop_definition():
    rng = np.random.default_rng(1)
    C = rng.normal(size=(12, 200))
    C[np.repeat(rng.random((6, 200)) < 0.6, 2, axis=0)] = np.nan
    pairs = [(0, 1), (2, 5), (4, 3), (1, 4)]
    mask = ~np.isnan(C[::2])
    want = [i for i in range(200) if any(a < b and mask[a, i] and mask[b, i] for a, b in pairs)]
    assert np.array_equal(ba_outliers.filter_C_using_pairs_to_triangulate(C, pairs), want)


def test_filter_pairs_matches_the_live_reference():
    """feature_tracks/ft_utils.py:38-62 of the unmodified reference, where the reference tree is present."""
    from oracle.ref_loader import load_reference, reference_available
    if not reference_available():
        pytest.skip("reference tree only exists in the build container")
    import sys
    load_reference()
    old = sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    try:
        from bundle_adjust.feature_tracks import ft_utils
    finally:
        sys.dont_write_bytecode = old
    rng = np.random.default_rng(3)
    for n_cam, n_tr, p_vis in ((6, 400, 0.5), (12, 300, 0.2), (3, 50, 0.9)):
        C = rng.normal(size=(2 * n_cam, n_tr))
        C[np.repeat(rng.random((n_cam, n_tr)) > p_vis, 2, axis=0)] = np.nan
        pairs = [(int(a), int(b)) for a, b in rng.integers(0, n_cam, size=(7, 2))]
        want = ft_utils.filter_C_using_pairs_to_triangulate(C, pairs)
        assert np.array_equal(ba_outliers.filter_C_using_pairs_to_triangulate(C, pairs), want)


def _scene_params():
    d = {"correction_params": ["R", "T"], "n_cam_fix": 0, "n_pts_fix": 30, "ref_cam_weight": 1.0, "reduce": False, "verbose": False}
    pairs = [tuple(int(v) for v in p) for p in G["scene/pairs"]]
    return BundleAdjustmentParameters(G["scene/C"], G["scene/pts3d"], list(G["scene/cameras"]), "perspective", pairs,
                                      list(G["scene/centers"]), d)


@pytest.mark.gpu
def test_reset_ba_params_from_golden_masks(built):
    """reset_ba_params_after_outlier_removal (ba_outliers.py:61-109) on the correspondence matrix the reference filtered (re-triangulation on the device)."""
    p = _scene_params()
    C_new = p.C.copy()
    C_new[G["scene/auto/C_new_nan"]] = np.nan
    new_p = ba_outliers.reset_ba_params_after_outlier_removal(C_new, p, verbose=False)
    assert np.array_equal(np.isnan(new_p.C), G["scene/new/C_nan"])
    assert np.array_equal(new_p.pts_ind, G["scene/new/pts_ind"]) and np.array_equal(new_p.cam_ind, G["scene/new/cam_ind"])
    assert np.array_equal(new_p.pts2d, G["scene/new/pts2d"])
    assert int(new_p.n_pts_fix) == int(G["scene/new/n_pts_fix"])
    assert np.array_equal(new_p.pts_prev_indices, G["scene/new/pts_prev_indices"])
    ref_pts = G["scene/new/pts3d"]
    diff = np.abs(new_p.pts3d.astype(np.float64) - ref_pts.astype(np.float64))
    assert new_p.pts3d.dtype == ref_pts.dtype and np.all(diff <= np.spacing(np.abs(ref_pts))) and np.mean(diff > 0) < 1e-3
    n_cam_vars = new_p.n_cam * new_p.n_params
    assert np.array_equal(new_p.params_opt[:n_cam_vars], G["scene/new/params_opt"][:n_cam_vars])


@pytest.mark.gpu
def test_elbow_values_bit_exact(built):
    for k in range(int(G["elbow/n"])):
        val, ok = ba_outliers.get_elbow_value(G["elbow/%d/err" % k])
        assert val == float(G["elbow/%d/value" % k]), k
        assert ok == bool(G["elbow/%d/success" % k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("tag,kw", [("auto", {}), ("predef", {"predef_thr": 2.345}), ("minthr", {"min_thr": 6.0})])
def test_compute_obs_to_remove_bit_exact(built, tag, kw):
    p = _scene_params()
    C_new, cam_thr, n_det = ba_outliers.compute_obs_to_remove(G["scene/err"], p, **kw)
    assert np.array_equal(np.array(cam_thr, dtype=np.float64), G["scene/%s/cam_thr" % tag])
    assert int(n_det) == int(G["scene/%s/n_detected" % tag])
    assert np.array_equal(np.isnan(C_new), G["scene/%s/C_new_nan" % tag])
    keep = ~np.isnan(C_new)
    assert np.array_equal(C_new[keep], p.C[keep])


@pytest.mark.gpu
def test_rm_outliers_rebuilds_the_problem_like_the_reference(built):
    p = _scene_params()
    new_p = ba_outliers.rm_outliers(G["scene/err"], p)
    assert np.array_equal(np.isnan(new_p.C), G["scene/new/C_nan"])
    assert np.array_equal(new_p.pts_ind, G["scene/new/pts_ind"]) and np.array_equal(new_p.cam_ind, G["scene/new/cam_ind"])
    assert np.array_equal(new_p.pts2d, G["scene/new/pts2d"])
    assert int(new_p.n_pts_fix) == int(G["scene/new/n_pts_fix"])
    assert np.array_equal(new_p.pts_prev_indices, G["scene/new/pts_prev_indices"])
    ref_pts = G["scene/new/pts3d"]
    assert new_p.pts3d.dtype == ref_pts.dtype and new_p.pts3d.shape == ref_pts.shape
    ulp = np.spacing(np.abs(ref_pts).astype(np.float32)).astype(np.float64)
    diff = np.abs(new_p.pts3d.astype(np.float64) - ref_pts.astype(np.float64))
    assert np.all(diff <= ulp) and np.mean(diff > 0) < 1e-3
    n_cam_vars = new_p.n_cam * new_p.n_params
    assert np.array_equal(new_p.params_opt[:n_cam_vars], G["scene/new/params_opt"][:n_cam_vars])


@pytest.mark.gpu
def test_sort_at_full_size_properties(built):
    """5e5 observations, 10 cameras: the elbow equals the numpy definition evaluated per camera on the host."""
    rng = np.random.default_rng(5)
    K, M = 500000, 10
    cam = rng.integers(0, M, K).astype(np.int32)
    err = np.hypot(rng.normal(0, 0.5, K), rng.normal(0, 0.5, K)) + np.where(rng.random(K) < 0.02, np.abs(rng.normal(0, 20, K)), 0.0)
    counts, elbow, perc, vmax = ba_outliers._elbow_stats(err, cam, M, 20)
    for c in range(M):
        v = np.sort(err[cam == c])
        n = v.size
        assert counts[c] == n and vmax[c] == v[-1] and perc[c] == np.percentile(v, 80)
        coord = np.vstack((np.arange(n), v)).T
        lv = coord[-1] - coord[0]
        lvn = lv / np.sqrt(np.sum(lv ** 2))
        vf = coord - coord[0]
        sp = np.sum(vf * np.tile(lvn, (n, 1)), axis=1)
        dist = np.sqrt(np.sum((vf - np.outer(sp, lvn)) ** 2, axis=1))
        assert elbow[c] == v[np.argmax(dist)], c
