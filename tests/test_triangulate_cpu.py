"""CPU checks of the batched DLT used by ft_triangulate.init_pts3d for matrix cameras (the reference calls cv2.triangulatePoints)."""
import os

import numpy as np
import pytest

from sat_bundleadjust_b200 import cam_utils, ft_triangulate, synth


def _cv2_triangulate(cv2, P1, P2, a, b):
    X = cv2.triangulatePoints(P1, P2, a.T, b.T)
    return (X[:3] / X[3]).T


def test_linear_triangulation_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    sc = synth.make_scene(n_cam=2, n_tracks=200, p_vis=1.0, cam_model="perspective", seed=4)
    P1, P2 = sc.cameras
    X = sc.pts3d_true
    u1, u2 = cam_utils.apply_projection_matrix(P1, X), cam_utils.apply_projection_matrix(P2, X)
    got = ft_triangulate.linear_triangulation_multiple_pts(P1, P2, u1, u2)
    assert np.abs(got - X).max() < 1e-6 and np.abs(got - _cv2_triangulate(cv2, P1, P2, u1, u2)).max() < 1e-6
    rng = np.random.default_rng(0)
    for sigma in (0.5, 20.0):        # pixel noise, then gross outliers: still the same minimiser as OpenCV's
        a, b = u1 + sigma * rng.standard_normal(u1.shape), u2 + sigma * rng.standard_normal(u2.shape)
        got = ft_triangulate.linear_triangulation_multiple_pts(P1, P2, a, b)
        assert np.abs(got - _cv2_triangulate(cv2, P1, P2, a, b)).max() < 1e-6, sigma


def test_init_pts3d_matrix_cameras_matches_reference_golden():
    """init_pts3d (ft_triangulate.py:57-127) on the filtered scene of the outlier golden: float32 points of the reference."""
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "outliers_golden.npz"))
    idx, nf = G["scene/new/pts_prev_indices"], int(G["scene/new/n_pts_fix"])
    C = G["scene/C"][:, idx].copy()
    C[G["scene/new/C_nan"]] = np.nan
    pairs = [tuple(int(v) for v in p) for p in G["scene/pairs"]]
    got = ft_triangulate.init_pts3d(C, list(G["scene/cameras"]), "perspective", pairs)
    ref = G["scene/new/pts3d"]
    assert got.dtype == ref.dtype == np.float32
    # rows < nf were overwritten with the frozen points afterwards (ba_outliers.py:91-92)
    diff = np.abs(got[nf:].astype(np.float64) - ref[nf:].astype(np.float64))
    assert np.all(diff <= np.spacing(np.abs(ref[nf:]))) and np.mean(diff > 0) < 1e-3
