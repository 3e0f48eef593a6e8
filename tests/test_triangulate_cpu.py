"""
CPU pins of the triangulation ORACLE (oracle/tri_oracle.py) that the GPU tests of ft_triangulate check the device kernels
against: cv2.triangulatePoints (the reference's call, ft_triangulate.py:18-34) when OpenCV is importable, and the float32
points the unmodified reference produced (tests/golden/outliers_golden.npz).
"""
import os

import numpy as np
import pytest

from oracle import tri_oracle
from sat_bundleadjust_b200 import cam_utils, ft_triangulate, synth


def _cv2_triangulate(cv2, P1, P2, a, b):
    X = cv2.triangulatePoints(P1, P2, a.T, b.T)
    return (X[:3] / X[3]).T


def test_oracle_linear_triangulation_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    sc = synth.make_scene(n_cam=2, n_tracks=200, p_vis=1.0, cam_model="perspective", seed=4)
    P1, P2 = sc.cameras
    X = sc.pts3d_true
    u1, u2 = cam_utils.apply_projection_matrix(P1, X), cam_utils.apply_projection_matrix(P2, X)
    got = tri_oracle.linear_triangulation_multiple_pts(P1, P2, u1, u2)
    assert np.abs(got - X).max() < 1e-6 and np.abs(got - _cv2_triangulate(cv2, P1, P2, u1, u2)).max() < 1e-6
    rng = np.random.default_rng(0)
    for sigma in (0.5, 20.0):        # pixel noise, then gross outliers: still the same minimiser as OpenCV's
        a, b = u1 + sigma * rng.standard_normal(u1.shape), u2 + sigma * rng.standard_normal(u2.shape)
        got = tri_oracle.linear_triangulation_multiple_pts(P1, P2, a, b)
        assert np.abs(got - _cv2_triangulate(cv2, P1, P2, a, b)).max() < 1e-6, sigma


def test_oracle_linear_triangulation_recovers_exact_points():
    """No OpenCV needed: noiseless projections triangulate back to the points (perspective and affine matrices)."""
    for model in ("perspective", "affine"):
        sc = synth.make_scene(n_cam=2, n_tracks=300, p_vis=1.0, cam_model=model, seed=5)
        P1, P2 = sc.cameras
        X = sc.pts3d_true
        u1, u2 = cam_utils.apply_projection_matrix(P1, X), cam_utils.apply_projection_matrix(P2, X)
        got = tri_oracle.linear_triangulation_multiple_pts(P1, P2, u1, u2)
        assert np.abs(got - X).max() < (1e-5 if model == "perspective" else 1e-3), model


def golden_filtered_scene():
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "outliers_golden.npz"))
    idx, nf = G["scene/new/pts_prev_indices"], int(G["scene/new/n_pts_fix"])
    C = G["scene/C"][:, idx].copy()
    C[G["scene/new/C_nan"]] = np.nan
    pairs = [tuple(int(v) for v in p) for p in G["scene/pairs"]]
    return C, list(G["scene/cameras"]), pairs, G["scene/new/pts3d"], nf


def test_oracle_init_pts3d_matches_reference_golden():
    """init_pts3d (ft_triangulate.py:57-127) on the filtered scene of the outlier golden: float32 points of the reference."""
    C, cams, pairs, ref, nf = golden_filtered_scene()
    got = tri_oracle.init_pts3d(C, cams, pairs)
    assert got.dtype == ref.dtype == np.float32
    # rows < nf were overwritten with the frozen points afterwards (ba_outliers.py:91-92)
    diff = np.abs(got[nf:].astype(np.float64) - ref[nf:].astype(np.float64))
    assert np.all(diff <= np.spacing(np.abs(ref[nf:]))) and np.mean(diff > 0) < 1e-3


def test_tracks_from_C():
    """Host-side conversion of the correspondence matrix to the CSR tracks the kernel reads."""
    C = np.full((6, 4), np.nan)
    C[0:2, 0] = (1, 2); C[4:6, 0] = (3, 4)          # track 0: cameras 0, 2
    C[2:4, 2] = (5, 6); C[0:2, 2] = (7, 8); C[4:6, 2] = (9, 10)   # track 2: cameras 0, 1, 2; tracks 1 and 3 empty
    ptr, cam, xy = ft_triangulate.tracks_from_C(C)
    assert ptr.tolist() == [0, 2, 2, 5, 5] and cam.tolist() == [0, 2, 0, 1, 2]
    assert xy.tolist() == [[1, 2], [3, 4], [7, 8], [5, 6], [9, 10]]
