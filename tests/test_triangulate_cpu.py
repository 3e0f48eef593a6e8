"""CPU check of the batched DLT used by ft_triangulate.init_pts3d for matrix cameras."""
import numpy as np
import pytest


def test_linear_triangulation_matches_cv2():   # CPU-only helper, runs in the not-gpu suite too
    cv2 = pytest.importorskip("cv2")
    from sat_bundleadjust_b200 import ft_triangulate, synth
    sc = synth.make_scene(n_cam=2, n_tracks=200, p_vis=1.0, cam_model="perspective", seed=4)
    P1, P2 = sc.cameras
    from sat_bundleadjust_b200 import cam_utils
    X = sc.pts3d_true
    u1, u2 = cam_utils.apply_projection_matrix(P1, X), cam_utils.apply_projection_matrix(P2, X)
    got = ft_triangulate.linear_triangulation_multiple_pts(P1, P2, u1, u2)
    ref = cv2.triangulatePoints(P1, P2, u1.T, u2.T)
    ref = (ref[:3] / ref[3]).T
    assert np.abs(got - X).max() < 1e-6 and np.abs(got - ref).max() < 1e-6
    rng = np.random.default_rng(0)
    u1, u2 = u1 + 0.5 * rng.standard_normal(u1.shape), u2 + 0.5 * rng.standard_normal(u2.shape)
    got = ft_triangulate.linear_triangulation_multiple_pts(P1, P2, u1, u2)
    ref = cv2.triangulatePoints(P1, P2, u1.T, u2.T)
    ref = (ref[:3] / ref[3]).T
    assert np.abs(got - ref).max() < 1e-4
