"""
Device triangulation of the initial 3-D points (ft_triangulate.py:18-127) against the oracle
(oracle/tri_oracle.py, pinned to cv2 / the reference's golden points in tests/test_triangulate_cpu.py) -- through the C ABI
(sba_linear_triangulation, sba_init_pts3d).  Bars: DLT points within 1e-6 m of the oracle (float64); the float32 running
mean of init_pts3d within one float32 ulp of the oracle and of the reference's golden points, differing in < 0.1 % of the
entries (a last-bit difference of the float64 triangulation can flip one float32 rounding).
"""
import numpy as np
import pytest

from oracle import tri_oracle
from sat_bundleadjust_b200 import _lib, cam_utils, ft_triangulate, synth
from test_triangulate_cpu import golden_filtered_scene

pytestmark = pytest.mark.gpu


def _within_one_ulp(got, ref, frac=1e-3):
    diff = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    return got.dtype == ref.dtype == np.float32 and np.all(diff <= np.spacing(np.abs(ref))) and np.mean(diff > 0) < frac


@pytest.mark.parametrize("model", ["perspective", "affine"])
def test_linear_triangulation_vs_oracle(built, model):
    sc = synth.make_scene(n_cam=2, n_tracks=5000, p_vis=1.0, cam_model=model, seed=4)
    P1, P2 = sc.cameras
    X = sc.pts3d_true
    u1, u2 = cam_utils.apply_projection_matrix(P1, X), cam_utils.apply_projection_matrix(P2, X)
    rng = np.random.default_rng(0)
    for sigma in (0.0, 0.5, 20.0):
        a, b = u1 + sigma * rng.standard_normal(u1.shape), u2 + sigma * rng.standard_normal(u2.shape)
        got = ft_triangulate.linear_triangulation_multiple_pts(P1, P2, a, b)
        ref = tri_oracle.linear_triangulation_multiple_pts(P1, P2, a, b)
        assert got.shape == ref.shape and np.abs(got - ref).max() < (1e-6 if model == "perspective" else 1e-4), (model, sigma)
    assert ft_triangulate.linear_triangulation_multiple_pts(P1, P2, u1[:0], u2[:0]).shape == (0, 3)


def test_init_pts3d_matches_reference_golden(built):
    C, cams, pairs, ref, nf = golden_filtered_scene()
    got = ft_triangulate.init_pts3d(C, cams, "perspective", pairs)
    assert _within_one_ulp(got[nf:], ref[nf:])
    assert _within_one_ulp(got, tri_oracle.init_pts3d(C, cams, pairs))


@pytest.mark.parametrize("model,n_cam", [("perspective", 10), ("affine", 6), ("perspective", 70)])
def test_init_pts3d_vs_oracle_pair_order_and_gaps(built, model, n_cam):
    """Pairs in arbitrary order, reversed, repeated and out of range; tracks seen by no listed pair stay zero; > 64 cameras."""
    sc = synth.make_scene(n_cam=n_cam, n_tracks=3000, p_vis=0.35 if n_cam < 20 else 0.06, cam_model=model, seed=9)
    rng = np.random.default_rng(1)
    pairs = [(i, j) for i in range(n_cam) for j in range(i + 1, n_cam) if rng.random() < (0.5 if n_cam < 20 else 0.1)]
    rng.shuffle(pairs)
    pairs = pairs + [(j, i) for i, j in pairs[:3]] + [pairs[0]] + [(0, n_cam + 2)]
    C = sc.correspondence_matrix()
    got = ft_triangulate.init_pts3d(C, sc.cameras, model, pairs)
    ref = tri_oracle.init_pts3d(C, sc.cameras, pairs)
    assert (np.abs(ref).sum(axis=1) == 0).any() and (np.abs(ref).sum(axis=1) > 0).sum() > 1000
    assert np.array_equal(got == 0, ref == 0)
    assert _within_one_ulp(got, ref, frac=3e-3)


def test_init_pts3d_edge_cases(built):
    sc = synth.make_scene(n_cam=4, n_tracks=50, p_vis=0.8, cam_model="perspective", seed=2)
    C = sc.correspondence_matrix()
    assert not ft_triangulate.init_pts3d(C, sc.cameras, "perspective", []).any()             # no pairs -> zeros
    assert ft_triangulate.init_pts3d(C[:, :0], sc.cameras, "perspective", [(0, 1)]).shape == (0, 3)
    with pytest.raises(ValueError):
        ft_triangulate.init_pts3d(C, sc.cameras, "pinhole", [(0, 1)])
    lib = _lib.load()
    assert lib.sba_init_pts3d(1, None, 4, None, None, None, 5, None, 0, None) == -1              # bad arguments are reported
