"""
Initial 3d points of the feature tracks by pairwise triangulation -- mirror of the part of the reference's
bundle_adjust/feature_tracks/ft_triangulate.py that sits next to the hot path (`init_pts3d`, :57-127, and
`rpc_triangulation`, :37-54).  For cam_model == "rpc" every pair goes through ONE batched GPU call
(triangulation.stereo_corresp_to_xyz -> k_rpc_triangulate) instead of the serial C loop of c/disp_to_h.c:50-64;
matrix cameras use a batched linear triangulation (the reference calls cv2.triangulatePoints, :18-34).
The running float32 mean over the pairs is kept exactly as the reference computes it (:77-81, :115-117): its float32
quantisation of the points (~0.5 m at ECEF magnitude) is part of the reference's input to bundle adjustment.
"""
import numpy as np

from .triangulation import rpc_triangulation


def linear_triangulation_multiple_pts(P1, P2, pts1, pts2):
    """
    Linear triangulation of N correspondences with 3x4 matrices (the reference calls cv2.triangulatePoints,
    ft_triangulate.py:18-34).  The four DLT equations are solved for the finite point (X, Y, Z, 1) as a batched 4x3
    least-squares problem: at ECEF magnitudes the homogeneous SVD loses ~5 mm to the 1 : 6e6 dynamic range of the
    null vector, this form agrees with cv2 to < 1e-6 m on noisy input.
    """
    A = np.stack([pts1[:, 0:1] * P1[2] - P1[0], pts1[:, 1:2] * P1[2] - P1[1],
                  pts2[:, 0:1] * P2[2] - P2[0], pts2[:, 1:2] * P2[2] - P2[1]], axis=1)      # (N, 4, 4)
    return (np.linalg.pinv(A[:, :, :3]) @ (-A[:, :, 3:4]))[:, :, 0]


def init_pts3d(C, cameras, cam_model, pairs_to_triangulate, verbose=False):
    """Average of all pairwise triangulations of every track; returns (N,3) float32 like the reference."""
    n_pts, n_cam = C.shape[1], C.shape[0] // 2
    avg = np.zeros((n_pts, 3), dtype=np.float32)
    cnt = np.zeros(n_pts, dtype=np.float32)
    mask = ~np.isnan(C[::2])
    for (ci, cj) in pairs_to_triangulate:
        if not (ci < n_cam and cj < n_cam):
            continue
        t = np.where(mask[ci] & mask[cj])[0]
        if t.shape[0] == 0:
            continue
        oi, oj = C[2 * ci: 2 * ci + 2, t].T, C[2 * cj: 2 * cj + 2, t].T
        if cam_model in ("affine", "perspective"):
            new = linear_triangulation_multiple_pts(cameras[ci], cameras[cj], oi, oj)
        else:
            new, _ = rpc_triangulation(cameras[ci], cameras[cj], oi, oj)
        new32 = np.zeros((n_pts, 3), dtype=np.float32)
        new32[t] = new
        cnt[t] += 1.0
        avg[t] = ((cnt[t, np.newaxis] - 1.0) * avg[t] + new32[t]) / cnt[t, np.newaxis]
    return avg
