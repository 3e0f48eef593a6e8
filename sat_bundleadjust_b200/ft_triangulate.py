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


def _smallest_right_singular_vector(A, max_sweeps=30):
    """
    Right singular vector of the smallest singular value of N 4x4 matrices by one-sided (Hestenes) Jacobi rotations,
    batched over N.  The DLT matrix has columns of magnitude 1e6 x (X, Y, Z ~ 6e6 m) next to 1e13: a bidiagonalising
    SVD (LAPACK) loses ~5 mm there, the Jacobi iteration keeps the small singular pair to full relative accuracy --
    which is what OpenCV's own SVD does inside cv2.triangulatePoints.
    """
    U = np.array(A, dtype=np.float64)
    n = U.shape[0]
    V = np.broadcast_to(np.eye(4), (n, 4, 4)).copy()
    eps = np.finfo(np.float64).eps
    for _ in range(max_sweeps):
        rotated = False
        for p in range(3):
            for q in range(p + 1, 4):
                up, uq = U[:, :, p], U[:, :, q]
                alpha, beta, gamma = np.sum(up * up, axis=1), np.sum(uq * uq, axis=1), np.sum(up * uq, axis=1)
                need = np.abs(gamma) > eps * np.sqrt(alpha * beta)
                if not need.any():
                    continue
                rotated = True
                zeta = (beta - alpha) / (2.0 * np.where(need, gamma, 1.0))
                t = np.where(zeta == 0, 1.0, np.sign(zeta) / (np.abs(zeta) + np.sqrt(1.0 + zeta * zeta)))
                c = 1.0 / np.sqrt(1.0 + t * t)
                s = np.where(need, c * t, 0.0)[:, np.newaxis]
                c = np.where(need, c, 1.0)[:, np.newaxis]
                U[:, :, p], U[:, :, q] = c * up - s * uq, s * up + c * uq
                vp, vq = V[:, :, p].copy(), V[:, :, q].copy()
                V[:, :, p], V[:, :, q] = c * vp - s * vq, s * vp + c * vq
        if not rotated:
            break
    k = np.argmin(np.sum(U * U, axis=1), axis=1)
    return V[np.arange(n), :, k]


def linear_triangulation_multiple_pts(P1, P2, pts1, pts2):
    """
    Linear (DLT) triangulation of N correspondences with 3x4 matrices: the homogeneous point minimising |A X| with
    |X| = 1, A = the four equations x P[2] - P[0], y P[2] - P[1] of both views -- the definition cv2.triangulatePoints
    implements (the reference's call, ft_triangulate.py:18-34).  Agrees with cv2 to < 1e-7 m, noisy / outlier
    observations and short baselines included.
    """
    A = np.stack([pts1[:, 0:1] * P1[2] - P1[0], pts1[:, 1:2] * P1[2] - P1[1],
                  pts2[:, 0:1] * P2[2] - P2[0], pts2[:, 1:2] * P2[2] - P2[1]], axis=1)      # (N, 4, 4)
    X = _smallest_right_singular_vector(A)
    return X[:, :3] / X[:, 3:4]


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
