"""
TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the triangulation step that initialises the 3-D points,
bundle_adjust/feature_tracks/ft_triangulate.py:18-127.  Never imported by the product path.

  linear_triangulation_multiple_pts   :18-34   the reference calls cv2.triangulatePoints (OpenCV, not vendored): the unit
                                               homogeneous X minimising |A X| for the four DLT equations of two views.
                                               Restated with a one-sided Jacobi SVD (what OpenCV's own SVD is); pinned
                                               against cv2 when it is importable (tests/test_triangulate_cpu.py) and
                                               against the reference's float32 points of tests/golden/outliers_golden.npz.
  init_pts3d                          :57-127  pair loop + float32 running mean, restated verbatim in numpy; the pair
                                               triangulation is a callable so that the RPC branch can use the compiled
                                               reference port (oracle/rpc_ctypes.py).
"""
import numpy as np


def smallest_right_singular_vector(A, max_sweeps=30):
    """Right singular vector of the smallest singular value of N 4x4 matrices, one-sided (Hestenes) Jacobi, batched over N."""
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
    """ft_triangulate.py:18-34."""
    A = np.stack([pts1[:, 0:1] * P1[2] - P1[0], pts1[:, 1:2] * P1[2] - P1[1],
                  pts2[:, 0:1] * P2[2] - P2[0], pts2[:, 1:2] * P2[2] - P2[1]], axis=1)      # (N, 4, 4)
    X = smallest_right_singular_vector(A)
    return X[:, :3] / X[:, 3:4]


def init_pts3d(C, cameras, pairs_to_triangulate, triangulate=linear_triangulation_multiple_pts):
    """ft_triangulate.py:57-127: float32 running mean over the pairs, in list order.  triangulate(cam_i, cam_j, obs_i, obs_j) -> (n,3)."""
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
        new = triangulate(cameras[ci], cameras[cj], C[2 * ci: 2 * ci + 2, t].T, C[2 * cj: 2 * cj + 2, t].T)
        new32 = np.zeros((n_pts, 3), dtype=np.float32)
        new32[t] = new
        cnt[t] += 1.0
        avg[t] = ((cnt[t, np.newaxis] - 1.0) * avg[t] + new32[t]) / cnt[t, np.newaxis]
    return avg
