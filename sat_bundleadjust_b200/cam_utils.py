"""
Camera-model helpers on the bundle-adjustment hot path (host side, per camera).

Mirrors the subset of the reference's bundle_adjust/cam_utils.py that BA needs:
  decompose_perspective_camera  (cam_utils.py:45-75)
  compose_perspective_camera    (cam_utils.py:78-89)
  decompose_affine_camera       (cam_utils.py:92-127)
  compose_affine_camera         (cam_utils.py:130-143)
  apply_projection_matrix       (cam_utils.py:201-214)
  apply_rpc_projection          (cam_utils.py:217-231)
  generate_point_mesh           (cam_utils.py:280-306)
The RPC -> projection-matrix fitting helpers (cam_utils.py:146-198, 234-356) are
one-off per-camera set-up that needs the `ad` autodiff package; out of scope.
"""
import numpy as np


def decompose_perspective_camera(P):
    """P = K [R | vecT] with diag(K) > 0; returns K, R, vecT, optical centre oC."""
    from scipy import linalg

    M = P[:, :3]
    K, R = linalg.rq(M)
    sgn = np.diag(np.sign(np.diag(K)))
    R = sgn.dot(R)
    K = K.dot(sgn)
    oC = -(np.linalg.inv(M).dot(P[:, 3]))
    vecT = (R @ -oC[:, np.newaxis]).T[0]
    return K, R, vecT, oC


def compose_perspective_camera(K, R, oC):
    """P = K R [I | -oC]."""
    return K @ R @ np.hstack((np.eye(3), -np.asarray(oC).reshape((3, 1))))


def decompose_affine_camera(P):
    """Affine camera (Hartley & Zisserman 6.3.3): returns K (2x2), R (3x3), vecT (2x1)."""
    M = P[:2, :3]
    G = M @ M.T
    fy = np.sqrt(G[1, 1])
    s = G[1, 0] / fy
    fx = np.sqrt(G[0, 0] - s ** 2)
    K = np.array([[fx, s], [0, fy]])
    Kinv = np.linalg.inv(K)
    R2 = Kinv @ M
    r3 = np.cross(R2[0], R2[1])
    R = np.vstack((R2[0], R2[1], r3))
    vecT = Kinv @ P[:2, 3].reshape(2, 1)
    return K, R, vecT


def compose_affine_camera(K, R, vecT):
    """Affine 3x4 matrix from K (2x2), R (3x3; first two rows used) and vecT (2,)."""
    E = np.zeros((3, 4))
    E[:2, :3] = R[:2]
    E[:2, 3] = np.asarray(vecT).ravel()
    E[2, 3] = 1.0
    I = np.zeros((3, 3))
    I[:2, :2] = K
    I[2, 2] = 1.0
    return I @ E


def apply_projection_matrix(P, pts3d):
    """Project Nx3 ECEF points with a 3x4 matrix -> Nx2 (col, row)."""
    h = P @ np.hstack((pts3d, np.ones((pts3d.shape[0], 1)))).T
    return (h[:2, :] / h[-1, :]).T


def apply_rpc_projection(rpc, pts3d):
    """
    Project Nx3 ECEF points with an RPC model -> Nx2 (col, row).
    `rpc` is any object with the rpcm-style `.projection(lon, lat, alt)`; with
    sat_bundleadjust_b200.rpc_model.RPCModel the geodetic conversion and the rational
    polynomial both run in one batched sm_100a kernel (csrc/sba_rpc.cu).
    """
    if hasattr(rpc, "projection_from_ecef"):
        return rpc.projection_from_ecef(pts3d)
    from . import geo_utils

    lat, lon, alt = geo_utils.ecef_to_latlon_custom(pts3d[:, 0], pts3d[:, 1], pts3d[:, 2])
    col, row = rpc.projection(lon, lat, alt)
    return np.vstack((col, row)).T


def generate_point_mesh(col_range, row_range, alt_range):
    """(min, max, n) triplets -> flattened col/row/alt lists, alt-major then row then col."""
    c, r, a = [np.linspace(v[0], v[1], v[2]) for v in (col_range, row_range, alt_range)]
    A, R, C = np.meshgrid(a, r, c, indexing="ij")
    return C.reshape(-1), R.reshape(-1), A.reshape(-1)
