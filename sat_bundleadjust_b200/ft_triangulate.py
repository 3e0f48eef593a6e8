"""
Initial 3d points of the feature tracks by pairwise triangulation -- mirror of the part of the reference's
bundle_adjust/feature_tracks/ft_triangulate.py that sits next to the hot path (`linear_triangulation_multiple_pts`, :18-34,
`rpc_triangulation`, :37-54, and `init_pts3d`, :57-127).

The reference loops over the triangulation pairs on the host, one cv2.triangulatePoints / one serial C loop
(c/disp_to_h.c:50-64) per pair, and keeps a float32 running mean per track.  Here ALL pairs x ALL tracks go through one
kernel launch (csrc/sba_rpc.cu k_init_pts3d, entry point sba_init_pts3d): one thread per track walks the pair list in
order, triangulates on the spot (4x4 Jacobi DLT for matrix cameras, the two-view RPC iteration + geodetic -> ECEF for
RPCs) and keeps the same float32 running mean -- its quantisation of the points (~0.5 m at ECEF magnitude) is part of
the reference's input to bundle adjustment.  No CPU fallback: without libsba_b200.so / a GPU these raise.
"""
import ctypes

import numpy as np

from . import _lib
from .triangulation import RPCStruct, rpc_triangulation  # noqa: F401  (rpc_triangulation re-exported like in the reference)

_MODEL = {"affine": 0, "perspective": 1, "rpc": 2}


def linear_triangulation_multiple_pts(P1, P2, pts1, pts2):
    """
    Linear (DLT) triangulation of N correspondences with 3x4 matrices (ft_triangulate.py:18-34; the reference calls
    cv2.triangulatePoints).  Returns (N,3) float64.  Agrees with cv2 to < 1e-6 m, noisy observations included.
    """
    lib = _lib.load()
    P1, P2 = _lib.f64(P1), _lib.f64(P2)
    a, b = _lib.f64(pts1), _lib.f64(pts2)
    if P1.shape != (3, 4) or P2.shape != (3, 4) or a.shape != b.shape or a.ndim != 2 or a.shape[1] != 2:
        raise ValueError("expected two 3x4 matrices and two (N,2) arrays")
    out = np.zeros((a.shape[0], 3), dtype=np.float64)
    _lib.check(lib.sba_linear_triangulation(_lib.dptr(P1), _lib.dptr(P2), _lib.dptr(a), _lib.dptr(b), a.shape[0], _lib.dptr(out)))
    return out


def tracks_from_C(C):
    """Correspondence matrix (2M x N, NaN = not seen) -> CSR tracks (track_ptr int64, cam_idx int32 ascending, pts2d (K,2))."""
    mask = ~np.isnan(C[::2])
    pts_ind, cam_ind = np.nonzero(mask.T)                     # track-major, cameras ascending inside a track
    track_ptr = np.zeros(C.shape[1] + 1, dtype=np.int64)
    np.cumsum(np.bincount(pts_ind, minlength=C.shape[1]), out=track_ptr[1:])
    pts2d = np.stack((C[2 * cam_ind, pts_ind], C[2 * cam_ind + 1, pts_ind]), axis=1).astype(np.float64)
    return track_ptr, cam_ind.astype(np.int32), np.ascontiguousarray(pts2d)


def init_pts3d(C, cameras, cam_model, pairs_to_triangulate, verbose=False):
    """Average of all pairwise triangulations of every track; returns (N,3) float32 like the reference (:57-127)."""
    if cam_model not in _MODEL:
        raise ValueError("cam_model must be 'affine', 'perspective' or 'rpc'")
    lib = _lib.load()
    n_pts, n_cam = C.shape[1], C.shape[0] // 2
    if verbose:
        print("Computing {} points 3d from feature tracks...".format(n_pts), flush=True)
    track_ptr, cam_idx, pts2d = tracks_from_C(np.asarray(C))
    if cam_model == "rpc":                                    # `struct rpc` per camera, the delta the reference's binding sets
        cams = np.concatenate([np.frombuffer(bytes(RPCStruct(c, delta=0.1)), dtype=np.float64) for c in cameras[:n_cam]])
    else:
        cams = np.concatenate([_lib.f64(P).reshape(12) for P in cameras[:n_cam]])
    pairs = np.ascontiguousarray(np.asarray(list(pairs_to_triangulate), dtype=np.int64).reshape(-1, 2).astype(np.int32))
    out = np.zeros((n_pts, 3), dtype=np.float32)
    i32p, i64p = ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int64)
    _lib.check(lib.sba_init_pts3d(_MODEL[cam_model], _lib.dptr(cams), n_cam, track_ptr.ctypes.data_as(i64p), cam_idx.ctypes.data_as(i32p),
                                  _lib.dptr(pts2d), n_pts, pairs.ctypes.data_as(i32p), pairs.shape[0], out.ctypes.data_as(_lib.c_float_p)))
    if verbose:
        print("done!", flush=True)
    return out
