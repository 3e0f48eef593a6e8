"""
Two-view RPC triangulation of keypoint matches on the GPU.

Drop-in for the reference's only native binding, bundle_adjust/s2p/triangulation.py (`RPCStruct`,
`stereo_corresp_to_xyz`), which loads lib/disp_to_h.so and calls `stereo_corresp_to_lonlatalt`
(c/disp_to_h.c:40-65).  libsba_b200.so exports that symbol with the same signature and the same
`struct rpc` layout (c/rpc.h:14-32), so the binding below is the reference's binding with another library
path; the loop over matches runs as one kernel, one match per thread.
"""
import ctypes
from ctypes import POINTER, byref, c_double, c_float, c_int

import numpy as np
from numpy.ctypeslib import ndpointer

from . import _lib


class RPCStruct(ctypes.Structure):
    """ctypes mirror of `struct rpc` (c/rpc.h:14-32)."""
    _fields_ = [("numx", c_double * 20), ("denx", c_double * 20), ("numy", c_double * 20), ("deny", c_double * 20),
                ("scale", c_double * 3), ("offset", c_double * 3),
                ("inumx", c_double * 20), ("idenx", c_double * 20), ("inumy", c_double * 20), ("ideny", c_double * 20),
                ("iscale", c_double * 3), ("ioffset", c_double * 3),
                ("dmval", c_double * 4), ("imval", c_double * 4), ("delta", c_double)]

    def __init__(self, rpc, delta=1.0):
        self.offset[:] = [rpc.col_offset, rpc.row_offset, rpc.alt_offset]
        self.ioffset[:] = [rpc.lon_offset, rpc.lat_offset, rpc.alt_offset]
        self.scale[:] = [rpc.col_scale, rpc.row_scale, rpc.alt_scale]
        self.iscale[:] = [rpc.lon_scale, rpc.lat_scale, rpc.alt_scale]
        self.inumx[:], self.idenx[:] = list(rpc.col_num), list(rpc.col_den)
        self.inumy[:], self.ideny[:] = list(rpc.row_num), list(rpc.row_den)
        if hasattr(rpc, "lat_num"):      # direct (localisation) model present
            self.numx[:], self.denx[:] = list(rpc.lon_num), list(rpc.lon_den)
            self.numy[:], self.deny[:] = list(rpc.lat_num), list(rpc.lat_den)
        else:                            # absent -> NaN -> iterative localisation (c/rpc.c:417-425)
            nan = [float("nan")] * 20
            self.numx[:], self.denx[:], self.numy[:], self.deny[:] = nan, nan, nan, nan
        self.delta = delta


def stereo_corresp_to_xyz(rpc1, rpc2, pts1, pts2, out_crs=None):
    """
    Point cloud (lon, lat, alt) from N matches between two RPC images, and the triangulation error.
    Same contract as the reference: keypoints are cast to float32, outputs are float64 (N,3) and float32 (N,1).
    `out_crs` other than geographic coordinates is not supported here (CRS conversion is out of scope).
    """
    if out_crs is not None:
        raise NotImplementedError("CRS conversion is outside the hot path; convert the returned lon/lat/alt")
    lib = _lib.load()
    s1, s2 = RPCStruct(rpc1, delta=0.1), RPCStruct(rpc2, delta=0.1)
    n = pts1.shape[0]
    # The package's own binding goes through the status-returning entry point: a missing device, an out-of-memory condition or
    # any CUDA error raises SbaError instead of handing zero-filled points to the caller.  (The void symbol
    # `stereo_corresp_to_lonlatalt` with the reference's exact signature stays exported for the reference's unmodified ctypes
    # stub; it fills its outputs with NaN on failure.)
    lonlatalt = np.zeros((n, 3), dtype="float64")
    err = np.zeros((n, 1), dtype="float32")
    k1, k2 = np.ascontiguousarray(pts1, dtype="float32"), np.ascontiguousarray(pts2, dtype="float32")
    _lib.check(lib.sba_stereo_corresp_to_lonlatalt(lonlatalt.ctypes.data_as(_lib.c_double_p), err.ctypes.data_as(_lib.c_float_p),
                                                   k1.ctypes.data_as(_lib.c_float_p), k2.ctypes.data_as(_lib.c_float_p), n,
                                                   byref(s1), byref(s2)))
    return lonlatalt, err


def rpc_triangulation(rpc1, rpc2, pts1, pts2):
    """ECEF points from matches (feature_tracks/ft_triangulate.py:37-54)."""
    from . import geo_utils
    lonlatalt, err = stereo_corresp_to_xyz(rpc1, rpc2, pts1, pts2)
    x, y, z = geo_utils.latlon_to_ecef_custom(lonlatalt[:, 1], lonlatalt[:, 0], lonlatalt[:, 2])
    return np.vstack((x, y, z)).T, err
