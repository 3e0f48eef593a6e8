"""
TEST INFRASTRUCTURE ONLY.  ctypes access to the two CPU checkers of the native RPC path:
  * oracle/librpc_oracle.so  -- our C restatement (oracle/rpc_oracle.c), symbols rpco_*
  * oracle/_ref/disp_to_h.so -- the reference's own c/rpc.c + c/disp_to_h.c compiled in place
                                (present only if `make -C oracle ref` ran where /root/reference exists)
The struct is the ABI of the reference's c/rpc.h:14-32.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
D20 = ctypes.c_double * 20


class RPCStruct(ctypes.Structure):
    _fields_ = [("numx", D20), ("denx", D20), ("numy", D20), ("deny", D20),
                ("scale", ctypes.c_double * 3), ("offset", ctypes.c_double * 3),
                ("inumx", D20), ("idenx", D20), ("inumy", D20), ("ideny", D20),
                ("iscale", ctypes.c_double * 3), ("ioffset", ctypes.c_double * 3),
                ("dmval", ctypes.c_double * 4), ("imval", ctypes.c_double * 4), ("delta", ctypes.c_double)]


def struct_from_model(rpc, delta=1.0):
    """Same field mapping as bundle_adjust/s2p/triangulation.py:38-78 (direct model absent -> NaN)."""
    s = RPCStruct()
    s.offset[:] = [rpc.col_offset, rpc.row_offset, rpc.alt_offset]
    s.scale[:] = [rpc.col_scale, rpc.row_scale, rpc.alt_scale]
    s.ioffset[:] = [rpc.lon_offset, rpc.lat_offset, rpc.alt_offset]
    s.iscale[:] = [rpc.lon_scale, rpc.lat_scale, rpc.alt_scale]
    s.inumx[:], s.idenx[:] = list(rpc.col_num), list(rpc.col_den)
    s.inumy[:], s.ideny[:] = list(rpc.row_num), list(rpc.row_den)
    nan = [float("nan")] * 20
    s.numx[:], s.denx[:], s.numy[:], s.deny[:] = nan, nan, nan, nan
    s.delta = delta
    return s


def build(target="all"):
    subprocess.run(["make", "-C", HERE, target], check=True, capture_output=True)


def load_port():
    path = os.path.join(HERE, "librpc_oracle.so")
    if not os.path.exists(path):
        build("oracle")
    lib = ctypes.CDLL(path)
    P = ctypes.POINTER(RPCStruct)
    dp, fp = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_float)
    lib.rpco_stereo_corresp_to_lonlatalt.argtypes = [dp, fp, fp, fp, ctypes.c_int, P, P]
    lib.rpco_project_batch.argtypes = [dp, P, dp, ctypes.c_int]
    lib.rpco_localize_batch.argtypes = [dp, P, dp, ctypes.c_int]
    return lib


def load_ref():
    """The compiled reference, or None when oracle/_ref/ was not built (no /root/reference at hand)."""
    path = os.path.join(HERE, "_ref", "disp_to_h.so")
    if not os.path.exists(path):
        return None
    lib = ctypes.CDLL(path)
    P = ctypes.POINTER(RPCStruct)
    dp, fp = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_float)
    lib.stereo_corresp_to_lonlatalt.argtypes = [dp, fp, fp, fp, ctypes.c_int, P, P]
    lib.eval_rpci.argtypes = [dp, P, ctypes.c_double, ctypes.c_double, ctypes.c_double]
    lib.eval_rpc.argtypes = [dp, P, ctypes.c_double, ctypes.c_double, ctypes.c_double]
    lib.rpc_height.argtypes = [P, P] + [ctypes.c_double] * 4 + [dp]
    lib.rpc_height.restype = ctypes.c_double
    return lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def triangulate(lib, rpc_a, rpc_b, kp_a, kp_b, delta=0.1, ref=False):
    """stereo_corresp_to_lonlatalt through either library; kp_* are (n,2), cast to float32 like the reference."""
    n = kp_a.shape[0]
    ka = np.ascontiguousarray(kp_a, dtype=np.float32)
    kb = np.ascontiguousarray(kp_b, dtype=np.float32)
    out = np.zeros((n, 3), dtype=np.float64)
    err = np.zeros((n, 1), dtype=np.float32)
    sa, sb = struct_from_model(rpc_a, delta), struct_from_model(rpc_b, delta)
    f = lib.stereo_corresp_to_lonlatalt if ref else lib.rpco_stereo_corresp_to_lonlatalt
    f(_dp(out), _fp(err), _fp(ka), _fp(kb), n, ctypes.byref(sa), ctypes.byref(sb))
    return out, err


def ref_project(lib, rpc, lonlatalt):
    s = struct_from_model(rpc)
    out = np.zeros((lonlatalt.shape[0], 2))
    tmp = (ctypes.c_double * 2)()
    for i, (a, b, c) in enumerate(lonlatalt):
        lib.eval_rpci(tmp, ctypes.byref(s), a, b, c)
        out[i] = tmp[0], tmp[1]
    return out


def ref_localize(lib, rpc, colrowalt, delta=1.0):
    s = struct_from_model(rpc, delta)
    out = np.zeros((colrowalt.shape[0], 2))
    tmp = (ctypes.c_double * 2)()
    for i, (a, b, c) in enumerate(colrowalt):
        lib.eval_rpc(tmp, ctypes.byref(s), a, b, c)
        out[i] = tmp[0], tmp[1]
    return out


def port_project(lib, rpc, lonlatalt):
    s = struct_from_model(rpc)
    x = np.ascontiguousarray(lonlatalt, dtype=np.float64)
    out = np.zeros((x.shape[0], 2))
    lib.rpco_project_batch(_dp(out), ctypes.byref(s), _dp(x), x.shape[0])
    return out


def port_localize(lib, rpc, colrowalt, delta=1.0):
    s = struct_from_model(rpc, delta)
    x = np.ascontiguousarray(colrowalt, dtype=np.float64)
    out = np.zeros((x.shape[0], 2))
    lib.rpco_localize_batch(_dp(out), ctypes.byref(s), _dp(x), x.shape[0])
    return out
