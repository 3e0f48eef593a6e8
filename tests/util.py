"""Shared helpers for the test-suite (test infrastructure)."""
import os

import numpy as np

from oracle import rpc_oracle
from sat_bundleadjust_b200.ba_params import BundleAdjustmentParameters

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def load_ba_golden():
    return np.load(os.path.join(GOLDEN, "ba_golden.npz"))


def load_rpc_golden():
    return np.load(os.path.join(GOLDEN, "rpc_golden.npz"))


def params_from_golden(G, name, params_cls=BundleAdjustmentParameters):
    pre = name + "/"
    ncf, npf, refw = G[pre + "opts"]
    d = {"correction_params": [str(s) for s in G[pre + "correction_params"]], "n_cam_fix": int(ncf),
         "n_pts_fix": int(npf), "ref_cam_weight": float(refw), "reduce": False, "verbose": False}
    return params_cls(G[pre + "C"], G[pre + "pts3d_init"], list(G[pre + "cameras_init"]), str(G[pre + "cam_model"]),
                      [(0, 1)], list(G[pre + "camera_centers"]), d)


def ls_from_golden(G, name):
    pre = name + "/"
    cfg = {}
    for k, v in zip(G[pre + "ls_keys"], G[pre + "ls_vals"]):
        k, v = str(k), str(v)
        cfg[k] = v if k == "loss" else (int(float(v)) if k in ("max_iter", "verbose") else float(v))
    return cfg


def rpc_from_array(a):
    r = rpc_oracle.RPCModel()
    (r.row_offset, r.col_offset, r.lat_offset, r.lon_offset, r.alt_offset,
     r.row_scale, r.col_scale, r.lat_scale, r.lon_scale, r.alt_scale) = [float(v) for v in a[:10]]
    r.row_num, r.row_den, r.col_num, r.col_den = [list(a[10 + 20 * i: 30 + 20 * i]) for i in range(4)]
    return r


def rpc_ba_params_from_golden(G, corr, params_cls=BundleAdjustmentParameters):
    cams = [rpc_from_array(a) for a in G["rpc_cams"]]
    d = {"correction_params": list(corr), "reduce": False, "verbose": False}
    return params_cls(G["rpcba/C"], G["rpcba/pts3d_init"], cams, "rpc", [(0, 1)], list(G["rpcba/camera_centers"]), d)


def dense_jacobian_from_blocks(p, Jc, Jp):
    """Assemble the dense (2K x n) Jacobian from per-observation blocks (small problems only)."""
    K, c = p.pts_ind.size, p.n_params
    n = p.n_cam * c + 3 * p.n_pts
    J = np.zeros((2 * K, n))
    for k in range(K):
        j, i = int(p.cam_ind[k]), int(p.pts_ind[k])
        J[2 * k:2 * k + 2, j * c:(j + 1) * c] = Jc[k]
        J[2 * k:2 * k + 2, p.n_cam * c + 3 * i: p.n_cam * c + 3 * i + 3] = Jp[k]
    return J
