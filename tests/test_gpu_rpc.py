"""
GPU parity tests of the batched RPC kernels and of cam_model='rpc' residuals (run with -m gpu).
Checkers: golden vectors computed by the compiled reference C (c/rpc.c, c/disp_to_h.c) and the oracle.

Tolerances: the kernels keep the reference's algorithm and stopping rules; CUDA contracts a*b+c into FMA,
so results agree to rounding, not bit for bit:
  projection 1e-7 px; localisation 1e-9 deg (the iteration stops at 1e-9 normalised image units, i.e. ~1e-6 px);
  triangulated height 1e-3 m and lon/lat 1e-8 deg (the height iteration stops at |lambda| < 1e-5 m and amplifies
  rounding through the base-to-height ratio).
  fun for cam_model='rpc': the reference rounds the projection to float32 (ba_core.py:150, one float32 ulp at
  ~3000 px is 2.4e-4 px); we reproduce the rounding, so values agree except where the FP64 value sits within
  rounding distance of a float32 tie: <= 1 float32 ulp on < 1% of the residuals.
"""
import ctypes

import numpy as np
import pytest

import util
from oracle import ba_oracle, rpc_ctypes
from sat_bundleadjust_b200 import _lib, ba_core
from sat_bundleadjust_b200.solver import DeviceProblem, rpc_table

pytestmark = pytest.mark.gpu

R = util.load_rpc_golden()
RA, RB = util.rpc_from_array(R["rpc_a"]), util.rpc_from_array(R["rpc_b"])


def test_projection_golden(built):
    lib = _lib.load()
    lla = R["lonlatalt"]
    n = lla.shape[0]
    for rpc, key in ((RA, "ref_proj_a"), (RB, "ref_proj_b")):
        col, row = np.empty(n), np.empty(n)
        lon, lat, alt = [np.ascontiguousarray(lla[:, k]) for k in range(3)]
        _lib.check(lib.sba_rpc_projection(_lib.dptr(rpc_table(rpc)), _lib.dptr(lon), _lib.dptr(lat), _lib.dptr(alt), n,
                                          _lib.dptr(col), _lib.dptr(row)))
        assert np.abs(np.stack((col, row), axis=1) - R[key]).max() < 1e-7


def test_localization_golden(built):
    lib = _lib.load()
    cra = R["colrowalt"]
    n = cra.shape[0]
    for delta, key in ((1.0, "ref_loc_a_delta1"), (0.1, "ref_loc_a_delta01")):
        lon, lat = np.empty(n), np.empty(n)
        col, row, alt = [np.ascontiguousarray(cra[:, k]) for k in range(3)]
        _lib.check(lib.sba_rpc_localization(_lib.dptr(rpc_table(RA)), _lib.dptr(col), _lib.dptr(row), _lib.dptr(alt), n,
                                            delta, _lib.dptr(lon), _lib.dptr(lat)))
        assert np.abs(np.stack((lon, lat), axis=1) - R[key]).max() < 1e-9


def test_triangulation_golden_reference_signature(built):
    """The reference's own entry point (same symbol, same struct, same buffers) now runs on the GPU."""
    lib = _lib.load()
    n = R["kp_a"].shape[0]
    ka, kb = R["kp_a"].astype(np.float32), R["kp_b"].astype(np.float32)
    out, err = np.zeros((n, 3)), np.zeros((n, 1), dtype=np.float32)
    sa, sb = rpc_ctypes.struct_from_model(RA, 0.1), rpc_ctypes.struct_from_model(RB, 0.1)
    lib.stereo_corresp_to_lonlatalt(_lib.dptr(out), err.ctypes.data_as(_lib.c_float_p), ka.ctypes.data_as(_lib.c_float_p),
                                    kb.ctypes.data_as(_lib.c_float_p), n, ctypes.byref(sa), ctypes.byref(sb))
    ref = R["ref_tri_lonlatalt"]
    assert np.abs(out[:, :2] - ref[:, :2]).max() < 1e-8
    assert np.abs(out[:, 2] - ref[:, 2]).max() < 1e-3
    assert np.abs(err - R["ref_tri_err"]).max() < 1e-3
    # empty input is a no-op
    lib.stereo_corresp_to_lonlatalt(_lib.dptr(out), err.ctypes.data_as(_lib.c_float_p), ka.ctypes.data_as(_lib.c_float_p),
                                    kb.ctypes.data_as(_lib.c_float_p), 0, ctypes.byref(sa), ctypes.byref(sb))


def test_large_batch_round_trip(built):
    """1e6 points: localisation inverts projection (size-independent property)."""
    lib = _lib.load()
    rng = np.random.default_rng(3)
    n = 1000000
    col, row = rng.uniform(0, 3200, n), rng.uniform(0, 1350, n)
    alt = RA.alt_offset + rng.uniform(-500, 500, n)
    lon, lat, c2, r2 = np.empty(n), np.empty(n), np.empty(n), np.empty(n)
    t = _lib.dptr(rpc_table(RA))
    _lib.check(lib.sba_rpc_localization(t, _lib.dptr(col), _lib.dptr(row), _lib.dptr(alt), n, 1.0, _lib.dptr(lon), _lib.dptr(lat)))
    _lib.check(lib.sba_rpc_projection(t, _lib.dptr(lon), _lib.dptr(lat), _lib.dptr(alt), n, _lib.dptr(c2), _lib.dptr(r2)))
    assert np.abs(c2 - col).max() < 1e-4 and np.abs(r2 - row).max() < 1e-4


@pytest.mark.parametrize("corr", [["R"], ["R", "T"]])
def test_rpc_fun_golden(built, corr):
    p = util.rpc_ba_params_from_golden(R, corr)
    tag = "rpcba/" + "".join(corr) + "/"
    for x, key in ((p.params_opt.copy(), "ref_fun_x0"), (R[tag + "x1"].copy(), "ref_fun_x1")):
        r = ba_core.fun(x, p)
        ref = R[tag + key]
        diff = np.abs(r - ref)
        assert diff.max() <= 2.5e-4, diff.max()          # one float32 ulp at <= 4096 px
        assert np.mean(diff > 1e-9) < 0.01
    # un-rounded FP64 residuals agree with an FP64 evaluation of the same model
    with DeviceProblem(p, rpc_float32=False) as prob:
        r64, _ = prob.residuals(p.params_opt.copy())
    pts3d, cam = ba_oracle.unpack_variables(p.params_opt.copy(), p)
    q = ba_oracle.adjust_pts3d(pts3d[p.pts_ind], cam[p.cam_ind])
    proj = np.zeros((p.n_obs, 2))
    for j in range(p.n_cam):
        sel = p.cam_ind == j
        proj[sel] = p.cameras[j].project_ecef(q[sel])
    assert np.abs(r64 - (proj - p.pts2d).ravel()).max() < 1e-6


def test_rpc_bundle_adjustment_converges(built):
    """cam_model='rpc' end to end: analytic Jacobian vs FD of the FP64 model, and the solve lowers the cost."""
    p = util.rpc_ba_params_from_golden(R, ["R", "T"])
    from sat_bundleadjust_b200.solver import initial_vars
    x0 = initial_vars(p)
    with DeviceProblem(p, rpc_float32=False) as prob:
        Jc, Jp = prob.jacobian_blocks(x0)
        def f(x):
            return prob.residuals(x)[0]
        J = util.dense_jacobian_from_blocks(p, Jc, Jp)
        rng = np.random.default_rng(0)
        for _ in range(5):      # directional derivatives (the dense FD Jacobian would need 2n GPU calls)
            v = rng.standard_normal(x0.size)
            v[: p.n_cam * p.n_params] *= 1e-6
            h = 1e-3
            fd = (f(x0 + h * v) - f(x0 - h * v)) / (2 * h)
            assert np.abs(J @ v - fd).max() <= 1e-5 * np.abs(fd).max()
        x, r, info = prob.solve(x0, loss="linear", ftol=1e-12, xtol=1e-14, max_nfev=200)
    assert info["cost"] < 0.2 * info["cost_init"]
    assert np.sqrt(np.mean(r ** 2)) < 1.0     # observations carry 0.5 px noise


def test_rpc_model_class_and_triangulation_binding(built):
    """The rpcm-style class and the reference-style ctypes binding (s2p/triangulation.py) on top of the GPU kernels."""
    from sat_bundleadjust_b200 import triangulation
    from sat_bundleadjust_b200.rpc_model import RPCModel
    ra, rb = RPCModel(RA.to_dict()), RPCModel(RB.to_dict())
    lla = R["lonlatalt"]
    col, row = ra.projection(lla[:, 0], lla[:, 1], lla[:, 2])
    assert np.abs(np.stack((col, row), axis=1) - R["ref_proj_a"]).max() < 1e-7
    c0, r0 = ra.projection(float(lla[0, 0]), float(lla[0, 1]), float(lla[0, 2]))
    assert abs(c0 - R["ref_proj_a"][0, 0]) < 1e-7 and isinstance(c0, float)
    lon, lat = ra.localization(R["colrowalt"][:, 0], R["colrowalt"][:, 1], R["colrowalt"][:, 2])
    assert np.abs(np.stack((lon, lat), axis=1) - R["ref_loc_a_delta1"]).max() < 1e-9
    from oracle import rpc_oracle
    X = np.stack(rpc_oracle.latlon_to_ecef(lla[:, 1], lla[:, 0], lla[:, 2]), axis=1)
    assert np.abs(ra.projection_from_ecef(X) - RA.project_ecef(X)).max() < 1e-6
    out, err = triangulation.stereo_corresp_to_xyz(ra, rb, R["kp_a"], R["kp_b"])
    assert out.dtype == np.float64 and err.dtype == np.float32 and err.shape == (R["kp_a"].shape[0], 1)
    assert np.abs(out[:, :2] - R["ref_tri_lonlatalt"][:, :2]).max() < 1e-8
    assert np.abs(out[:, 2] - R["ref_tri_lonlatalt"][:, 2]).max() < 1e-3
    xyz, _ = triangulation.rpc_triangulation(ra, rb, R["kp_a"], R["kp_b"])
    ref_xyz = np.stack(rpc_oracle.latlon_to_ecef(R["ref_tri_lonlatalt"][:, 1], R["ref_tri_lonlatalt"][:, 0],
                                                 R["ref_tri_lonlatalt"][:, 2]), axis=1)
    assert np.abs(xyz - ref_xyz).max() < 2e-3


def test_init_pts3d_rpc_branch(built):
    """init_pts3d (ft_triangulate.py:57-127) for cam_model='rpc': GPU triangulation + the reference's float32 running mean."""
    from sat_bundleadjust_b200 import ft_triangulate
    from sat_bundleadjust_b200.rpc_model import RPCModel
    from oracle import rpc_oracle
    cams_o = [util.rpc_from_array(a) for a in R["rpc_cams"][:3]]
    cams = [RPCModel(c.to_dict()) for c in cams_o]
    C = R["rpcba/C"][:6]
    pairs = [(0, 1), (0, 2), (1, 2)]
    got = ft_triangulate.init_pts3d(C, cams, "rpc", pairs)
    assert got.dtype == np.float32 and got.shape == (C.shape[1], 3)
    # checker: the compiled-reference-pinned C port + the same running mean
    port = rpc_ctypes.load_port()
    avg = np.zeros((C.shape[1], 3), dtype=np.float32)
    cnt = np.zeros(C.shape[1], dtype=np.float32)
    mask = ~np.isnan(C[::2])
    for ci, cj in pairs:
        t = np.where(mask[ci] & mask[cj])[0]
        lla, _ = rpc_ctypes.triangulate(port, cams_o[ci], cams_o[cj], C[2 * ci:2 * ci + 2, t].T, C[2 * cj:2 * cj + 2, t].T, delta=0.1)
        xyz = np.stack(rpc_oracle.latlon_to_ecef(lla[:, 1], lla[:, 0], lla[:, 2]), axis=1)
        new = np.zeros((C.shape[1], 3), dtype=np.float32)
        new[t] = xyz
        cnt[t] += 1.0
        avg[t] = ((cnt[t, None] - 1.0) * avg[t] + new[t]) / cnt[t, None]
    seen = cnt > 0
    assert seen.sum() > 100
    assert np.abs(got[seen] - avg[seen]).max() <= 1.0      # float32 ulp at 6.4e6 m is 0.5 m
    assert np.array_equal(got[~seen], avg[~seen])


def test_batched_projection_and_localization(built):
    """sba_rpc_projection_batch / sba_rpc_localization_batch: all cameras in one launch, bit-identical to the per-camera calls."""
    from sat_bundleadjust_b200 import rpc_model
    models = [rpc_model.RPCModel(util.rpc_from_array(a).to_dict()) for a in R["rpc_cams"][:3]]
    lla = R["lonlatalt"]
    col, row = rpc_model.projection_batch(models, lla[:, 0], lla[:, 1], lla[:, 2])              # shared points
    for j, m in enumerate(models):
        c1, r1 = m.projection(lla[:, 0], lla[:, 1], lla[:, 2])
        assert np.array_equal(col[j], c1) and np.array_equal(row[j], r1)
    lon, lat = rpc_model.localization_batch(models, col, row, np.broadcast_to(lla[:, 2], col.shape))   # per-camera points
    for j, m in enumerate(models):
        l1, a1 = m.localization(col[j], row[j], lla[:, 2])
        assert np.array_equal(lon[j], l1) and np.array_equal(lat[j], a1)
        assert np.abs(lon[j] - lla[:, 0]).max() < 1e-7 and np.abs(lat[j] - lla[:, 1]).max() < 1e-7
    with pytest.raises(ValueError):
        rpc_model.projection_batch(models, np.zeros((2, 5)), np.zeros((2, 5)), np.zeros((2, 5)))
