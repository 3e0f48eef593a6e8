"""
RPC refit (ba_rpcfit.weighted_lsq).  CPU: the oracle restatement against golden vectors of the UNMODIFIED reference
function.  GPU: the batched kernel against the same golden vectors.

Tolerances for the GPU fit.  The ridge (h^2 = 1e-6 on normalised variables) bounds the condition number of the
re-weighted normal matrices at ~1e9, so FP64 solutions agree to ~1e-7 relative; the reference inverts with LAPACK LU
(np.linalg.inv), the kernel eliminates with partial pivoting.  We require
  * normalisation constants: 1e-12 relative (same min/max arithmetic)
  * coefficients: |delta| <= 1e-6 * max|coef| of the same polynomial           (north_star: 1e-6 relative)
  * projections of the fitted model on the samples: 1e-6 px from the reference fit's projections
"""
import numpy as np
import pytest

import util
from oracle import rpcfit_oracle

F = np.load(util.GOLDEN + "/rpcfit_golden.npz")
NCASES = int(F["n_cases"])


def _oracle_table(rpc):
    return np.concatenate([[rpc.row_offset, rpc.col_offset, rpc.lat_offset, rpc.lon_offset, rpc.alt_offset,
                            rpc.row_scale, rpc.col_scale, rpc.lat_scale, rpc.lon_scale, rpc.alt_scale],
                           rpc.row_num, rpc.row_den, rpc.col_num, rpc.col_den])


@pytest.mark.parametrize("k", range(NCASES))
def test_oracle_weighted_lsq_bit_exact(k):
    pre = "case%d/" % k
    fit = rpcfit_oracle.weighted_lsq(F[pre + "target"], F[pre + "input_locs"])
    assert np.array_equal(_oracle_table(fit), F[pre + "ref_rpc"])


def test_oracle_sampling_matches_golden():
    R = util.load_rpc_golden()
    pre = "case0/"
    rpc = util.rpc_from_array(R[str(F[pre + "src"])])
    crop = {"col0": 0.0, "row0": 0.0, "width": 3199.0, "height": 1349.0}
    target, locs, _ = rpcfit_oracle.rt_corrected_samples(F[pre + "Rt"], rpc, crop, margin=int(F[pre + "margin"]))
    assert np.array_equal(target, F[pre + "target"]) and np.array_equal(locs, F[pre + "input_locs"])


@pytest.mark.gpu
def test_gpu_weighted_lsq_batch_vs_reference_golden(built):
    from sat_bundleadjust_b200 import ba_rpcfit
    targets = np.stack([F["case%d/target" % k] for k in range(NCASES)])
    locs = np.stack([F["case%d/input_locs" % k] for k in range(NCASES)])
    models, iters, rmse = ba_rpcfit.weighted_lsq_batch(targets, locs)
    for k, m in enumerate(models):
        ref = F["case%d/ref_rpc" % k]
        got = m.table()
        assert np.allclose(got[:10], ref[:10], rtol=1e-12, atol=0)
        for a in range(4):
            r, g = ref[10 + 20 * a: 30 + 20 * a], got[10 + 20 * a: 30 + 20 * a]
            assert np.abs(g - r).max() <= 1e-6 * np.abs(r).max(), (k, a, np.abs(g - r).max(), np.abs(r).max())
        x = locs[k]
        ref_model = util.rpc_from_array(ref)
        pr = np.stack(ref_model.projection(x[:, 0], x[:, 1], x[:, 2]), axis=1)
        pg = np.stack(m.projection(x[:, 0], x[:, 1], x[:, 2]), axis=1)
        assert np.abs(pr - pg).max() < 1e-6
        err = ba_rpcfit.check_errors(m, x, targets[k])
        assert np.abs(err - F["case%d/ref_err" % k]).max() < 1e-6
        assert 1 <= iters[k] <= 20 and rmse[k] < 0.01
    # single-camera entry point, same arguments as the reference
    one = ba_rpcfit.weighted_lsq(targets[3], locs[3])
    assert np.array_equal(one.table(), models[3].table())


@pytest.mark.gpu
def test_gpu_fit_Rt_corrected_rpc_driver(built):
    """The whole refit driver (grid -> localisation -> corrective mapping -> projection -> fit -> coverage test)."""
    from sat_bundleadjust_b200 import ba_rpcfit
    from sat_bundleadjust_b200.rpc_model import RPCModel
    R = util.load_rpc_golden()
    src = util.rpc_from_array(R["rpc_a"])
    rpc = RPCModel(src.to_dict())
    Rt = F["case0/Rt"]
    crop = {"col0": 0.0, "row0": 0.0, "width": 3199.0, "height": 1349.0}
    pts = np.array([[1771000.0, -5693000.0, 1201000.0]])
    from oracle import rpc_oracle
    g = R["lonlatalt"][:50]
    pts = np.stack(rpc_oracle.latlon_to_ecef(g[:, 1], g[:, 0], np.full(50, src.alt_offset)), axis=1)
    fit, err, margin = ba_rpcfit.fit_Rt_corrected_rpc(Rt, None, rpc, crop, pts)
    assert margin in (10, 20, 40, 80, 160, 320, 640, 1280) and err.shape == (1000,) and err.max() < 0.01
    # the refitted RPC reproduces the corrected mapping at fresh points
    rng = np.random.default_rng(0)
    lla = np.stack([src.lon_offset + rng.uniform(-.01, .01, 200), src.lat_offset + rng.uniform(-.004, .004, 200),
                    src.alt_offset + rng.uniform(-500, 500, 200)], axis=1)
    X = np.stack(rpc_oracle.latlon_to_ecef(lla[:, 1], lla[:, 0], lla[:, 2]), axis=1)
    from oracle import ba_oracle
    want = src.project_ecef(ba_oracle.adjust_pts3d(X, np.tile(Rt, (200, 1))))
    got = np.stack(fit.projection(lla[:, 0], lla[:, 1], lla[:, 2]), axis=1)
    inside = (want[:, 0] > 0) & (want[:, 0] < 3199) & (want[:, 1] > 0) & (want[:, 1] < 1349)
    assert inside.sum() > 20 and np.abs(got - want)[inside].max() < 0.02
