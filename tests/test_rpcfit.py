"""
RPC refit (ba_rpcfit.weighted_lsq).  CPU: the oracle restatement against golden vectors of the UNMODIFIED reference
function.  GPU: the batched kernel against the same golden vectors.

Tolerances for the GPU fit.  Coefficient-level parity with the reference is ILL-POSED, not merely hard: the first,
unregularised normal matrix M^T M has condition number ~4e16 (numerically singular) and the ridge-regularised ones
~6e11, and the number of re-weighting passes is decided by a 1e-2 px threshold on an RMSE that the first solve
determines.  Replacing np.linalg.inv(A) @ b by np.linalg.solve(A, b) inside the reference's own function already
changes the pass count (1 instead of 3), the coefficients by 3e-4 and the projections on the samples by 1e-3 px
(test_reference_fit_is_ill_conditioned below documents this).  What IS well-posed, and what we require:
  * normalisation constants: 1e-12 relative (same min/max arithmetic)
  * the fitted function on the samples: a fit error against the targets (check_errors) no worse than 3x the
    reference's or 2.5e-2 px (max) / 1e-2 px (rms), i.e. the stopping threshold itself; tightening `tol` gives the
    converged fit (test below)
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


def test_reference_fit_is_ill_conditioned():
    """np.linalg.solve instead of np.linalg.inv in the reference algorithm: coefficients move by > 1e-5 relative."""
    t, x = F["case0/target"], F["case0/input_locs"]
    ref = util.rpc_from_array(F["case0/ref_rpc"])
    lon = (x[:, 0] - ref.lon_offset) / ref.lon_scale
    lat = (x[:, 1] - ref.lat_offset) / ref.lat_scale
    alt = (x[:, 2] - ref.alt_offset) / ref.alt_scale
    R = ((t[:, 1] - ref.row_offset) / ref.row_scale)[:, None]
    pv = rpcfit_oracle.poly_terms(lon, lat, alt).T
    MR = np.hstack([np.ones((lon.size, 1)), pv, -R * pv])
    A = MR.T @ MR
    assert np.linalg.cond(A) > 1e14
    a = (np.linalg.inv(A) @ (MR.T @ R)).ravel()
    b = np.linalg.solve(A, MR.T @ R).ravel()
    assert np.abs(a - b).max() > 1e-6 * np.abs(a).max()


def test_reference_fit_noise_on_a_held_out_grid():
    """
    How much "the same RPC" can mean.  The reference's fit against the same algorithm, same number of passes, with accurate
    linear solves (oracle.weighted_lsq_accurate): the fitted FUNCTIONS differ by 4e-4 .. 5e-2 px on a dense held-out grid.  That
    is the reference's own numerical noise (np.linalg.inv at condition numbers 4e16 / 6e11), so no other implementation can be
    closer to it than this; the GPU test below uses it as its bar.
    """
    worst = 0.0
    for k in range(NCASES):
        t, x = F["case%d/target" % k], F["case%d/input_locs" % k]
        ref, n_it, _ = rpcfit_oracle.weighted_lsq(t, x, return_iters=True)
        acc = rpcfit_oracle.weighted_lsq_accurate(t, x, n_it)
        g = rpcfit_oracle.held_out_grid(x)
        d = np.abs(np.stack(acc.projection(g[:, 0], g[:, 1], g[:, 2]), 1) - np.stack(ref.projection(g[:, 0], g[:, 1], g[:, 2]), 1)).max()
        worst = max(worst, d)
    assert 1e-3 < worst < 0.1, worst


@pytest.mark.gpu
def test_gpu_fit_function_parity_on_a_held_out_grid(built):
    """
    GPU-fitted RPC against the reference-fitted RPC as FUNCTIONS (what "same RPC out" means downstream), on a dense held-out
    lon/lat/alt grid inside the sample hull:
      * default stopping rule: within the reference's own numerical noise (see the test above): measured 4e-4 .. 4.9e-2 px, case
        by case the same figures as the accurate CPU restatement shows against the reference; bar 6e-2 px;
      * pass count forced equal to the reference's and accurate arithmetic on both sides (GPU vs oracle.weighted_lsq_accurate):
        the two well-conditioned implementations of the same algorithm agree to 1e-3 px (measured); bar 2e-3 px.
    """
    from sat_bundleadjust_b200 import ba_rpcfit
    targets = np.stack([F["case%d/target" % k] for k in range(NCASES)])
    locs = np.stack([F["case%d/input_locs" % k] for k in range(NCASES)])
    models, iters, _ = ba_rpcfit.weighted_lsq_batch(targets, locs)
    d_ref, d_acc = [], []
    for k in range(NCASES):
        g = rpcfit_oracle.held_out_grid(locs[k])
        ref = util.rpc_from_array(F["case%d/ref_rpc" % k])
        want = np.stack(ref.projection(g[:, 0], g[:, 1], g[:, 2]), 1)
        got = np.stack(models[k].projection(g[:, 0], g[:, 1], g[:, 2]), 1)
        d_ref.append(np.abs(got - want).max())
        _, n_it, _ = rpcfit_oracle.weighted_lsq(targets[k], locs[k], return_iters=True)
        forced, it_f, _ = ba_rpcfit.weighted_lsq_batch(targets[k][None], locs[k][None], tol=0.0, max_iter=n_it)
        assert it_f[0] == n_it
        acc = rpcfit_oracle.weighted_lsq_accurate(targets[k], locs[k], n_it)
        a = np.stack(acc.projection(g[:, 0], g[:, 1], g[:, 2]), 1)
        f = np.stack(forced[0].projection(g[:, 0], g[:, 1], g[:, 2]), 1)
        d_acc.append(np.abs(f - a).max())
    print("held-out grid, GPU vs reference fit (px):", ["%.1e" % v for v in d_ref])
    print("held-out grid, GPU vs accurate restatement, equal passes (px):", ["%.1e" % v for v in d_acc])
    assert max(d_ref) < 6e-2
    assert max(d_acc) < 2e-3


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
        for a in (0, 2):        # leading (affine) terms of the numerators are well determined; the denominators are not
            r, g = ref[10 + 20 * a: 14 + 20 * a], got[10 + 20 * a: 14 + 20 * a]
            assert np.abs(g - r).max() <= 1e-3 * np.abs(ref[10 + 20 * a: 30 + 20 * a]).max()
        x = locs[k]
        err = ba_rpcfit.check_errors(m, x, targets[k])
        ref_err = F["case%d/ref_err" % k]
        # the pass count hangs on |dRMSE| < 1e-2 px: with an accurate first solve the loop may stop after one pass where the
        # reference's inaccurate inverse makes it run three, leaving up to ~2e-2 px at the corners of the widest samplings
        assert err.max() <= max(3 * ref_err.max(), 2.5e-2) and np.sqrt(np.mean(err ** 2)) <= max(3 * np.sqrt(np.mean(ref_err ** 2)), 1e-2)
        assert 1 <= iters[k] <= 20 and rmse[k] < 0.01
    # a tighter stopping threshold reaches the reference's fit quality everywhere
    tight, _, _ = ba_rpcfit.weighted_lsq_batch(targets, locs, tol=1e-5)
    for k, m in enumerate(tight):
        err = ba_rpcfit.check_errors(m, locs[k], targets[k])
        assert err.max() <= max(3 * F["case%d/ref_err" % k].max(), 2e-3)
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
