"""
CPU tests: the oracle restatement (oracle/ba_oracle.py) against the golden vectors produced by the
unmodified reference (tests/golden/make_golden.py), and -- when the reference tree is at hand -- against
the reference itself.  Bit-exact: the restatement keeps the reference's operation order.
"""
import numpy as np
import pytest

import util
from oracle import ba_oracle
from oracle.ref_loader import load_reference, reference_available

G = util.load_ba_golden()
CASES = [str(s) for s in G["cases"]]


@pytest.mark.parametrize("name", CASES)
def test_oracle_fun_bit_exact(name):
    p = util.params_from_golden(G, name)
    pre = name + "/"
    assert np.array_equal(ba_oracle.residuals(p.params_opt.copy(), p), G[pre + "ref_fun_x0"])
    assert np.array_equal(ba_oracle.residuals(G[pre + "x1"].copy(), p), G[pre + "ref_fun_x1"])


@pytest.mark.parametrize("name", CASES)
def test_oracle_sparsity_bit_exact(name):
    p = util.params_from_golden(G, name)
    pre = name + "/"
    A = ba_oracle.jacobian_sparsity(p)
    A.sort_indices()
    assert tuple(A.shape) == tuple(G[pre + "ref_sparsity_shape"])
    assert np.array_equal(A.indptr, G[pre + "ref_sparsity_indptr"])
    assert np.array_equal(A.indices, G[pre + "ref_sparsity_indices"])


@pytest.mark.parametrize("name", ["persp_R", "persp_RT_softl1", "affine_RT_softl1", "persp_RT_huber"])
def test_oracle_solve_matches_reference_run(name):
    """scipy TRF driven exactly like ba_core.py:284-297 reproduces the reference's run bit for bit."""
    p = util.params_from_golden(G, name)
    pre = name + "/"
    v0, v1, e0, e1, nfev = ba_oracle.solve(p, util.ls_from_golden(G, name))
    assert nfev == int(G[pre + "ref_nfev"])
    assert np.array_equal(v1, G[pre + "ref_vars_ba"])
    assert np.array_equal(e0, G[pre + "ref_err_init"])
    assert np.array_equal(e1, G[pre + "ref_err_ba"])


def test_oracle_rpc_fun_bit_exact():
    R = util.load_rpc_golden()
    for corr in (["R"], ["R", "T"]):
        p = util.rpc_ba_params_from_golden(R, corr)
        tag = "rpcba/" + "".join(corr) + "/"
        assert np.array_equal(p.params_opt, R[tag + "params_opt"])
        assert np.array_equal(ba_oracle.residuals(p.params_opt.copy(), p), R[tag + "ref_fun_x0"])
        assert np.array_equal(ba_oracle.residuals(R[tag + "x1"].copy(), p), R[tag + "ref_fun_x1"])


@pytest.mark.skipif(not reference_available(), reason="reference tree only exists in the build container")
def test_oracle_against_live_reference():
    ref = load_reference()
    from sat_bundleadjust_b200 import synth
    sc = synth.make_scene(n_cam=5, n_tracks=200, p_vis=0.6, cam_model="perspective", seed=42)
    q = synth.scene_to_params(sc, ["R", "T"], n_cam_fix=1, params_cls=ref.ba_params.BundleAdjustmentParameters)
    p = synth.scene_to_params(sc, ["R", "T"], n_cam_fix=1)
    x = p.params_opt.copy()
    assert np.array_equal(ba_oracle.residuals(x.copy(), p), ref.ba_core.fun(x.copy(), q))
    A, B = ba_oracle.jacobian_sparsity(p), ref.ba_core.build_jacobian_sparsity(q).tocsr()
    assert (A != B).nnz == 0
    r = ref.ba_core.fun(x.copy(), q)
    assert np.array_equal(ba_oracle.reprojection_error(r, p.pts2d_w), ref.ba_core.compute_reprojection_error(r, q.pts2d_w))
