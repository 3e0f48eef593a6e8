"""
sat_bundleadjust_b200 -- B200-native implementation of the bundle-adjustment hot path of
centreborelli/sat-bundleadjust (residual -> analytic Jacobian -> Schur/Cholesky trust-region solve,
plus batched RPC projection / localisation / triangulation), behind the reference's Python API.

    from sat_bundleadjust_b200 import ba_core, ba_params
    p = ba_params.BundleAdjustmentParameters(C, pts3d, cameras, cam_model, pairs, centers, d)
    vars_init, vars_ba, err_init, err_ba, nfev = ba_core.run_ba_optimization(p, ls_params)

All arithmetic runs in hand-written sm_100a CUDA kernels (csrc/) reached through the C ABI of
include/sba_b200.h; there is no CPU fallback.
"""
__version__ = "0.1.0"
