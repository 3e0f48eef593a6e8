"""
Golden vectors for the outlier-removal step, produced by the UNMODIFIED reference
(bundle_adjust/ba_outliers.py, imported through oracle/ref_loader.py).  Run in the build container only:

    python tests/golden/make_outliers_golden.py        ->  tests/golden/outliers_golden.npz

  elbow/<k>/...   get_elbow_value on hand-made and random samples (L-shaped, flat, ties, tiny, constant)
  scene/...       compute_obs_to_remove and rm_outliers on a synthetic perspective scene whose reprojection errors
                  come from the reference `fun` at the initial point (restricted pairs_to_triangulate, fixed points)
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_loader import load_reference  # noqa: E402
from sat_bundleadjust_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def import_ref_outliers():
    ref = load_reference()
    # modules the reference's ft_triangulate pulls in and that cannot load here (no rasterio; the C library is
    # looked up under /root/reference/lib, which does not exist and is read-only): the matrix-camera branch used
    # below only needs cv2
    sys.modules.setdefault("geojson", types.ModuleType("geojson"))
    tri = types.ModuleType("bundle_adjust.s2p.triangulation")
    tri.stereo_corresp_to_xyz = None
    s2p = types.ModuleType("bundle_adjust.s2p")
    s2p.__path__ = []
    s2p.triangulation = tri
    sys.modules["bundle_adjust.s2p"] = s2p
    sys.modules["bundle_adjust.s2p.triangulation"] = tri
    old = sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    try:
        from bundle_adjust import ba_outliers
        from bundle_adjust.feature_tracks import ft_triangulate, ft_utils  # noqa: F401
    finally:
        sys.dont_write_bytecode = old
    return ref, ba_outliers


def elbow_samples():
    rng = np.random.default_rng(7)
    out = []
    out.append(np.abs(np.concatenate([rng.normal(0, 0.5, 900), rng.normal(0, 20, 100)])))          # L shape
    out.append(rng.uniform(0, 3, 500))                                                               # flat: no elbow
    out.append(np.abs(rng.normal(0, 1.0, 2000)) + np.where(rng.random(2000) < 0.02, 15.0, 0.0))
    out.append(np.array([0.7]))                                                                      # one sample (0/0 chord)
    out.append(np.array([0.3, 5.0]))
    out.append(np.array([2.0, 0.1, 0.1]))
    out.append(np.full(40, 1.25))                                                                    # constant
    out.append(np.round(np.abs(rng.normal(0, 2.0, 300)), 1))                                         # many ties
    out.append(np.concatenate([np.zeros(50), [0.0, 1e-300, 1e300]]))                                 # extreme magnitudes
    out.append(np.hypot(rng.normal(0, 0.5, 50000), rng.normal(0, 0.5, 50000)) + np.where(rng.random(50000) < 0.02, np.abs(rng.normal(0, 20, 50000)), 0))
    return out


def main():
    ref, ba_outliers = import_ref_outliers()
    out = {}
    samples = elbow_samples()
    out["elbow/n"] = np.array(len(samples))
    for k, e in enumerate(samples):
        with np.errstate(all="ignore"):
            val, ok = ba_outliers.get_elbow_value(e.copy())
        out["elbow/%d/err" % k] = e
        out["elbow/%d/value" % k] = np.array(val)
        out["elbow/%d/success" % k] = np.array(bool(ok))
    # scene: 6 cameras, fixed points, only some pairs suitable for triangulation
    sc = synth.make_scene(n_cam=6, n_tracks=1500, p_vis=0.5, cam_model="perspective", seed=321, outlier_frac=0.05)
    d = {"correction_params": ["R", "T"], "n_cam_fix": 0, "n_pts_fix": 30, "ref_cam_weight": 1.0, "reduce": False, "verbose": False}
    pairs = [(0, 1), (0, 2), (1, 2), (2, 3), (3, 4), (4, 5), (1, 4)]
    C = sc.correspondence_matrix()
    p = ref.ba_params.BundleAdjustmentParameters(C, sc.pts3d_init, list(sc.cameras_init), sc.cam_model, pairs,
                                                 list(sc.camera_centers), d)
    r = ref.ba_core.fun(p.params_opt.copy(), p)
    err = ref.ba_core.compute_reprojection_error(r, p.pts2d_w)
    out["scene/C"] = C
    out["scene/pts3d"] = sc.pts3d_init
    out["scene/cameras"] = np.array(sc.cameras_init)
    out["scene/centers"] = np.array(sc.camera_centers)
    out["scene/pairs"] = np.array(pairs)
    out["scene/err"] = err
    for tag, kw in (("auto", {}), ("predef", {"predef_thr": 2.345}), ("minthr", {"min_thr": 6.0})):
        C_new, cam_thr, n_det = ba_outliers.compute_obs_to_remove(err, p, **kw)
        out["scene/%s/C_new_nan" % tag] = np.isnan(C_new)
        out["scene/%s/cam_thr" % tag] = np.array(cam_thr, dtype=np.float64)
        out["scene/%s/n_detected" % tag] = np.array(int(n_det))
    with contextlib.redirect_stdout(io.StringIO()):
        new_p = ba_outliers.rm_outliers(err, p, verbose=False)
    out["scene/new/C_nan"] = np.isnan(new_p.C)
    out["scene/new/pts_ind"] = new_p.pts_ind
    out["scene/new/cam_ind"] = new_p.cam_ind
    out["scene/new/pts2d"] = new_p.pts2d
    out["scene/new/pts3d"] = new_p.pts3d
    out["scene/new/n_pts_fix"] = np.array(int(new_p.n_pts_fix))
    out["scene/new/pts_prev_indices"] = new_p.pts_prev_indices
    out["scene/new/params_opt"] = new_p.params_opt
    path = os.path.join(HERE, "outliers_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", "scene: %d obs, removed %d, tracks %d -> %d, n_pts_fix %d -> %d"
          % (err.size, int(out["scene/auto/n_detected"]), C.shape[1], new_p.C.shape[1], 30, int(new_p.n_pts_fix)))
    print("thresholds:", out["scene/auto/cam_thr"], [(float(out["elbow/%d/value" % k]), bool(out["elbow/%d/success" % k])) for k in range(len(samples))])


if __name__ == "__main__":
    main()
