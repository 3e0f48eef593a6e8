"""
Bundle-adjustment problem packing: correspondence matrix C -> observation list, camera-parameter
vectors and the variable vector `params_opt`.

Drop-in mirror of the reference class `BundleAdjustmentParameters`
(bundle_adjust/ba_params.py:78-286): same constructor signature, same public fields
(`C, pts3d, cameras, cam_model, pairs_to_triangulate, camera_centers, cam_params_to_optimize,
ref_cam_weight, n_cam, n_pts, n_cam_fix, n_pts_fix, n_cam_opt, n_pts_opt, cam_prev_indices,
pts_prev_indices, cam_params, pts_ind, cam_ind, pts2d, n_obs, n_params, params_opt, pts2d_w`) and
the same methods (`reduce`, `get_vars_ready_for_fun`, `reconstruct_vars`).

What is different: the reference builds the observation list with a Python double loop
(ba_params.py:138-149, 1.4 s at 5e5 observations); here it is one `np.nonzero` over the
transposed visibility mask, which yields the identical (point-major, camera-ascending) order
bit for bit.  The device solver consumes the arrays of this class directly
(`sat_bundleadjust_b200.solver`), so an object built by the *reference* class works too.

Reference quirks that are inherited on purpose:
  * "T" is only honoured when "R" is requested, "K" only when "T" is (ba_params.py:153-163);
  * the K slots of `params_opt` are initialised from `cam_params[:, 3:3+nK]` (ba_params.py:163);
  * `get_vars_ready_for_fun` writes the fixed cameras' parameters into the caller's vector
    (ba_params.py:246-249).
"""
import numpy as np

from . import ba_rotate, cam_utils


class Error(Exception):
    pass


def n_intrinsics(cam_model):
    """number of calibration parameters of a camera model (ba_params.py:161)"""
    return 3 if cam_model == "affine" else 5


def load_cam_params_from_camera(camera, camera_center, cam_model):
    """
    Camera -> flat parameter vector used by the residual (ba_params.py:19-44):
      affine       [roll, pitch, yaw, T0, T1, fx, fy, skew]                    (8)
      perspective  [roll, pitch, yaw, T0, T1, T2, fx, fy, skew, cx, cy]        (11)
      rpc          [0, 0, 0, 0, 0, 0, Cx, Cy, Cz]                              (9)
    """
    if cam_model == "affine":
        K, R, vecT = cam_utils.decompose_affine_camera(camera)
        angles = np.array(ba_rotate.euler_angles_from_R(R))
        return np.hstack((angles.ravel(), vecT.ravel(), K[0, 0], K[1, 1], K[0, 1]))
    if cam_model == "perspective":
        K, R, vecT, _ = cam_utils.decompose_perspective_camera(camera)
        K = K / K[2, 2]
        angles = np.array(ba_rotate.euler_angles_from_R(R))
        return np.hstack((angles.ravel(), vecT.ravel(), K[0, 0], K[1, 1], K[0, 1], K[0, 2], K[1, 2]))
    return np.hstack([np.zeros(6, dtype=np.float32), camera_center])


def load_camera_from_cam_params(cam_params, cam_model):
    """Inverse of `load_cam_params_from_camera` (ba_params.py:47-75)."""
    if cam_model == "affine":
        K = np.array([[cam_params[5], cam_params[7]], [0, cam_params[6]]])
        R = ba_rotate.euler_angles_to_R(*cam_params[0:3].tolist())
        P = cam_utils.compose_affine_camera(K, R, cam_params[3:5])
        return P / P[2, 3]
    if cam_model == "perspective":
        K = np.array([[cam_params[6], cam_params[8], cam_params[9]], [0, cam_params[7], cam_params[10]], [0, 0, 1]])
        R = ba_rotate.euler_angles_to_R(*cam_params[0:3].tolist())
        P = K @ np.hstack((R, cam_params[3:6].reshape((3, 1))))
        return P / P[2, 3]
    return cam_params.reshape((1, 9))


def observations_from_C(C):
    """
    Flatten a (2M x N) correspondence matrix (NaN = unobserved) into the observation list,
    ordered by point index then camera index -- the layout contract of ba_params.py:138-149.

    Returns pts_ind (K,) int64, cam_ind (K,) int64, pts2d (K,2) float64.
    """
    seen = ~np.isnan(C[::2, :])                 # (M, N)
    pts_ind, cam_ind = np.nonzero(seen.T)       # row-major over (N, M): point-major, camera ascending
    pts2d = np.empty((pts_ind.size, 2), dtype=C.dtype)
    pts2d[:, 0] = C[2 * cam_ind, pts_ind]
    pts2d[:, 1] = C[2 * cam_ind + 1, pts_ind]
    return pts_ind.astype(np.int64), cam_ind.astype(np.int64), pts2d


class BundleAdjustmentParameters:
    def __init__(self, C, pts3d, cameras, cam_model, pairs_to_triangulate, camera_centers, d):
        """
        Args (identical to the reference, ba_params.py:79-100):
            C: 2M x N correspondence matrix, NaN where camera m does not see track n
            pts3d: N x 3 initial ECEF coordinates of the tracks
            cameras: M projection matrices (3x4) or M RPC models
            cam_model: "affine" | "perspective" | "rpc"
            pairs_to_triangulate: list of camera-index pairs
            camera_centers: M camera centres (ECEF)
            d: options -- n_cam_fix, n_pts_fix, reduce, verbose, correction_params, ref_cam_weight
        """
        # inputs are copied, options come from `d` with the reference's defaults (ba_params.py:103-115)
        self.C, self.pts3d = C.copy(), pts3d.copy()
        self.cameras, self.camera_centers = cameras.copy(), camera_centers.copy()
        self.pairs_to_triangulate = pairs_to_triangulate.copy()
        self.cam_model = cam_model
        defaults = {"correction_params": ["R"], "ref_cam_weight": 1.0, "n_cam_fix": 0, "n_pts_fix": 0, "verbose": True,
                    "reduce": True}
        opts = {k: d.get(k, v) for k, v in defaults.items()}
        self.cam_params_to_optimize, self.ref_cam_weight = opts["correction_params"], opts["ref_cam_weight"]
        self.n_cam_fix, self.n_pts_fix = opts["n_cam_fix"], opts["n_pts_fix"]
        verbose = opts["verbose"]
        if verbose:
            print("\nDefining bundle adjustment parameters...\n     - cam_params_to_optimize: {}\n".format(
                self.cam_params_to_optimize))

        self.n_cam, self.n_pts = C.shape[0] // 2, C.shape[1]
        self.n_cam_opt, self.n_pts_opt = self.n_cam - self.n_cam_fix, self.n_pts - self.n_pts_fix
        self.cam_prev_indices, self.pts_prev_indices = np.arange(self.n_cam), np.arange(self.n_pts)
        if opts["reduce"]:
            shape_before = C.shape
            self.reduce(C, pts3d, cameras, pairs_to_triangulate, camera_centers)
            if verbose:
                print("C.shape before reduce", shape_before)
                print("C.shape after reduce", self.C.shape)

        # per-camera parameter vectors (M x 8 | 11 | 9)
        self.cam_params = np.array([load_cam_params_from_camera(cam, center, self.cam_model)
                                    for cam, center in zip(self.cameras, self.camera_centers)])

        # observation list (vectorised; same order as the reference's double loop)
        self.pts_ind, self.cam_ind, self.pts2d = observations_from_C(self.C)
        self.n_obs = self.pts2d.shape[0]
        if self.n_obs == 0:
            # the reference dies in np.vstack([]) here (ba_params.py:148)
            raise ValueError("need at least one array to concatenate")

        # variable vector
        opt = self.cam_params_to_optimize
        nK = n_intrinsics(self.cam_model)
        self.n_params = 0
        blocks = []
        if "R" in opt:
            self.n_params += 3
            blocks.append(self.cam_params[:, :3])
            if "T" in opt:
                nT = 2 if self.cam_model == "affine" else 3
                self.n_params += nT
                blocks.append(self.cam_params[:, 3:3 + nT])
                if "K" in opt:
                    self.n_params += nK
                    blocks.append(self.cam_params[:, 3:3 + nK])   # sic: reference slice (ba_params.py:163)
        if not blocks:
            # the reference leaves a python list here and fails on `.ravel()` (ba_params.py:165,172)
            raise AttributeError("'list' object has no attribute 'ravel'")
        cam_params_opt = np.hstack(blocks)
        if "K" in opt and "COMMON_K" in opt:
            K = cam_params_opt[0, -nK:]
            cam_params_opt = np.hstack([cam_params_opt[i, :-nK] for i in range(self.n_cam_opt)])
            cam_params_opt = np.hstack((K, cam_params_opt))
        self.params_opt = np.hstack((cam_params_opt.ravel(), self.pts3d.ravel()))
        self.pts2d_w = np.ones(self.pts2d.shape[0])
        if self.ref_cam_weight > 1.0:
            self.pts2d_w[self.cam_ind == 0] = self.ref_cam_weight

        if verbose:
            for what, total, fixed, free in (("3d points", self.n_pts, self.n_pts_fix, self.n_pts_opt),
                                             ("cameras", self.n_cam, self.n_cam_fix, self.n_cam_opt)):
                print("{} {}, {} fixed and {} to be optimized".format(total, what, fixed, free))
            print("{} parameters to optimize per camera\n".format(self.n_params))

    # ------------------------------------------------------------------------------------------
    def reduce(self, C, pts3d, cameras, pairs_to_triangulate, camera_centers):
        """
        Keep only tracks seen by at least one camera that will be optimised, then drop cameras left
        without observations (ba_params.py:183-219).  Updates the counts of fixed / free items and
        re-indexes `pairs_to_triangulate`.
        """
        seen = ~np.isnan(C[::2, :])
        keep_pt = seen[-self.n_cam_opt:].sum(axis=0).astype(bool)
        self.C = C[:, keep_pt].copy()
        self.pts_prev_indices = np.arange(self.n_pts, dtype=int)[keep_pt]
        self.n_pts_fix -= np.sum(~keep_pt[: self.n_pts_fix])
        self.n_pts_opt -= np.sum(~keep_pt[-self.n_pts_opt:])
        self.pts3d = pts3d[self.pts_prev_indices, :].copy()

        keep_cam = np.sum(~np.isnan(self.C[::2]), axis=1) > 0
        self.cam_prev_indices = np.arange(self.n_cam, dtype=int)[keep_cam]
        self.C = self.C[np.repeat(keep_cam, 2), :]
        self.n_cam = int(self.C.shape[0] / 2)
        self.n_pts = int(self.C.shape[1])
        self.n_cam_fix -= np.sum(~keep_cam[: self.n_cam_fix])
        self.n_cam_opt -= np.sum(~keep_cam[-self.n_cam_opt:])
        self.cameras = [cameras[idx] for idx in self.cam_prev_indices]
        self.camera_centers = [camera_centers[idx] for idx in self.cam_prev_indices]

        new_index = np.full(len(keep_cam), -1)
        new_index[keep_cam] = np.arange(np.sum(keep_cam))
        self.pairs_to_triangulate = [
            (new_index[a], new_index[b]) for [a, b] in pairs_to_triangulate if keep_cam[a] and keep_cam[b]
        ]

    # ------------------------------------------------------------------------------------------
    def common_K(self):
        return "K" in self.cam_params_to_optimize and "COMMON_K" in self.cam_params_to_optimize

    def get_vars_ready_for_fun(self, v):
        """
        Variable vector -> (pts3d (N,3), cam_params (M,P)) as the residual needs them
        (ba_params.py:221-257).  Like the reference this writes the frozen cameras' values into `v`.
        """
        n_params = self.n_params
        K = None
        if self.common_K():
            nK = n_intrinsics(self.cam_model)
            K, v = v[:nK], v[nK:]
            n_params -= nK

        split = self.n_cam * n_params
        pts3d = v[split:].reshape((self.n_pts, 3)).copy()
        if self.n_pts_fix > 0:
            pts3d[: self.n_pts_fix, :] = self.pts3d[: self.n_pts_fix, :]

        cam_opt = v[:split].reshape((self.n_cam, n_params))
        if self.n_cam_fix > 0:
            cam_opt[: self.n_cam_fix, :] = self.cam_params[: self.n_cam_fix, :n_params]
        cam_params = np.hstack((cam_opt, self.cam_params[:, n_params:]))
        if K is not None:
            cam_params[:, -nK:] = K[np.newaxis, :]
        return pts3d, cam_params

    def reconstruct_vars(self, v, pts3d, cameras):
        """
        Variable vector -> corrected 3d points and cameras, written back at the indices they had
        before `reduce` (ba_params.py:259-286).
        """
        self.pts3d_ba, cam_params = self.get_vars_ready_for_fun(v)
        self.cameras_ba = [load_camera_from_cam_params(row, self.cam_model) for row in cam_params]
        wanted = [(key, sl) for key, sl in (("R", slice(0, 3)), ("T", slice(3, 6))) if key in self.cam_params_to_optimize]
        if self.cam_model == "rpc":
            wanted.append(("C", slice(6, 9)))
        self.estimated_params = [{key: row[sl] for key, sl in wanted} for row in cam_params]
        print("\n")      # the reference's driver output has this blank block here (ba_params.py:279)
        corrected_pts3d, corrected_cameras = pts3d.copy(), cameras.copy()
        corrected_pts3d[self.pts_prev_indices] = self.pts3d_ba
        for ba_idx, prev_idx in enumerate(self.cam_prev_indices):
            corrected_cameras[prev_idx] = self.cameras_ba[ba_idx]
        return corrected_pts3d, corrected_cameras
