"""
Seeded synthetic satellite scenes for parity tests and benchmarks.

Pure numpy, host side, no reference code involved.  A scene is:
  * ground points uniform in a `box_m` box (default 10 km x 10 km x 500 m) around
    lat 11.02, lon -72.71, alt 3500 m (the footprint of the reference's test RPCs),
  * M pinhole cameras at ~500 km altitude, off-nadir up to +-25 deg, f ~ 1e6 px
    (GSD ~0.5 m), P = K R [I | -C] normalised so that P[2,3] = 1,
  * optionally their first-order (affine) expansion at the scene centre,
  * each track seen by each camera with probability `p_vis`; tracks with <2 observations dropped,
  * observations = exact projection + N(0, noise_px^2), `outlier_frac` of them with an
    extra N(0, outlier_px^2) term,
  * initial state = true points + N(0, pts_sigma_m^2) cast to float32 (as the reference's
    triangulation returns float32, ft_triangulate.py:89), Euler angles + N(0, ang_sigma^2).
"""
from dataclasses import dataclass, field

import numpy as np

from . import cam_utils, geo_utils
from .ba_params import BundleAdjustmentParameters, load_cam_params_from_camera

SCENE_LAT, SCENE_LON, SCENE_ALT = 11.02, -72.71, 3500.0


@dataclass
class Scene:
    cam_model: str
    cameras: list              # true cameras (3x4) -- or RPC models for cam_model == "rpc"
    cameras_init: list         # perturbed cameras handed to BA
    camera_centers: list
    pts3d_true: np.ndarray     # (N,3) float64
    pts3d_init: np.ndarray     # (N,3) float32 values
    pts_ind: np.ndarray        # (K,) int64, non-decreasing
    cam_ind: np.ndarray        # (K,) int64
    pts2d: np.ndarray          # (K,2) float64
    seed: int = 0
    meta: dict = field(default_factory=dict)

    @property
    def n_cam(self):
        return len(self.cameras)

    @property
    def n_pts(self):
        return self.pts3d_true.shape[0]

    @property
    def n_obs(self):
        return self.pts_ind.size

    def correspondence_matrix(self):
        """(2M x N) matrix with NaN where unobserved -- the reference's input format."""
        C = np.full((2 * self.n_cam, self.n_pts), np.nan)
        C[2 * self.cam_ind, self.pts_ind] = self.pts2d[:, 0]
        C[2 * self.cam_ind + 1, self.pts_ind] = self.pts2d[:, 1]
        return C


def enu_basis(lat_deg, lon_deg):
    phi, lam = np.deg2rad(lat_deg), np.deg2rad(lon_deg)
    east = np.array([-np.sin(lam), np.cos(lam), 0.0])
    north = np.array([-np.sin(phi) * np.cos(lam), -np.sin(phi) * np.sin(lam), np.cos(phi)])
    up = np.array([np.cos(phi) * np.cos(lam), np.cos(phi) * np.sin(lam), np.sin(phi)])
    return east, north, up


def make_perspective_camera(rng, centre_ecef, basis, altitude_m=500e3, max_off_nadir_deg=25.0,
                            focal_px=1.0e6, principal=(10000.0, 10000.0)):
    """A pinhole camera looking at `centre_ecef` from `altitude_m` with a random off-nadir tilt."""
    east, north, up = basis
    tilt = np.deg2rad(rng.uniform(-max_off_nadir_deg, max_off_nadir_deg))
    az = rng.uniform(0, 2 * np.pi)
    horiz = np.cos(az) * east + np.sin(az) * north
    direction = np.cos(tilt) * up + np.sin(tilt) * horiz
    C = centre_ecef + direction * (altitude_m / np.cos(tilt))
    zc = (centre_ecef - C) / np.linalg.norm(centre_ecef - C)      # optical axis, towards the ground
    heading = rng.uniform(0, 2 * np.pi)
    a = np.cos(heading) * east + np.sin(heading) * north
    xc = a - np.dot(a, zc) * zc
    xc /= np.linalg.norm(xc)
    yc = np.cross(zc, xc)
    R = np.vstack((xc, yc, zc))
    f = focal_px * rng.uniform(0.95, 1.05)
    K = np.array([[f, rng.uniform(-2, 2), principal[0]], [0, f * rng.uniform(0.999, 1.001), principal[1]], [0, 0, 1.0]])
    P = cam_utils.compose_perspective_camera(K, R, C)
    return P / P[2, 3], C


def affine_expansion(P, X0):
    """First-order Taylor expansion of a pinhole camera at X0 -> affine 3x4 matrix."""
    h = P @ np.append(X0, 1.0)
    u0 = h[:2] / h[2]
    J = (P[:2, :3] - np.outer(u0, P[2, :3])) / h[2]
    A = np.zeros((3, 4))
    A[:2, :3] = J
    A[:2, 3] = u0 - J @ X0
    A[2, 3] = 1.0
    return A


def perturb_camera(P, cam_model, d_angles):
    """Add `d_angles` to the Euler angles of a camera, keeping T and K (what correction 'R' moves)."""
    from .ba_params import load_camera_from_cam_params
    v = load_cam_params_from_camera(P, None, cam_model).astype(np.float64)
    v[:3] += d_angles
    return load_camera_from_cam_params(v, cam_model)


def make_scene(n_cam=10, n_tracks=1000, p_vis=0.5, cam_model="perspective", seed=0,
               box_m=(10e3, 10e3, 500.0), noise_px=0.5, outlier_frac=0.02, outlier_px=20.0,
               pts_sigma_m=1.0, ang_sigma=1e-6, min_obs=2, visibility=None):
    """
    Build a seeded synthetic scene.  `n_tracks` is the number of candidate ground points; tracks seen
    by fewer than `min_obs` cameras are dropped, so the scene ends up with slightly fewer.
    `visibility` (optional): callable(rng, n_tracks, n_cam) -> boolean (n_tracks, n_cam) matrix replacing the
    independent draws with probability p_vis (used to build scenes whose shards differ in structure).
    """
    if cam_model not in ("perspective", "affine"):
        raise ValueError("make_scene builds matrix cameras; see make_rpc_scene for cam_model='rpc'")
    rng = np.random.default_rng(seed)
    basis = enu_basis(SCENE_LAT, SCENE_LON)
    centre = np.array(geo_utils.latlon_to_ecef_custom(SCENE_LAT, SCENE_LON, SCENE_ALT))

    cams, centers = [], []
    for _ in range(n_cam):
        P, C = make_perspective_camera(rng, centre, basis)
        if cam_model == "affine":
            P = affine_expansion(P, centre)
        cams.append(P)
        centers.append(C)

    enu = rng.uniform(-0.5, 0.5, size=(n_tracks, 3)) * np.array(box_m)
    pts = centre + enu[:, :1] * basis[0] + enu[:, 1:2] * basis[1] + enu[:, 2:3] * basis[2]

    seen = rng.random((n_tracks, n_cam)) < p_vis if visibility is None else np.asarray(visibility(rng, n_tracks, n_cam), dtype=bool)
    keep = seen.sum(axis=1) >= min_obs
    pts, seen = pts[keep], seen[keep]
    pts_ind, cam_ind = np.nonzero(seen)

    pts2d = np.empty((pts_ind.size, 2))
    for j in range(n_cam):
        sel = cam_ind == j
        pts2d[sel] = cam_utils.apply_projection_matrix(cams[j], pts[pts_ind[sel]])
    pts2d += rng.normal(0.0, noise_px, size=pts2d.shape)
    bad = rng.random(pts_ind.size) < outlier_frac
    pts2d[bad] += rng.normal(0.0, outlier_px, size=(int(bad.sum()), 2))

    pts_init = (pts + rng.normal(0.0, pts_sigma_m, size=pts.shape)).astype(np.float32)
    cams_init = [perturb_camera(P, cam_model, rng.normal(0.0, ang_sigma, size=3)) for P in cams]

    return Scene(cam_model=cam_model, cameras=cams, cameras_init=cams_init, camera_centers=centers,
                 pts3d_true=pts, pts3d_init=pts_init, pts_ind=pts_ind.astype(np.int64),
                 cam_ind=cam_ind.astype(np.int64), pts2d=pts2d, seed=seed,
                 meta=dict(n_cam=n_cam, n_tracks=n_tracks, p_vis=p_vis, noise_px=noise_px,
                           outlier_frac=outlier_frac, pts_sigma_m=pts_sigma_m, ang_sigma=ang_sigma))


def scene_to_params(scene, correction_params=("R",), n_cam_fix=0, n_pts_fix=0, ref_cam_weight=1.0,
                    params_cls=BundleAdjustmentParameters, verbose=False):
    """
    Feed a scene to a `BundleAdjustmentParameters` class (ours by default; the reference's own class
    can be passed as `params_cls` in the build container to cross-check the packing).
    """
    d = {"correction_params": list(correction_params), "n_cam_fix": n_cam_fix, "n_pts_fix": n_pts_fix,
         "ref_cam_weight": ref_cam_weight, "reduce": False, "verbose": verbose}
    pairs = [(a, b) for a in range(scene.n_cam) for b in range(a + 1, scene.n_cam)]
    return params_cls(scene.correspondence_matrix(), scene.pts3d_init, list(scene.cameras_init), scene.cam_model,
                      pairs, list(scene.camera_centers), d)


class SparseParams:
    """
    The fields of `BundleAdjustmentParameters` that the solver reads, built WITHOUT the dense
    (2M x N) correspondence matrix: at time-series scale (BASELINE config 4: 300 views x 5e6 tracks) that matrix alone
    is 24 GB.  Same observation order (by track, camera ascending), same parameter vector layout
    (bundle_adjust/ba_params.py:138-173) -- test_sparse_scene_matches_dense_packing pins it against the class.
    """

    def __init__(self, scene, correction_params=("R", "T"), n_cam_fix=0, n_pts_fix=0):
        self.cam_model = scene.cam_model
        self.cameras, self.camera_centers = list(scene.cameras_init), list(scene.camera_centers)
        self.cam_params_to_optimize = list(correction_params)
        self.n_cam, self.n_pts = scene.n_cam, scene.n_pts
        self.n_cam_fix, self.n_pts_fix = n_cam_fix, n_pts_fix
        self.pts3d = np.asarray(scene.pts3d_init)
        self.pts_ind, self.cam_ind, self.pts2d = scene.pts_ind, scene.cam_ind, scene.pts2d
        self.n_obs = int(self.pts_ind.size)
        self.cam_params = np.array([load_cam_params_from_camera(c, ctr, self.cam_model)
                                    for c, ctr in zip(self.cameras, self.camera_centers)])
        nT = 2 if self.cam_model == "affine" else 3
        self.n_params = 3 + (nT if "T" in correction_params else 0) if "R" in correction_params else 0
        if "K" in correction_params:
            raise NotImplementedError("SparseParams covers the R / R+T corrections of the large synthetic configs")
        self.params_opt = np.hstack((self.cam_params[:, : self.n_params].ravel(), self.pts3d.astype(np.float64).ravel()))
        self.pts2d_w = np.ones(self.n_obs)


def make_scene_sparse(n_cam=300, n_tracks=100000, p_vis=0.02, cam_model="perspective", seed=0, box_m=(10e3, 10e3, 500.0),
                      noise_px=0.5, outlier_frac=0.02, outlier_px=20.0, pts_sigma_m=1.0, ang_sigma=1e-6, min_obs=2):
    """
    `make_scene` for many cameras: the visibility is drawn sparsely (geometric gaps between the observed (track, camera)
    cells of the row-major N x M table), so memory and time are O(observations), not O(N x M).  Not bit-identical to
    make_scene for the same seed (different random draws), same statistics.
    """
    if cam_model not in ("perspective", "affine"):
        raise ValueError("matrix cameras only")
    rng = np.random.default_rng(seed)
    basis = enu_basis(SCENE_LAT, SCENE_LON)
    centre = np.array(geo_utils.latlon_to_ecef_custom(SCENE_LAT, SCENE_LON, SCENE_ALT))
    cams, centers = [], []
    for _ in range(n_cam):
        P, C = make_perspective_camera(rng, centre, basis)
        if cam_model == "affine":
            P = affine_expansion(P, centre)
        cams.append(P)
        centers.append(C)
    total = n_tracks * n_cam
    n_draw = int(total * p_vis * 1.05 + 10 * np.sqrt(total * p_vis) + 100)
    flat = np.cumsum(rng.geometric(p_vis, size=n_draw)) - 1
    flat = flat[flat < total]
    trk, cam_ind = np.divmod(flat, n_cam)
    counts = np.bincount(trk, minlength=n_tracks)
    keep_trk = counts >= min_obs
    new_id = np.cumsum(keep_trk) - 1
    sel = keep_trk[trk]
    pts_ind, cam_ind = new_id[trk[sel]].astype(np.int64), cam_ind[sel].astype(np.int64)
    n_pts = int(keep_trk.sum())
    enu = rng.uniform(-0.5, 0.5, size=(n_pts, 3)) * np.array(box_m)
    pts = centre + enu[:, :1] * basis[0] + enu[:, 1:2] * basis[1] + enu[:, 2:3] * basis[2]
    pts2d = np.empty((pts_ind.size, 2))
    order = np.argsort(cam_ind, kind="stable")
    bounds = np.searchsorted(cam_ind[order], np.arange(n_cam + 1))
    for j in range(n_cam):
        o = order[bounds[j]: bounds[j + 1]]
        pts2d[o] = cam_utils.apply_projection_matrix(cams[j], pts[pts_ind[o]])
    pts2d += rng.normal(0.0, noise_px, size=pts2d.shape)
    bad = rng.random(pts_ind.size) < outlier_frac
    pts2d[bad] += rng.normal(0.0, outlier_px, size=(int(bad.sum()), 2))
    pts_init = (pts + rng.normal(0.0, pts_sigma_m, size=pts.shape)).astype(np.float32)
    cams_init = [perturb_camera(P, cam_model, rng.normal(0.0, ang_sigma, size=3)) for P in cams]
    return Scene(cam_model=cam_model, cameras=cams, cameras_init=cams_init, camera_centers=centers, pts3d_true=pts,
                 pts3d_init=pts_init, pts_ind=pts_ind, cam_ind=cam_ind, pts2d=pts2d, seed=seed,
                 meta=dict(n_cam=n_cam, n_tracks=n_tracks, p_vis=p_vis, sparse=True))


def make_rpc_scene(rpc_tables, camera_centers, n_tracks=100000, p_vis=0.8, seed=0, noise_px=0.5, outlier_frac=0.02, outlier_px=20.0,
                   pts_sigma_m=1.0, min_obs=2):
    """
    Synthetic scene for cam_model == "rpc" (the pipeline's default model, ba_pipeline.py:83): ground points drawn inside the
    validity box shared by the given RPC cameras ((M, 90) coefficient tables, layout of include/sba_b200.h), observations =
    their RPC projections (one batched GPU launch, sba_rpc_projection_batch) + Gaussian noise + a few outliers, random
    visibility.  Needs the CUDA library (the generator is used by bench.py and the GPU tests only).
    """
    from . import rpc_model
    from .ba_rpcfit import _model_from_table
    rng = np.random.default_rng(seed)
    T = np.asarray(rpc_tables, dtype=np.float64)
    M = T.shape[0]
    cams = [_model_from_table(T[j]) for j in range(M)]
    # common validity box in normalised coordinates: |x| <= 0.45 of every camera's scale around the mean offset
    lat0, lon0, alt0 = T[:, 2].mean(), T[:, 3].mean(), T[:, 4].mean()
    dlat = 0.45 * T[:, 7].min() - np.abs(T[:, 2] - lat0).max()
    dlon = 0.45 * T[:, 8].min() - np.abs(T[:, 3] - lon0).max()
    dalt = 0.45 * T[:, 9].min() - np.abs(T[:, 4] - alt0).max()
    if min(dlat, dlon, dalt) <= 0:
        raise ValueError("the RPC cameras do not share a validity box")
    lat = lat0 + dlat * rng.uniform(-1, 1, n_tracks)
    lon = lon0 + dlon * rng.uniform(-1, 1, n_tracks)
    alt = alt0 + dalt * rng.uniform(-1, 1, n_tracks)
    seen = rng.random((n_tracks, M)) < p_vis
    keep = seen.sum(axis=1) >= min_obs
    lat, lon, alt, seen = lat[keep], lon[keep], alt[keep], seen[keep]
    col, row = rpc_model.projection_batch(cams, lon, lat, alt)                     # (M, N)
    pts_ind, cam_ind = np.nonzero(seen)
    pts2d = np.stack((col[cam_ind, pts_ind], row[cam_ind, pts_ind]), axis=1)
    pts2d += noise_px * rng.standard_normal(pts2d.shape)
    bad = rng.random(pts_ind.size) < outlier_frac
    pts2d[bad] += outlier_px * rng.standard_normal((int(bad.sum()), 2))
    x, y, z = geo_utils.latlon_to_ecef_custom(lat, lon, alt)
    pts_true = np.stack((x, y, z), axis=1)
    pts_init = (pts_true + pts_sigma_m * rng.standard_normal(pts_true.shape)).astype(np.float32).astype(np.float64)
    return Scene(cam_model="rpc", cameras=cams, cameras_init=cams, camera_centers=[np.asarray(c, dtype=np.float64) for c in camera_centers],
                 pts3d_true=pts_true, pts3d_init=pts_init, pts_ind=pts_ind.astype(np.int64), cam_ind=cam_ind.astype(np.int64),
                 pts2d=pts2d, seed=seed, meta={"box_deg_m": (float(dlat), float(dlon), float(dalt))})
