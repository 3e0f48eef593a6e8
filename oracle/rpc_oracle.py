"""
TEST INFRASTRUCTURE ONLY.  CPU restatement (numpy) of the RPC camera model used on the
bundle-adjustment hot path.  Nothing under sat_bundleadjust_b200/ imports this module.

The reference delegates RPC evaluation to the third-party package `rpcm`
(requirements.txt:9, `rpcm @ git+https://github.com/centreborelli/rpcm.git@localization-origin`,
a branch, not vendored, not installed here).  Its published algorithm is restated from the
places where the reference itself spells it out:
  monomial order (RPC00B)        c/rpc.c:279-298 (eval_pol20), bundle_adjust/ba_rpcfit.py:17-44
  projection + normalisation     c/rpc.c:442-452 (eval_rpci), bundle_adjust/ba_rpcfit.py:47-74
  iterative localisation         c/rpc.c:378-411 (eval_nrpc_iterative), :429-439 (eval_rpc)
  two-view height                c/rpc.c:480-514 (rpc_height)
  triangulation of matches       c/disp_to_h.c:40-65 (stereo_corresp_to_lonlatalt)
  ECEF <-> geodetic              bundle_adjust/geo_utils.py:218-255
  RPC text format                tests/data/images/*.rpc (90 lines `KEY: value [unit]`)
Parity pin: `oracle/_ref/disp_to_h.so`, the reference's own C sources compiled where they lie
(oracle/Makefile), is the known-answer generator for projection / localisation / triangulation
(tests/test_oracle_pin.py) and the golden vectors in tests/golden/rpc_golden.npz come from it.
"""
import numpy as np

A_WGS84 = 6378137.0
ECC = 8.1819190842622e-2


def latlon_to_ecef(lat, lon, alt):
    phi, lam = lat * (np.pi / 180.0), lon * (np.pi / 180.0)
    f = 1 / 298.257223563
    e2 = 1 - (1 - f) * (1 - f)
    nu = A_WGS84 / np.sqrt(1 - e2 * np.sin(phi) * np.sin(phi))
    return ((nu + alt) * np.cos(phi) * np.cos(lam), (nu + alt) * np.cos(phi) * np.sin(lam),
            (nu * (1 - e2) + alt) * np.sin(phi))


def ecef_to_latlon(x, y, z):
    a = A_WGS84
    asq, esq = a ** 2, ECC ** 2
    b = np.sqrt(asq * (1 - esq))
    bsq = b ** 2
    ep = np.sqrt((asq - bsq) / bsq)
    p = np.sqrt((x ** 2) + (y ** 2))
    th = np.arctan2(a * z, b * p)
    lon = np.arctan2(y, x)
    lat = np.arctan2((z + (ep ** 2) * b * (np.sin(th) ** 3)), (p - esq * a * (np.cos(th) ** 3)))
    N = a / (np.sqrt(1 - esq * (np.sin(lat) ** 2)))
    alt = p / np.cos(lat) - N
    return lat * 180 / np.pi, lon * 180 / np.pi, alt


def monomials(lon, lat, alt):
    """The 20 cubic monomials in RPC00B order, rows = terms."""
    one = np.ones_like(lon)
    return np.array([one, lon, lat, alt, lon * lat, lon * alt, lat * alt, lon * lon, lat * lat, alt * alt,
                     lat * lon * alt, lon * lon * lon, lon * lat * lat, lon * alt * alt, lon * lon * lat,
                     lat * lat * lat, lat * alt * alt, lon * lon * alt, lat * lat * alt, alt * alt * alt])


def poly20(c, lon, lat, alt):
    """sum_i c[i] * m_i, accumulated term by term in index order as c/rpc.c:294-297 does."""
    m = monomials(lon, lat, alt)
    r = np.zeros_like(lon, dtype=np.float64)
    for i in range(20):
        r = r + c[i] * m[i]
    return r


_KEYS = [("row_offset", "LINE_OFF", "pixels"), ("col_offset", "SAMP_OFF", "pixels"),
         ("lat_offset", "LAT_OFF", "degrees"), ("lon_offset", "LONG_OFF", "degrees"),
         ("alt_offset", "HEIGHT_OFF", "meters"), ("row_scale", "LINE_SCALE", "pixels"),
         ("col_scale", "SAMP_SCALE", "pixels"), ("lat_scale", "LAT_SCALE", "degrees"),
         ("lon_scale", "LONG_SCALE", "degrees"), ("alt_scale", "HEIGHT_SCALE", "meters")]
_POLYS = [("row_num", "LINE_NUM_COEFF"), ("row_den", "LINE_DEN_COEFF"),
          ("col_num", "SAMP_NUM_COEFF"), ("col_den", "SAMP_DEN_COEFF")]


class RPCModel:
    """Minimal rpcm.RPCModel look-alike (attributes + projection / localization / file IO)."""

    def __init__(self, d=None):
        for attr, _, _ in _KEYS:
            setattr(self, attr, 0.0)
        for attr, _ in _POLYS:
            setattr(self, attr, [0.0] * 20)
        if d:
            for attr, key, _ in _KEYS:
                setattr(self, attr, float(d[key]))
            for attr, key in _POLYS:
                if key in d:      # GeoTIFF-tag style: one string with the coefficients (what ba_rpcfit.initialize_rpc passes)
                    setattr(self, attr, [float(v) for v in str(d[key]).split()])
                else:
                    setattr(self, attr, [float(d["%s_%d" % (key, i + 1)]) for i in range(20)])

    @classmethod
    def from_file(cls, path):
        d = {}
        with open(path) as f:
            for line in f:
                if ":" in line:
                    k, v = line.split(":", 1)
                    d[k.strip()] = v.split()[0]
        return cls(d)

    def to_dict(self):
        d = {key: getattr(self, attr) for attr, key, _ in _KEYS}
        for attr, key in _POLYS:
            for i, c in enumerate(getattr(self, attr)):
                d["%s_%d" % (key, i + 1)] = float(c)
        return d

    def write_to_file(self, path):
        with open(path, "w") as f:
            for attr, key, unit in _KEYS:
                f.write("%s: %.12f %s\n" % (key, getattr(self, attr), unit))
            for attr, key in _POLYS:
                for i, c in enumerate(getattr(self, attr)):
                    f.write("%s_%d: %.12f\n" % (key, i + 1, c))

    # -- forward model ---------------------------------------------------------------------------
    def _normalised_projection(self, nlon, nlat, nalt):
        ncol = poly20(self.col_num, nlon, nlat, nalt) / poly20(self.col_den, nlon, nlat, nalt)
        nrow = poly20(self.row_num, nlon, nlat, nalt) / poly20(self.row_den, nlon, nlat, nalt)
        return ncol, nrow

    def projection(self, lon, lat, alt):
        lon, lat, alt = [np.asarray(v, dtype=np.float64) for v in (lon, lat, alt)]
        nlon = (lon - self.lon_offset) / self.lon_scale
        nlat = (lat - self.lat_offset) / self.lat_scale
        nalt = (alt - self.alt_offset) / self.alt_scale
        ncol, nrow = self._normalised_projection(nlon, nlat, nalt)
        return ncol * self.col_scale + self.col_offset, nrow * self.row_scale + self.row_offset

    # -- inverse model, iterative (c/rpc.c:378-411) ---------------------------------------------------
    def localization(self, col, row, alt, delta=1.0, max_iter=200):
        col, row, alt = np.broadcast_arrays(*[np.asarray(v, dtype=np.float64) for v in (col, row, alt)])
        shape = col.shape
        xf = ((col - self.col_offset) / self.col_scale).ravel()
        yf = ((row - self.row_offset) / self.row_scale).ravel()
        nalt = ((alt - self.alt_offset) / self.alt_scale).ravel()
        lon = np.full(xf.shape, -1.0 * delta)
        lat = np.full(xf.shape, -1.0 * delta)
        eps = 2.0 * delta
        active = np.ones(xf.shape, dtype=bool)
        for _ in range(max_iter):
            x0, y0 = self._normalised_projection(lon, lat, nalt)
            active = active & ((x0 - xf) ** 2 + (y0 - yf) ** 2 > 1e-18)
            if not active.any():
                break
            x1, y1 = self._normalised_projection(lon + eps, lat, nalt)
            x2, y2 = self._normalised_projection(lon, lat + eps, nalt)
            ux, uy = xf - x0, yf - y0
            e1x, e1y, e2x, e2y = x1 - x0, y1 - y0, x2 - x0, y2 - y0
            det = e1x * e2y - e1y * e2x
            a0 = (e2y * ux - e2x * uy) / det
            a1 = (-e1y * ux + e1x * uy) / det
            lon = np.where(active, lon + a0 * eps, lon)
            lat = np.where(active, lat + a1 * eps, lat)
            eps = 0.1
        return (lon * self.lon_scale + self.lon_offset).reshape(shape), \
               (lat * self.lat_scale + self.lat_offset).reshape(shape)

    def project_ecef(self, pts3d):
        lat, lon, alt = ecef_to_latlon(pts3d[:, 0], pts3d[:, 1], pts3d[:, 2])
        col, row = self.projection(lon, lat, alt)
        return np.stack((col, row), axis=1)
