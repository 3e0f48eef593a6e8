"""
RPC camera model with batched GPU evaluation.

The reference takes its RPC objects from the third-party package `rpcm` (requirements.txt:9, not vendored,
not installed here); the hot path only touches the duck-typed surface listed in SURVEY.md section 8b:
attributes `row_num, row_den, col_num, col_den, {row,col,lat,lon,alt}_{offset,scale}`, the methods
`projection(lon, lat, alt)`, `localization(col, row, alt)`, `write_to_file(path)`, the constructor from a dict
of RPC00B keys and `rpc_from_rpc_file(path)`.  This class provides exactly that surface; projection and
localisation of whole arrays run in one sm_100a kernel each (csrc/sba_rpc.cu) instead of numpy expressions
(`rpcm`) or a per-point C call (c/rpc.c:442-452, :378-439).
"""
import numpy as np

from . import _lib

_SCALARS = [("row_offset", "LINE_OFF", "pixels"), ("col_offset", "SAMP_OFF", "pixels"),
            ("lat_offset", "LAT_OFF", "degrees"), ("lon_offset", "LONG_OFF", "degrees"),
            ("alt_offset", "HEIGHT_OFF", "meters"), ("row_scale", "LINE_SCALE", "pixels"),
            ("col_scale", "SAMP_SCALE", "pixels"), ("lat_scale", "LAT_SCALE", "degrees"),
            ("lon_scale", "LONG_SCALE", "degrees"), ("alt_scale", "HEIGHT_SCALE", "meters")]
_POLYS = [("row_num", "LINE_NUM_COEFF"), ("row_den", "LINE_DEN_COEFF"),
          ("col_num", "SAMP_NUM_COEFF"), ("col_den", "SAMP_DEN_COEFF")]


class RPCModel:
    def __init__(self, d=None):
        for attr, _, _ in _SCALARS:
            setattr(self, attr, 0.0)
        for attr, _ in _POLYS:
            setattr(self, attr, [0.0] * 20)
        if d:
            for attr, key, _ in _SCALARS:
                setattr(self, attr, float(d[key]))
            for attr, key in _POLYS:
                if key in d:      # "LINE_NUM_COEFF": "c1 c2 ..." (GeoTIFF tag style) or a list
                    v = d[key].split() if isinstance(d[key], str) else d[key]
                    vals = [float(x) for x in v]
                    vals = vals + [0.0] * (20 - len(vals)) if len(vals) < 20 else vals[:20]
                else:
                    vals = [float(d["%s_%d" % (key, i + 1)]) for i in range(20)]
                setattr(self, attr, vals)

    # -- table used by the C ABI (layout of include/sba_b200.h) --------------------------------
    def table(self):
        return np.concatenate([[getattr(self, a) for a, _, _ in _SCALARS]] +
                              [np.asarray(getattr(self, a), dtype=np.float64) for a, _ in _POLYS])

    def projection(self, lon, lat, alt):
        """(lon, lat, alt) -> (col, row); scalars or arrays (broadcast like rpcm)."""
        lon, lat, alt = np.broadcast_arrays(*[np.asarray(v, dtype=np.float64) for v in (lon, lat, alt)])
        shape = lon.shape
        a, b, c = [np.ascontiguousarray(v.ravel()) for v in (lon, lat, alt)]
        col, row = np.empty(a.size), np.empty(a.size)
        lib = _lib.load()
        _lib.check(lib.sba_rpc_projection(_lib.dptr(self.table()), _lib.dptr(a), _lib.dptr(b), _lib.dptr(c), a.size,
                                          _lib.dptr(col), _lib.dptr(row)))
        if shape == ():
            return float(col[0]), float(row[0])
        return col.reshape(shape), row.reshape(shape)

    def localization(self, col, row, alt, delta=1.0):
        """(col, row, alt) -> (lon, lat) by the iterative inversion of the projection."""
        col, row, alt = np.broadcast_arrays(*[np.asarray(v, dtype=np.float64) for v in (col, row, alt)])
        shape = col.shape
        a, b, c = [np.ascontiguousarray(v.ravel()) for v in (col, row, alt)]
        lon, lat = np.empty(a.size), np.empty(a.size)
        lib = _lib.load()
        _lib.check(lib.sba_rpc_localization(_lib.dptr(self.table()), _lib.dptr(a), _lib.dptr(b), _lib.dptr(c), a.size,
                                            float(delta), _lib.dptr(lon), _lib.dptr(lat)))
        if shape == ():
            return float(lon[0]), float(lat[0])
        return lon.reshape(shape), lat.reshape(shape)

    def projection_from_ecef(self, pts3d):
        """ECEF (N,3) -> (N,2) (col,row): geodetic conversion + projection fused in one kernel
        (replaces cam_utils.apply_rpc_projection, cam_utils.py:217-231)."""
        x = np.ascontiguousarray(pts3d, dtype=np.float64)
        out = np.empty((x.shape[0], 2))
        lib = _lib.load()
        _lib.check(lib.sba_rpc_projection_ecef(_lib.dptr(self.table()), _lib.dptr(x), x.shape[0], _lib.dptr(out)))
        return out

    def to_geotiff_dict(self):
        d = {key: getattr(self, attr) for attr, key, _ in _SCALARS}
        for attr, key in _POLYS:
            d[key] = " ".join("%.15e" % c for c in getattr(self, attr))
        return d

    def write_to_file(self, path):
        with open(path, "w") as f:
            for attr, key, unit in _SCALARS:
                f.write("%s: %.12f %s\n" % (key, getattr(self, attr), unit))
            for attr, key in _POLYS:
                for i, c in enumerate(getattr(self, attr)):
                    f.write("%s_%d: %.12f\n" % (key, i + 1, c))


def rpc_from_rpc_file(path):
    """Reads the RPC00B text format of the reference's tests/data/images/*.rpc (`KEY: value [unit]`)."""
    d = {}
    with open(path) as f:
        for line in f:
            if ":" in line:
                k, v = line.split(":", 1)
                d[k.strip()] = v.split()[0]
    return RPCModel(d)


def _batch(kind, models, a, b, c, delta=1.0):
    """kind 0: projection, 1: localisation, for a list of RPC models in one launch.  a, b, c: (n,) shared by all models or (len(models), n)."""
    tables = np.ascontiguousarray(np.stack([np.asarray(m.table(), dtype=np.float64) for m in models]))
    a, b, c = [np.ascontiguousarray(v, dtype=np.float64) for v in np.broadcast_arrays(a, b, c)]
    shared = a.ndim == 1
    if not shared and (a.ndim != 2 or a.shape[0] != len(models)):
        raise ValueError("expected (n,) arrays shared by all models or (n_models, n) arrays")
    n = a.shape[-1]
    o0, o1 = np.empty((len(models), n)), np.empty((len(models), n))
    lib = _lib.load()
    if kind == 0:
        _lib.check(lib.sba_rpc_projection_batch(_lib.dptr(tables), len(models), _lib.dptr(a), _lib.dptr(b), _lib.dptr(c), n, int(shared),
                                                _lib.dptr(o0), _lib.dptr(o1)))
    else:
        _lib.check(lib.sba_rpc_localization_batch(_lib.dptr(tables), len(models), _lib.dptr(a), _lib.dptr(b), _lib.dptr(c), n, int(shared),
                                                  float(delta), _lib.dptr(o0), _lib.dptr(o1)))
    return o0, o1


def projection_batch(models, lon, lat, alt):
    """(lon, lat, alt) -> (col, row) through every model of `models` in ONE kernel launch; returns two (len(models), n) arrays."""
    return _batch(0, models, lon, lat, alt)


def localization_batch(models, col, row, alt, delta=1.0):
    """(col, row, alt) -> (lon, lat) through every model of `models` in ONE kernel launch; returns two (len(models), n) arrays."""
    return _batch(1, models, col, row, alt, delta)
