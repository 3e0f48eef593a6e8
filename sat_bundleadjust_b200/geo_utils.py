"""
Geodesy helpers that sit on the bundle-adjustment hot path (host-side numpy form).

Only the two closed-form converters the reference calls inside every RPC projection are
provided (reference: bundle_adjust/geo_utils.py:218-233 `latlon_to_ecef_custom`,
:236-255 `ecef_to_latlon_custom`).  The batched device versions live in
csrc/sba_geodesy.cuh and are what the solver and the RPC kernels use; these numpy
forms exist for scene set-up and for the host-side mirror of the reference API.
UTM / GeoJSON helpers of the reference are out of scope.
"""
import numpy as np

WGS84_A = 6378137.0
WGS84_INV_F = 298.257223563
# eccentricity constant hard-coded by the reference's inverse conversion (geo_utils.py:241)
REF_ECC = 8.1819190842622e-2


def latlon_to_ecef_custom(lat, lon, alt):
    """geodetic (deg, deg, m) -> geocentric (x, y, z) in metres."""
    phi = lat * (np.pi / 180.0)
    lam = lon * (np.pi / 180.0)
    f = 1 / WGS84_INV_F
    e2 = 1 - (1 - f) * (1 - f)
    s = np.sin(phi)
    nu = WGS84_A / np.sqrt(1 - e2 * s * s)
    x = (nu + alt) * np.cos(phi) * np.cos(lam)
    y = (nu + alt) * np.cos(phi) * np.sin(lam)
    z = (nu * (1 - e2) + alt) * s
    return x, y, z


def ecef_to_latlon_custom(x, y, z):
    """geocentric (x, y, z) -> geodetic (lat deg, lon deg, alt m), one Bowring step."""
    a = WGS84_A
    a2 = a ** 2
    e2 = REF_ECC ** 2
    b = np.sqrt(a2 * (1 - e2))
    b2 = b ** 2
    ep = np.sqrt((a2 - b2) / b2)
    p = np.sqrt((x ** 2) + (y ** 2))
    th = np.arctan2(a * z, b * p)
    lon = np.arctan2(y, x)
    lat = np.arctan2((z + (ep ** 2) * b * (np.sin(th) ** 3)), (p - e2 * a * (np.cos(th) ** 3)))
    nu = a / (np.sqrt(1 - e2 * (np.sin(lat) ** 2)))
    alt = p / np.cos(lat) - nu
    return lat * 180 / np.pi, lon * 180 / np.pi, alt
