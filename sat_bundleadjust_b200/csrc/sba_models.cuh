// Per-observation reprojection residual and analytic Jacobian for the three camera models of
// sat-bundleadjust, plus the robust-loss re-weighting.  Header-only, __host__ __device__, FP64.
//
// Reference behaviour restated here (not translated: the reference evaluates the residual with numpy
// array temporaries and differentiates it by finite differences; this file is the closed form):
//   rotation  R = Rz(g) Ry(b) Rx(a) applied x-, y-, z-axis in turn      bundle_adjust/ba_core.py:36-56
//   perspective projection                                             bundle_adjust/ba_core.py:84-107
//   affine projection                                                  bundle_adjust/ba_core.py:59-81
//   RPC with corrective rotation  X' = R (X - T - C) + C               bundle_adjust/ba_core.py:110-154
//   ECEF -> geodetic (one Bowring step, e = 8.1819190842622e-2)        bundle_adjust/geo_utils.py:236-255
//   20-term cubic, RPC00B monomial order                               c/rpc.c:279-298
//   robust losses and the Jacobian/residual rescale                    scipy/optimize/_lsq/least_squares.py:183-240,
//                                                                      scipy/optimize/_lsq/common.py:720-731
// Camera parameter vector (bundle_adjust/ba_params.py:19-44):
//   affine       [a b g | T0 T1 | fx fy skew]                (8)
//   perspective  [a b g | T0 T1 T2 | fx fy skew cx cy]       (11)
//   rpc          [a b g | T0 T1 T2 | Cx Cy Cz]               (9; C is never optimised)
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define SBA_HD __host__ __device__ __forceinline__
#else
#define SBA_HD inline
#endif

namespace sba {

// Reciprocal and reciprocal square root for the per-observation hot loops: FP32 special-function seed + two FP64
// Newton steps (1-2 ulp).  Unlike 1.0 / x, __drcp_rn or rsqrt() this has no slow-path branch (denormal / huge
// arguments do not occur: depths ~5e5 m, damped 3x3 pivots, 1 + z of the robust losses below 1e10), so a warp issues
// ~8 straight-line instructions instead of ~15 plus a convergence barrier.
SBA_HD double fast_rcp(double x)
{
#ifdef __CUDA_ARCH__
    double r = (double)__frcp_rn((float)x);
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
#else
    return 1.0 / x;
#endif
}
SBA_HD double fast_rsqrt(double x)
{
#ifdef __CUDA_ARCH__
    double r = (double)rsqrtf((float)x);
    const double hx = 0.5 * x;
    r = r * fma(-hx * r, r, 1.5);
    r = r * fma(-hx * r, r, 1.5);
    return r;
#else
    return 1.0 / sqrt(x);
#endif
}

enum Model { MODEL_AFFINE = 0, MODEL_PERSPECTIVE = 1, MODEL_RPC = 2 };
enum Loss { LOSS_LINEAR = 0, LOSS_HUBER = 1, LOSS_SOFT_L1 = 2, LOSS_CAUCHY = 3, LOSS_ARCTAN = 4 };

constexpr int CAMREC_STRIDE = 40;   // doubles per prepared camera record
constexpr int RPC_TAB_STRIDE = 90;  // 10 normalisation constants + 4 x 20 coefficients
constexpr int MAX_CAM_PARAMS = 11;

// Prepared camera record (built once per evaluation by k_prepare_cameras / k_step), CAMREC_STRIDE doubles:
//   [0..5]   cos a, sin a, cos b, sin b, cos g, sin g
//   [6..8]   T (affine: T0 T1 -)
//   [9..13]  perspective fx fy skew cx cy ; affine fx fy skew ; rpc C
//   [14..22] R = Rz Ry Rx, row-major
//   [23..31] KR, row-major: perspective (fx R0 + skew R1 + cx R2, fy R1 + cy R2, R2); affine (fx R0 + skew R1, fy R1, 0)
//   [32..34] KT: perspective (fx T0 + skew T1 + cx T2, fy T1 + cy T2, T2); affine (fx T0 + skew T1, fy T1, 0)
constexpr int CR_R = 14, CR_KR = 23, CR_KT = 32;
struct CamRec {
    double ca, sa, cb, sb, cg, sg;
    double t0, t1, t2;
    double k0, k1, k2, k3, k4;
};

SBA_HD CamRec load_camrec(const double* __restrict__ r)
{
    CamRec c;
    c.ca = r[0]; c.sa = r[1]; c.cb = r[2]; c.sb = r[3]; c.cg = r[4]; c.sg = r[5];
    c.t0 = r[6]; c.t1 = r[7]; c.t2 = r[8];
    c.k0 = r[9]; c.k1 = r[10]; c.k2 = r[11]; c.k3 = r[12]; c.k4 = r[13];
    return c;
}

// Rotation of a point with all the intermediates the derivative needs.
struct Rotated {
    double y1, z1;      // after Rx  (x1 = x)
    double x2, z2;      // after Ry  (y2 = y1)
    double x3, y3;      // after Rz  (z3 = z2)
};

SBA_HD Rotated rotate(const CamRec& c, double x, double y, double z)
{
    Rotated r;
    r.y1 = c.ca * y - c.sa * z;
    r.z1 = c.sa * y + c.ca * z;
    r.x2 = c.cb * x + c.sb * r.z1;
    r.z2 = -c.sb * x + c.cb * r.z1;
    r.x3 = c.cg * r.x2 - c.sg * r.y1;
    r.y3 = c.sg * r.x2 + c.cg * r.y1;
    return r;
}

// Ry then Rz applied to an arbitrary vector (used for d/d(alpha) and the columns of R)
SBA_HD void rot_yz(const CamRec& c, double px, double py, double pz, double& ox, double& oy, double& oz)
{
    const double x2 = c.cb * px + c.sb * pz;
    oz = -c.sb * px + c.cb * pz;
    ox = c.cg * x2 - c.sg * py;
    oy = c.sg * x2 + c.cg * py;
}

// Full rotation matrix, row-major m[3][3], same composition as `rotate`.
SBA_HD void rotation_matrix(const CamRec& c, double m[9])
{
    double x, y, z;
    rot_yz(c, 1.0, 0.0, 0.0, x, y, z);     m[0] = x; m[3] = y; m[6] = z;
    rot_yz(c, 0.0, c.ca, c.sa, x, y, z);   m[1] = x; m[4] = y; m[7] = z;
    rot_yz(c, 0.0, -c.sa, c.ca, x, y, z);  m[2] = x; m[5] = y; m[8] = z;
}

// d(rotated point)/d(alpha, beta, gamma): three column vectors
SBA_HD void rotation_derivs(const CamRec& c, const Rotated& r, double x,
                            double da[3], double db[3], double dg[3])
{
    (void)x;
    rot_yz(c, 0.0, -r.z1, r.y1, da[0], da[1], da[2]);
    db[0] = c.cg * r.z2;  db[1] = c.sg * r.z2;  db[2] = -r.x2;
    dg[0] = -r.y3;        dg[1] = r.x3;         dg[2] = 0.0;
}

// Fills a prepared camera record from the full parameter vector v (layout of ba_params.py:19-44).
SBA_HD void build_camrec(const double* v, int model, double* __restrict__ r)
{
    double sn, cs;
#ifdef __CUDA_ARCH__
    sincos(v[0], &sn, &cs); r[0] = cs; r[1] = sn;
    sincos(v[1], &sn, &cs); r[2] = cs; r[3] = sn;
    sincos(v[2], &sn, &cs); r[4] = cs; r[5] = sn;
#else
    r[0] = cos(v[0]); r[1] = sin(v[0]); r[2] = cos(v[1]); r[3] = sin(v[1]); r[4] = cos(v[2]); r[5] = sin(v[2]);
    (void)sn; (void)cs;
#endif
    if (model == MODEL_PERSPECTIVE) {
        r[6] = v[3]; r[7] = v[4]; r[8] = v[5];
        r[9] = v[6]; r[10] = v[7]; r[11] = v[8]; r[12] = v[9]; r[13] = v[10];
    } else if (model == MODEL_AFFINE) {
        r[6] = v[3]; r[7] = v[4]; r[8] = 0.0;
        r[9] = v[5]; r[10] = v[6]; r[11] = v[7]; r[12] = 0.0; r[13] = 0.0;
    } else {
        r[6] = v[3]; r[7] = v[4]; r[8] = v[5];
        r[9] = v[6]; r[10] = v[7]; r[11] = v[8]; r[12] = 0.0; r[13] = 0.0;
    }
    const CamRec c = load_camrec(r);
    double R[9];
    rotation_matrix(c, R);
#pragma unroll
    for (int k = 0; k < 9; ++k) r[CR_R + k] = R[k];
    if (model == MODEL_PERSPECTIVE) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            r[CR_KR + k] = c.k0 * R[k] + c.k2 * R[3 + k] + c.k3 * R[6 + k];
            r[CR_KR + 3 + k] = c.k1 * R[3 + k] + c.k4 * R[6 + k];
            r[CR_KR + 6 + k] = R[6 + k];
        }
        r[CR_KT] = c.k0 * c.t0 + c.k2 * c.t1 + c.k3 * c.t2;
        r[CR_KT + 1] = c.k1 * c.t1 + c.k4 * c.t2;
        r[CR_KT + 2] = c.t2;
    } else if (model == MODEL_AFFINE) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            r[CR_KR + k] = c.k0 * R[k] + c.k2 * R[3 + k];
            r[CR_KR + 3 + k] = c.k1 * R[3 + k];
            r[CR_KR + 6 + k] = 0.0;
        }
        r[CR_KT] = c.k0 * c.t0 + c.k2 * c.t1;
        r[CR_KT + 1] = c.k1 * c.t1;
        r[CR_KT + 2] = 0.0;
    } else {
#pragma unroll
        for (int k = 0; k < 12; ++k) r[CR_KR + k] = 0.0;
    }
#pragma unroll
    for (int k = CR_KT + 3; k < CAMREC_STRIDE; ++k) r[k] = 0.0;
}

// ---------------------------------------------------------------------------------------------
// geodesy + RPC polynomial with first derivatives (forward mode, 3 directions)
// ---------------------------------------------------------------------------------------------
struct D3 {   // value + gradient w.r.t. the three ECEF coordinates
    double v, d0, d1, d2;
};
SBA_HD D3 d3c(double v) { return D3{v, 0.0, 0.0, 0.0}; }
SBA_HD D3 operator+(D3 a, D3 b) { return D3{a.v + b.v, a.d0 + b.d0, a.d1 + b.d1, a.d2 + b.d2}; }
SBA_HD D3 operator-(D3 a, D3 b) { return D3{a.v - b.v, a.d0 - b.d0, a.d1 - b.d1, a.d2 - b.d2}; }
SBA_HD D3 operator*(D3 a, D3 b)
{
    return D3{a.v * b.v, a.d0 * b.v + a.v * b.d0, a.d1 * b.v + a.v * b.d1, a.d2 * b.v + a.v * b.d2};
}
SBA_HD D3 operator*(double s, D3 a) { return D3{s * a.v, s * a.d0, s * a.d1, s * a.d2}; }
SBA_HD D3 operator+(D3 a, double s) { return D3{a.v + s, a.d0, a.d1, a.d2}; }
SBA_HD D3 operator-(D3 a, double s) { return D3{a.v - s, a.d0, a.d1, a.d2}; }
SBA_HD D3 operator/(D3 a, D3 b)
{
    const double q = a.v / b.v, ib = 1.0 / b.v;
    return D3{q, (a.d0 - q * b.d0) * ib, (a.d1 - q * b.d1) * ib, (a.d2 - q * b.d2) * ib};
}
SBA_HD D3 d3_chain(double v, double dv, D3 a) { return D3{v, dv * a.d0, dv * a.d1, dv * a.d2}; }
SBA_HD D3 d3_sqrt(D3 a) { const double r = fast_rsqrt(a.v); return d3_chain(a.v * r, 0.5 * r, a); }
SBA_HD D3 d3_atan2(D3 y, D3 x)
{
    const double n = fast_rcp(x.v * x.v + y.v * y.v);
    const double gy = x.v * n, gx = -y.v * n;
    return D3{atan2(y.v, x.v), gy * y.d0 + gx * x.d0, gy * y.d1 + gx * x.d1, gy * y.d2 + gx * x.d2};
}

constexpr double WGS84_A = 6378137.0;
constexpr double REF_ECC = 8.1819190842622e-2;
constexpr double RAD2DEG = 57.295779513082320876798154814105;   // 180/pi

// value-only conversion, operation order of geo_utils.py:236-255
SBA_HD void ecef_to_geodetic(double x, double y, double z, double& lat, double& lon, double& alt)
{
    const double a = WGS84_A, asq = a * a, esq = REF_ECC * REF_ECC;
    const double b = sqrt(asq * (1.0 - esq)), bsq = b * b;
    const double ep = sqrt((asq - bsq) / bsq);
    const double p = sqrt(x * x + y * y);
    const double th = atan2(a * z, b * p);
    const double lonr = atan2(y, x);
    const double sth = sin(th), cth = cos(th);
    const double latr = atan2(z + (ep * ep) * b * (sth * sth * sth), p - esq * a * (cth * cth * cth));
    const double sl = sin(latr);
    const double N = a / sqrt(1.0 - esq * (sl * sl));
    alt = p / cos(latr) - N;
    lon = lonr * 180.0 / 3.141592653589793;
    lat = latr * 180.0 / 3.141592653589793;
}

// conversion with gradient w.r.t. (x, y, z)
SBA_HD void ecef_to_geodetic_d(double x, double y, double z, D3& lat, D3& lon, D3& alt)
{
    const double a = WGS84_A, asq = a * a, esq = REF_ECC * REF_ECC;
    const double b = sqrt(asq * (1.0 - esq)), bsq = b * b;
    const double ep2 = (asq - bsq) / bsq;
    const D3 X{x, 1.0, 0.0, 0.0}, Y{y, 0.0, 1.0, 0.0}, Z{z, 0.0, 0.0, 1.0};
    const D3 p = d3_sqrt(X * X + Y * Y);
    // sin / cos of the auxiliary angle th = atan2(a z, b p) and of the latitude follow algebraically from the arguments of
    // the arc tangents (u / hypot(u, v), v / hypot(u, v)): no atan2 + sincos round trips on this path (the value-only
    // conversion above keeps the reference's operations for the parity of `fun`; the two agree to rounding)
    const D3 tu = a * Z, tv = b * p;
    const D3 th2 = tu * tu + tv * tv;
    const double rth = fast_rsqrt(th2.v);
    const D3 irt = d3_chain(rth, -0.5 * rth * rth * rth, th2);            // 1 / hypot(tu, tv)
    const D3 sth = tu * irt, cth = tv * irt;
    const D3 s3 = sth * sth * sth, c3 = cth * cth * cth;
    const D3 lnum = Z + (ep2 * b) * s3, lden = p - (esq * a) * c3;
    const D3 latr = d3_atan2(lnum, lden);
    const D3 lonr = d3_atan2(Y, X);
    const double rl = fast_rsqrt(lnum.v * lnum.v + lden.v * lden.v);
    const double sl = lnum.v * rl, cl = lden.v * rl;
    const double iroot = fast_rsqrt(1.0 - esq * sl * sl);
    // N = a / root ; dN/dlat = a esq sl cl / root^3
    const D3 N = d3_chain(a * iroot, a * esq * sl * cl * (iroot * iroot * iroot), latr);
    const double icl = fast_rcp(cl);
    const D3 invc = d3_chain(icl, sl * icl * icl, latr);
    alt = p * invc - N;
    lon = RAD2DEG * lonr;
    lat = RAD2DEG * latr;
    // use the reference's literal scaling for the values so that they match `ecef_to_geodetic`
    lon.v = lonr.v * 180.0 / 3.141592653589793;
    lat.v = latr.v * 180.0 / 3.141592653589793;
}

// cubic polynomial, term-by-term accumulation in index order (c/rpc.c:294-297)
SBA_HD double poly20(const double* __restrict__ c, double x, double y, double z)
{
    // x = lon, y = lat, z = alt (all normalised)
    double r = 0.0;
    r += c[0];
    r += c[1] * x;
    r += c[2] * y;
    r += c[3] * z;
    r += c[4] * (x * y);
    r += c[5] * (x * z);
    r += c[6] * (y * z);
    r += c[7] * (x * x);
    r += c[8] * (y * y);
    r += c[9] * (z * z);
    r += c[10] * (y * x * z);
    r += c[11] * (x * x * x);
    r += c[12] * (x * y * y);
    r += c[13] * (x * z * z);
    r += c[14] * (x * x * y);
    r += c[15] * (y * y * y);
    r += c[16] * (y * z * z);
    r += c[17] * (x * x * z);
    r += c[18] * (y * y * z);
    r += c[19] * (z * z * z);
    return r;
}

// polynomial value and its partials w.r.t. (x, y, z)
SBA_HD void poly20_grad(const double* __restrict__ c, double x, double y, double z,
                        double& v, double& gx, double& gy, double& gz)
{
    const double xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
    v = poly20(c, x, y, z);
    gx = c[1] + c[4] * y + c[5] * z + 2.0 * c[7] * x + c[10] * yz + 3.0 * c[11] * xx + c[12] * yy + c[13] * zz +
         2.0 * c[14] * xy + 2.0 * c[17] * xz;
    gy = c[2] + c[4] * x + c[6] * z + 2.0 * c[8] * y + c[10] * xz + 2.0 * c[12] * xy + c[14] * xx +
         3.0 * c[15] * yy + c[16] * zz + 2.0 * c[18] * yz;
    gz = c[3] + c[5] * x + c[6] * y + 2.0 * c[9] * z + c[10] * xy + 2.0 * c[13] * xz + 2.0 * c[16] * yz +
         c[17] * xx + c[18] * yy + 3.0 * c[19] * zz;
}

// RPC table layout (RPC_TAB_STRIDE doubles per camera):
//   [0] row_off [1] col_off [2] lat_off [3] lon_off [4] alt_off
//   [5] row_scl [6] col_scl [7] lat_scl [8] lon_scl [9] alt_scl
//   [10..29] row_num [30..49] row_den [50..69] col_num [70..89] col_den
SBA_HD void rpc_project(const double* __restrict__ t, double lon, double lat, double alt, double& col, double& row)
{
    const double nlon = (lon - t[3]) / t[8];
    const double nlat = (lat - t[2]) / t[7];
    const double nalt = (alt - t[4]) / t[9];
    const double ncol = poly20(t + 50, nlon, nlat, nalt) / poly20(t + 70, nlon, nlat, nalt);
    const double nrow = poly20(t + 10, nlon, nlat, nalt) / poly20(t + 30, nlon, nlat, nalt);
    col = ncol * t[6] + t[1];
    row = nrow * t[5] + t[0];
}

// RPC projection of an ECEF point with the 2x3 Jacobian d(col,row)/d(x,y,z)
SBA_HD void rpc_project_ecef_d(const double* __restrict__ t, double x, double y, double z,
                               double& col, double& row, double A[6])
{
    D3 lat, lon, alt;
    ecef_to_geodetic_d(x, y, z, lat, lon, alt);
    // reciprocals of the three ground scales and of the two denominators once (13 FP64 divisions otherwise: ~25 instructions each);
    // the values differ from rpc_project's (which keeps the reference's divisions for the parity of `fun`) in the last bit only
    const double i_lon = fast_rcp(t[8]), i_lat = fast_rcp(t[7]), i_alt = fast_rcp(t[9]);
    const double nlon = (lon.v - t[3]) * i_lon;
    const double nlat = (lat.v - t[2]) * i_lat;
    const double nalt = (alt.v - t[4]) * i_alt;
    double cn, cnx, cny, cnz, cd, cdx, cdy, cdz, rn, rnx, rny, rnz, rd, rdx, rdy, rdz;
    poly20_grad(t + 50, nlon, nlat, nalt, cn, cnx, cny, cnz);
    poly20_grad(t + 70, nlon, nlat, nalt, cd, cdx, cdy, cdz);
    poly20_grad(t + 10, nlon, nlat, nalt, rn, rnx, rny, rnz);
    poly20_grad(t + 30, nlon, nlat, nalt, rd, rdx, rdy, rdz);
    const double rcd = fast_rcp(cd), rrd = fast_rcp(rd);
    const double ncol = cn * rcd, nrow = rn * rrd;
    col = ncol * t[6] + t[1];
    row = nrow * t[5] + t[0];
    // d(ncol)/d(nlon, nlat, nalt), scaled back to pixels per (deg, deg, m)
    const double icd = t[6] * rcd, ird = t[5] * rrd;
    const double c_lon = (cnx - ncol * cdx) * icd * i_lon;
    const double c_lat = (cny - ncol * cdy) * icd * i_lat;
    const double c_alt = (cnz - ncol * cdz) * icd * i_alt;
    const double r_lon = (rnx - nrow * rdx) * ird * i_lon;
    const double r_lat = (rny - nrow * rdy) * ird * i_lat;
    const double r_alt = (rnz - nrow * rdz) * ird * i_alt;
    A[0] = c_lon * lon.d0 + c_lat * lat.d0 + c_alt * alt.d0;
    A[1] = c_lon * lon.d1 + c_lat * lat.d1 + c_alt * alt.d1;
    A[2] = c_lon * lon.d2 + c_lat * lat.d2 + c_alt * alt.d2;
    A[3] = r_lon * lon.d0 + r_lat * lat.d0 + r_alt * alt.d0;
    A[4] = r_lon * lon.d1 + r_lat * lat.d1 + r_alt * alt.d1;
    A[5] = r_lon * lon.d2 + r_lat * lat.d2 + r_alt * alt.d2;
}

// ---------------------------------------------------------------------------------------------
// projection only (the residual `fun`)
// ---------------------------------------------------------------------------------------------
template <int MODEL>
SBA_HD void project(const CamRec& c, const double* __restrict__ rpc_tab, double X, double Y, double Z,
                    double& u, double& v)
{
    if (MODEL == MODEL_PERSPECTIVE) {
        const Rotated r = rotate(c, X, Y, Z);
        const double a = r.x3 + c.t0, b = r.y3 + c.t1, d = r.z2 + c.t2;
        u = (c.k0 * a + c.k2 * b + c.k3 * d) / d;
        v = (c.k1 * b + c.k4 * d) / d;
    } else if (MODEL == MODEL_AFFINE) {
        const Rotated r = rotate(c, X, Y, Z);
        const double a = r.x3 + c.t0, b = r.y3 + c.t1;
        u = c.k0 * a + c.k2 * b;
        v = c.k1 * b;
    } else {
        // X' = R (X - T - C) + C
        double qx = X - c.t0, qy = Y - c.t1, qz = Z - c.t2;
        qx -= c.k0; qy -= c.k1; qz -= c.k2;
        const Rotated r = rotate(c, qx, qy, qz);
        double lat, lon, alt;
        ecef_to_geodetic(r.x3 + c.k0, r.y3 + c.k1, r.z2 + c.k2, lat, lon, alt);
        rpc_project(rpc_tab, lon, lat, alt, u, v);
    }
}

// ---------------------------------------------------------------------------------------------
// projection + Jacobian.  NC = number of leading camera parameters that are variables
//   Jc[2][NC] : d(u,v)/d(camera parameter s)     Jp[2][3] : d(u,v)/d(point)
// (un-weighted; the caller multiplies by the observation weight and the robust scale)
// ---------------------------------------------------------------------------------------------
template <int MODEL, int NC>
SBA_HD void project_jac(const CamRec& c, const double* __restrict__ rpc_tab, double X, double Y, double Z,
                        double& u, double& v, double Jc[2 * (NC > 0 ? NC : 1)], double Jp[6])
{
    double A[6];   // d(u,v)/d(rotated-translated point), row-major 2x3
    double da[3], db[3], dg[3], R[9];
    double px = X, py = Y, pz = Z;
    if (MODEL == MODEL_RPC) {
        px = X - c.t0; py = Y - c.t1; pz = Z - c.t2;
        px -= c.k0; py -= c.k1; pz -= c.k2;
    }
    const Rotated r = rotate(c, px, py, pz);
    rotation_derivs(c, r, px, da, db, dg);
    rotation_matrix(c, R);
    double a = 0.0, b = 0.0, invd = 0.0;
    if (MODEL == MODEL_PERSPECTIVE) {
        a = r.x3 + c.t0; b = r.y3 + c.t1;
        const double d = r.z2 + c.t2;
        invd = 1.0 / d;
        u = (c.k0 * a + c.k2 * b + c.k3 * d) / d;
        v = (c.k1 * b + c.k4 * d) / d;
        A[0] = c.k0 * invd; A[1] = c.k2 * invd; A[2] = (c.k3 - u) * invd;
        A[3] = 0.0;         A[4] = c.k1 * invd; A[5] = (c.k4 - v) * invd;
    } else if (MODEL == MODEL_AFFINE) {
        a = r.x3 + c.t0; b = r.y3 + c.t1;
        u = c.k0 * a + c.k2 * b;
        v = c.k1 * b;
        A[0] = c.k0; A[1] = c.k2; A[2] = 0.0;
        A[3] = 0.0;  A[4] = c.k1; A[5] = 0.0;
    } else {
        rpc_project_ecef_d(rpc_tab, r.x3 + c.k0, r.y3 + c.k1, r.z2 + c.k2, u, v, A);
    }
    // point block: A * R   (RPC: d(X')/dX = R as well)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        Jp[k] = A[0] * R[k] + A[1] * R[3 + k] + A[2] * R[6 + k];
        Jp[3 + k] = A[3] * R[k] + A[4] * R[3 + k] + A[5] * R[6 + k];
    }
    if (NC >= 3) {
        Jc[0] = A[0] * da[0] + A[1] * da[1] + A[2] * da[2];
        Jc[1] = A[0] * db[0] + A[1] * db[1] + A[2] * db[2];
        Jc[2] = A[0] * dg[0] + A[1] * dg[1] + A[2] * dg[2];
        Jc[NC + 0] = A[3] * da[0] + A[4] * da[1] + A[5] * da[2];
        Jc[NC + 1] = A[3] * db[0] + A[4] * db[1] + A[5] * db[2];
        Jc[NC + 2] = A[3] * dg[0] + A[4] * dg[1] + A[5] * dg[2];
    }
    if (MODEL == MODEL_PERSPECTIVE) {
        if (NC >= 6) {
            Jc[3] = A[0]; Jc[4] = A[1]; Jc[5] = A[2];
            Jc[NC + 3] = A[3]; Jc[NC + 4] = A[4]; Jc[NC + 5] = A[5];
        }
        if (NC >= 11) {
            Jc[6] = a * invd; Jc[7] = 0.0;      Jc[8] = b * invd; Jc[9] = 1.0; Jc[10] = 0.0;
            Jc[NC + 6] = 0.0; Jc[NC + 7] = b * invd; Jc[NC + 8] = 0.0; Jc[NC + 9] = 0.0; Jc[NC + 10] = 1.0;
        }
    } else if (MODEL == MODEL_AFFINE) {
        if (NC >= 5) {
            Jc[3] = A[0]; Jc[4] = A[1];
            Jc[NC + 3] = A[3]; Jc[NC + 4] = A[4];
        }
        if (NC >= 8) {
            Jc[5] = a;  Jc[6] = 0.0; Jc[7] = b;
            Jc[NC + 5] = 0.0; Jc[NC + 6] = b; Jc[NC + 7] = 0.0;
        }
    } else {
        if (NC >= 6) {   // d(X')/dT = -R
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                Jc[3 + k] = -Jp[k];
                Jc[NC + 3 + k] = -Jp[3 + k];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Fast Jacobian paths used by the solver's passes (the residual `fun` keeps the reference's operation
// order above).  They use one reciprocal instead of the reference's two divisions and, for the point
// block, the per-camera product K R:   (pu, pv, d) = KR X + KT ,  u = pu/d ,
//     d(u,v)/dX = ( KR_0 - u KR_2 ; KR_1 - v KR_2 ) / d .
// Values agree with `project` to ~1e-9 px at ECEF magnitudes (same rounding level as R X + T itself).
// ---------------------------------------------------------------------------------------------
template <int MODEL>
SBA_HD void point_side(const double* __restrict__ rec, const double* __restrict__ rpc_tab, double X, double Y,
                       double Z, double& u, double& v, double Jp[6])
{
    if (MODEL == MODEL_PERSPECTIVE) {
        const double* m = rec + CR_KR;
        const double pu = m[0] * X + m[1] * Y + m[2] * Z + rec[CR_KT];
        const double pv = m[3] * X + m[4] * Y + m[5] * Z + rec[CR_KT + 1];
        const double d = m[6] * X + m[7] * Y + m[8] * Z + rec[CR_KT + 2];
        const double invd = fast_rcp(d);
        u = pu * invd; v = pv * invd;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            Jp[k] = (m[k] - u * m[6 + k]) * invd;
            Jp[3 + k] = (m[3 + k] - v * m[6 + k]) * invd;
        }
    } else if (MODEL == MODEL_AFFINE) {
        const double* m = rec + CR_KR;
        u = m[0] * X + m[1] * Y + m[2] * Z + rec[CR_KT];
        v = m[3] * X + m[4] * Y + m[5] * Z + rec[CR_KT + 1];
#pragma unroll
        for (int k = 0; k < 6; ++k) Jp[k] = m[k];
    } else {
        double dummy[2];
        const CamRec c = load_camrec(rec);
        project_jac<MODEL_RPC, 0>(c, rpc_tab, X, Y, Z, u, v, dummy, Jp);
    }
}

// Camera block and (optionally) point block.  Jc[2][NC], Jp[2][3].
template <int MODEL, int NC, bool WITH_JP>
SBA_HD void full_side(const double* __restrict__ rec, const double* __restrict__ rpc_tab, double X, double Y,
                      double Z, double& u, double& v, double Jc[2 * (NC > 0 ? NC : 1)], double Jp[6])
{
    const CamRec c = load_camrec(rec);
    if (MODEL == MODEL_RPC) {
        project_jac<MODEL_RPC, NC>(c, rpc_tab, X, Y, Z, u, v, Jc, Jp);
        return;
    }
    double A[6], da[3], db[3], dg[3];
    const Rotated r = rotate(c, X, Y, Z);
    rotation_derivs(c, r, X, da, db, dg);
    double a, b, invd = 1.0;
    if (MODEL == MODEL_PERSPECTIVE) {
        a = r.x3 + c.t0; b = r.y3 + c.t1;
        const double d = r.z2 + c.t2;
        invd = fast_rcp(d);
        u = (c.k0 * a + c.k2 * b + c.k3 * d) * invd;
        v = (c.k1 * b + c.k4 * d) * invd;
        A[0] = c.k0 * invd; A[1] = c.k2 * invd; A[2] = (c.k3 - u) * invd;
        A[3] = 0.0;         A[4] = c.k1 * invd; A[5] = (c.k4 - v) * invd;
    } else {
        a = r.x3 + c.t0; b = r.y3 + c.t1;
        u = c.k0 * a + c.k2 * b;
        v = c.k1 * b;
        A[0] = c.k0; A[1] = c.k2; A[2] = 0.0;
        A[3] = 0.0;  A[4] = c.k1; A[5] = 0.0;
    }
    if (WITH_JP) {
        const double* R = rec + CR_R;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            Jp[k] = A[0] * R[k] + A[1] * R[3 + k] + A[2] * R[6 + k];
            Jp[3 + k] = A[4] * R[3 + k] + A[5] * R[6 + k];
        }
    }
    if (NC >= 3) {
        Jc[0] = A[0] * da[0] + A[1] * da[1] + A[2] * da[2];
        Jc[1] = A[0] * db[0] + A[1] * db[1] + A[2] * db[2];
        Jc[2] = A[0] * dg[0] + A[1] * dg[1] + A[2] * dg[2];
        Jc[NC + 0] = A[4] * da[1] + A[5] * da[2];
        Jc[NC + 1] = A[4] * db[1] + A[5] * db[2];
        Jc[NC + 2] = A[4] * dg[1] + A[5] * dg[2];
    }
    if (MODEL == MODEL_PERSPECTIVE) {
        if (NC >= 6) {
            Jc[3] = A[0]; Jc[4] = A[1]; Jc[5] = A[2];
            Jc[NC + 3] = 0.0; Jc[NC + 4] = A[4]; Jc[NC + 5] = A[5];
        }
        if (NC >= 11) {
            Jc[6] = a * invd; Jc[7] = 0.0; Jc[8] = b * invd; Jc[9] = 1.0; Jc[10] = 0.0;
            Jc[NC + 6] = 0.0; Jc[NC + 7] = b * invd; Jc[NC + 8] = 0.0; Jc[NC + 9] = 0.0; Jc[NC + 10] = 1.0;
        }
    } else {
        if (NC >= 5) {
            Jc[3] = A[0]; Jc[4] = A[1];
            Jc[NC + 3] = 0.0; Jc[NC + 4] = A[4];
        }
        if (NC >= 8) {
            Jc[5] = a; Jc[6] = 0.0; Jc[7] = b;
            Jc[NC + 5] = 0.0; Jc[NC + 6] = b; Jc[NC + 7] = 0.0;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// robust loss on one residual component (scipy semantics: per component, z = (f/f_scale)^2)
//   rho0 = f_scale^2 rho(z)   (cost = 0.5 * sum rho0)
//   returns the Jacobian row scale; f is replaced by the rescaled residual
// ---------------------------------------------------------------------------------------------
SBA_HD void loss_rho(int loss, double z, double& r0, double& r1, double& r2)
{
    switch (loss) {
    case LOSS_HUBER:
        if (z <= 1.0) { r0 = z; r1 = 1.0; r2 = 0.0; }
        else { const double s = sqrt(z); r0 = 2.0 * s - 1.0; r1 = 1.0 / s; r2 = -0.5 / (z * s); }
        break;
    case LOSS_SOFT_L1: {
        const double t = 1.0 + z, s = sqrt(t);
        r0 = 2.0 * (s - 1.0); r1 = 1.0 / s; r2 = -0.5 / (t * s);
    } break;
    case LOSS_CAUCHY: {
        const double t = 1.0 + z;
        r0 = log1p(z); r1 = 1.0 / t; r2 = -1.0 / (t * t);
    } break;
    case LOSS_ARCTAN: {
        const double t = 1.0 + z * z;
        r0 = atan(z); r1 = 1.0 / t; r2 = -2.0 * z / (t * t);
    } break;
    default:
        r0 = z; r1 = 1.0; r2 = 0.0;
    }
}

// cost contribution 0.5 * f_scale^2 * rho(z) of one residual component
SBA_HD double loss_cost(int loss, double f, double f_scale)
{
    if (loss == LOSS_LINEAR) return 0.5 * f * f;
    if (loss == LOSS_SOFT_L1) {
        const double q = f * fast_rcp(f_scale);
        const double t = 1.0 + q * q;
        return f_scale * f_scale * (t * fast_rsqrt(t) - 1.0);
    }
    const double q = f / f_scale;
    double r0, r1, r2;
    loss_rho(loss, q * q, r0, r1, r2);
    return 0.5 * f_scale * f_scale * r0;
}

// Returns the row scale s; on exit f = f * rho' / s, cost = 0.5 f_scale^2 rho
SBA_HD double loss_rescale(int loss, double f_scale, double& f, double& cost)
{
    if (loss == LOSS_LINEAR) { cost = 0.5 * f * f; return 1.0; }
    if (loss == LOSS_SOFT_L1) {
        // rho' = t^-1/2, rho'' = -1/2 t^-3/2  =>  rho' + 2 rho'' z = t^-3/2 exactly (t = 1 + z):
        // scale = t^-3/4, f <- f t^1/4.  One sqrt and one reciprocal-sqrt instead of two sqrt + three divisions.
        const double q = f * fast_rcp(f_scale);
        const double t = 1.0 + q * q;
        if (t < 1e10) {                        // beyond that scipy's EPS clamp of the scale applies: generic path
            const double r2 = fast_rsqrt(t);   // t^-1/2
            const double s = t * r2;           // t^1/2
            const double r4 = fast_rsqrt(s);   // t^-1/4
            cost = f_scale * f_scale * (s - 1.0);
            f = f * (s * r4);                  // f t^1/4
            return r2 * r4;                    // t^-3/4
        }
    }
    const double q = f / f_scale;
    double r0, r1, r2;
    loss_rho(loss, q * q, r0, r1, r2);
    cost = 0.5 * f_scale * f_scale * r0;
    double js = r1 + 2.0 * (r2 / (f_scale * f_scale)) * f * f;
    if (js < 2.220446049250313e-16) js = 2.220446049250313e-16;
    js = sqrt(js);
    f = f * r1 / js;
    return js;
}

}  // namespace sba
