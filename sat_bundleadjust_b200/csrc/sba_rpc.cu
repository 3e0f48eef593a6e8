// Batched RPC kernels: projection, iterative localisation and two-view triangulation, one thread per
// point / match.  They replace, for whole arrays at once,
//   rpcm.RPCModel.projection / .localization  (third-party, call sites bundle_adjust/cam_utils.py:229,247,
//                                              bundle_adjust/ba_rpcfit.py:81,245,323,364)
//   eval_rpci / eval_rpc / rpc_height         (c/rpc.c:442-452, :378-439, :480-514)
//   stereo_corresp_to_lonlatalt               (c/disp_to_h.c:40-65, a serial loop over the matches)
// The reference algorithms are kept (same probes, same stopping rules) so that results agree to
// rounding; only the execution is different: coefficients staged once per block in shared memory,
// one match per thread, arrays of structure-of-arrays inputs read coalesced.
#include "sba_internal.cuh"
#include <algorithm>
#include <vector>

namespace sba {

// `struct rpc` of the reference (c/rpc.h:14-32) as offsets into an array of 181 doubles
constexpr int R_NUMX = 0, R_DENX = 20, R_NUMY = 40, R_DENY = 60, R_SCALE = 80, R_OFFSET = 83, R_INUMX = 86,
              R_IDENX = 106, R_INUMY = 126, R_IDENY = 146, R_ISCALE = 166, R_IOFFSET = 169, R_DELTA = 180,
              R_STRUCT_DOUBLES = 181;
constexpr int LOCALIZE_MAX_IT = 200;

// normalised (lon,lat,alt) -> normalised (col,row)      c/rpc.c:337-349 (eval_nrpci)
__device__ __forceinline__ void s_nproject(const double* r, double lon, double lat, double alt, double& x, double& y)
{
    x = poly20(r + R_INUMX, lon, lat, alt) / poly20(r + R_IDENX, lon, lat, alt);
    y = poly20(r + R_INUMY, lon, lat, alt) / poly20(r + R_IDENY, lon, lat, alt);
}

// (lon,lat,alt) -> (col,row)                               c/rpc.c:442-452 (eval_rpci)
__device__ __forceinline__ void s_project(const double* r, double lon, double lat, double alt, double& col, double& row)
{
    double x, y;
    s_nproject(r, (lon - r[R_IOFFSET]) / r[R_ISCALE], (lat - r[R_IOFFSET + 1]) / r[R_ISCALE + 1],
               (alt - r[R_IOFFSET + 2]) / r[R_ISCALE + 2], x, y);
    col = x * r[R_SCALE] + r[R_OFFSET];
    row = y * r[R_SCALE + 1] + r[R_OFFSET + 1];
}

// (col,row,alt) -> (lon,lat): direct model when present, else the iterative inversion of the projection
// by repeated affine fits (c/rpc.c:378-411): first probe at -delta with step 2 delta, then step 0.1,
// stop when the squared normalised image distance drops to 1e-18.
__device__ __forceinline__ void s_localize(const double* r, double col, double row, double alt, double& lon, double& lat)
{
    const double x = (col - r[R_OFFSET]) / r[R_SCALE];
    const double y = (row - r[R_OFFSET + 1]) / r[R_SCALE + 1];
    const double z = (alt - r[R_OFFSET + 2]) / r[R_SCALE + 2];
    double nlon, nlat;
    if (isfinite(r[R_NUMX])) {
        nlon = poly20(r + R_NUMX, x, y, z) / poly20(r + R_DENX, x, y, z);
        nlat = poly20(r + R_NUMY, x, y, z) / poly20(r + R_DENY, x, y, z);
    } else {
        const double d = r[R_DELTA] != 0.0 ? r[R_DELTA] : 1.0;
        nlon = -d; nlat = -d;
        double eps = 2.0 * d;
        double x0, y0, x1, y1, x2, y2;
        s_nproject(r, nlon, nlat, z, x0, y0);
        s_nproject(r, nlon + eps, nlat, z, x1, y1);
        s_nproject(r, nlon, nlat + eps, z, x2, y2);
        for (int it = 0; it < LOCALIZE_MAX_IT && (x0 - x) * (x0 - x) + (y0 - y) * (y0 - y) > 1e-18; ++it) {
            const double ux = x - x0, uy = y - y0;
            const double ax = x1 - x0, ay = y1 - y0, bx = x2 - x0, by = y2 - y0;
            const double det = ax * by - ay * bx;
            const double c0 = (by * ux - bx * uy) / det;
            const double c1 = (-ay * ux + ax * uy) / det;
            nlon += c0 * eps;
            nlat += c1 * eps;
            eps = 0.1;
            s_nproject(r, nlon, nlat, z, x0, y0);
            s_nproject(r, nlon + eps, nlat, z, x1, y1);
            s_nproject(r, nlon, nlat + eps, z, x2, y2);
        }
    }
    lon = nlon * r[R_ISCALE] + r[R_IOFFSET];
    lat = nlat * r[R_ISCALE + 1] + r[R_IOFFSET + 1];
}

__device__ __forceinline__ void load_struct(double* sh, const double* g, int count)
{
    for (int k = threadIdx.x; k < count; k += blockDim.x) sh[k] = g[k];
    __syncthreads();
}

__global__ void __launch_bounds__(256)
k_rpc_projection(const double* __restrict__ rs, const double* __restrict__ lon, const double* __restrict__ lat,
                 const double* __restrict__ alt, long long n, double* __restrict__ col, double* __restrict__ row)
{
    __shared__ double r[R_STRUCT_DOUBLES];
    load_struct(r, rs, R_STRUCT_DOUBLES);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        s_project(r, lon[i], lat[i], alt[i], col[i], row[i]);
}

__global__ void __launch_bounds__(256)
k_rpc_projection_ecef(const double* __restrict__ rs, const double* __restrict__ xyz, long long n,
                      double* __restrict__ colrow)
{
    __shared__ double r[R_STRUCT_DOUBLES];
    load_struct(r, rs, R_STRUCT_DOUBLES);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double la, lo, al, c, w;
        ecef_to_geodetic(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], la, lo, al);
        s_project(r, lo, la, al, c, w);
        colrow[2 * i] = c;
        colrow[2 * i + 1] = w;
    }
}

__global__ void __launch_bounds__(256)
k_rpc_localization(const double* __restrict__ rs, const double* __restrict__ col, const double* __restrict__ row,
                   const double* __restrict__ alt, long long n, double* __restrict__ lon, double* __restrict__ lat)
{
    __shared__ double r[R_STRUCT_DOUBLES];
    load_struct(r, rs, R_STRUCT_DOUBLES);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        s_localize(r, col[i], row[i], alt[i], lon[i], lat[i]);
}

// ---- initial 3-D points of the feature tracks (feature_tracks/ft_triangulate.py:18-127) ---------------------------------
// Two-view RPC triangulation of one match, c/rpc.c:480-514 + c/disp_to_h.c:50-64, result in ECEF metres
// (ft_triangulate.py:51-53 -> geo_utils.py:218-233).  ra / rb may live in global memory.
__device__ __forceinline__ void rpc_triangulate_one(const double* ra, const double* rb, double xa, double ya, double xb, double yb,
                                                    double& lon, double& lat, double& h, double& e)
{
    h = 0.0; e = 0.0;
    for (int t = 0; t < 100; ++t) {
        double lo, la, px, py, qx, qy;
        s_localize(ra, xa, ya, h, lo, la);
        s_project(rb, lo, la, h, px, py);
        s_localize(ra, xa, ya, h + 1.0, lo, la);
        s_project(rb, lo, la, h + 1.0, qx, qy);
        const double dx = qx - px, dy = qy - py, ex = xb - px, ey = yb - py;
        const double lambda = (dx * ex + dy * ey) / (dx * dx + dy * dy);
        const double zx = px + lambda * dx, zy = py + lambda * dy;
        e = hypot(zx - xb, zy - yb);
        h += lambda;
        if (fabs(lambda) < 0.00001) break;
    }
    s_localize(ra, xa, ya, h, lon, lat);
}

__device__ __forceinline__ void geodetic_to_ecef(double lat, double lon, double alt, double& x, double& y, double& z)
{
    const double phi = lat * (3.141592653589793 / 180.0), lam = lon * (3.141592653589793 / 180.0);
    const double f = 1 / 298.257223563, e2 = 1 - (1 - f) * (1 - f);
    double s, c, sl, cl;
    sincos(phi, &s, &c);
    sincos(lam, &sl, &cl);
    const double nu = 6378137.0 / sqrt(1 - e2 * s * s);
    x = (nu + alt) * c * cl;
    y = (nu + alt) * c * sl;
    z = (nu * (1 - e2) + alt) * s;
}

// Linear (DLT) triangulation of one match with two 3x4 matrices: the unit homogeneous point minimising |A X|, A = the four
// equations x P[2] - P[0], y P[2] - P[1] of both views -- what cv2.triangulatePoints computes (ft_triangulate.py:18-34).
// Right singular vector of the smallest singular value by one-sided (Hestenes) Jacobi rotations on the columns of A: the
// columns differ by 1e7 in magnitude (X, Y, Z ~ 6e6 m next to 1), and the Jacobi iteration keeps the small singular pair
// to full relative accuracy where a bidiagonalising SVD loses millimetres.
__device__ __forceinline__ void dlt_triangulate_one(const double* P1, const double* P2, double x1, double y1, double x2, double y2,
                                                    double& X, double& Y, double& Z)
{
    double U[4][4], V[4][4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        U[0][c] = x1 * P1[8 + c] - P1[c];
        U[1][c] = y1 * P1[8 + c] - P1[4 + c];
        U[2][c] = x2 * P2[8 + c] - P2[c];
        U[3][c] = y2 * P2[8 + c] - P2[4 + c];
#pragma unroll
        for (int r = 0; r < 4; ++r) V[r][c] = r == c ? 1.0 : 0.0;
    }
    const double eps = 2.220446049250313e-16;
    for (int sweep = 0; sweep < 30; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
#pragma unroll
            for (int q = p + 1; q < 4; ++q) {
                double alpha = 0.0, beta = 0.0, gamma = 0.0;
#pragma unroll
                for (int r = 0; r < 4; ++r) { alpha += U[r][p] * U[r][p]; beta += U[r][q] * U[r][q]; gamma += U[r][p] * U[r][q]; }
                if (fabs(gamma) > eps * sqrt(alpha * beta)) {
                    rotated = true;
                    const double zeta = (beta - alpha) / (2.0 * gamma);
                    const double t = zeta == 0.0 ? 1.0 : copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const double up = U[r][p], uq = U[r][q], vp = V[r][p], vq = V[r][q];
                        U[r][p] = c * up - s * uq; U[r][q] = s * up + c * uq;
                        V[r][p] = c * vp - s * vq; V[r][q] = s * vp + c * vq;
                    }
                }
            }
        }
        if (!rotated) break;
    }
    double best = 0.0, w = 1.0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        double nrm = 0.0;
#pragma unroll
        for (int r = 0; r < 4; ++r) nrm += U[r][c] * U[r][c];
        if (c == 0 || nrm < best) { best = nrm; X = V[0][c]; Y = V[1][c]; Z = V[2][c]; w = V[3][c]; }
    }
    X /= w; Y /= w; Z /= w;
}

__global__ void __launch_bounds__(128)
k_dlt_pairs(const double* __restrict__ P, const double* __restrict__ a, const double* __restrict__ b, long long n, double* __restrict__ out)
{
    __shared__ double sP[24];
    if (threadIdx.x < 24) sP[threadIdx.x] = P[threadIdx.x];
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dlt_triangulate_one(sP, sP + 12, a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1], out[3 * i], out[3 * i + 1], out[3 * i + 2]);
}

// init_pts3d (ft_triangulate.py:57-127) in one launch: one thread per track walks the list of triangulation pairs IN ORDER,
// triangulates every pair both of whose cameras see the track, and keeps the reference's float32 running mean
// (avg = ((count - 1) avg + new) / count, every operation rounded to float32, :77-81) -- the float32 quantisation of the
// start points (~0.5 m at ECEF magnitude) is part of the reference's input to bundle adjustment.  The lanes of a warp
// first advance to their next matching pair (cheap bit tests), then triangulate together (the expensive part converged).
// Observations of a track are stored with ascending cameras, one per camera, so the observation of camera c is found
// by a population count of the track's camera bit mask below c.
constexpr int TRI_MASK_WORDS = 16;                       // up to 1024 cameras
__global__ void __launch_bounds__(128)
k_init_pts3d(int rpc_model, const double* __restrict__ cams, int n_cam, const long long* __restrict__ track_ptr,
             const int* __restrict__ cam_idx, const double* __restrict__ pts2d, long long n_tracks,
             const int2* __restrict__ pairs, int n_pairs, float* __restrict__ out)
{
    const int cam_stride = rpc_model ? R_STRUCT_DOUBLES : 12;
    for (long long t0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) & ~31ll; t0 < n_tracks; t0 += (long long)gridDim.x * blockDim.x) {
        const long long t = t0 + (threadIdx.x & 31);
        const bool live = t < n_tracks;
        unsigned long long m[TRI_MASK_WORDS];
#pragma unroll
        for (int k = 0; k < TRI_MASK_WORDS; ++k) m[k] = 0ull;
        long long o0 = 0;
        if (live) {
            o0 = track_ptr[t];
            const long long o1 = track_ptr[t + 1];
            for (long long o = o0; o < o1; ++o) { const int c = cam_idx[o]; m[c >> 6] |= 1ull << (c & 63); }
        }
        float ax = 0.f, ay = 0.f, az = 0.f, cnt = 0.f;
        int q = live ? 0 : n_pairs;
        while (true) {
            int ci = -1, cj = -1;
            for (; q < n_pairs; ++q) {
                const int2 pr = pairs[q];
                if (pr.x >= 0 && pr.y >= 0 && pr.x < n_cam && pr.y < n_cam && ((m[pr.x >> 6] >> (pr.x & 63)) & 1ull) &&
                    ((m[pr.y >> 6] >> (pr.y & 63)) & 1ull)) { ci = pr.x; cj = pr.y; ++q; break; }
            }
            if (__ballot_sync(0xffffffffu, ci >= 0) == 0u) break;
            if (ci >= 0) {
                int pi = 0, pj = 0;
                for (int k = 0; k < TRI_MASK_WORDS; ++k) {
                    if (k < (ci >> 6)) pi += __popcll(m[k]);
                    if (k < (cj >> 6)) pj += __popcll(m[k]);
                }
                pi += __popcll(m[ci >> 6] & ((1ull << (ci & 63)) - 1ull));
                pj += __popcll(m[cj >> 6] & ((1ull << (cj & 63)) - 1ull));
                const double xi = pts2d[2 * (o0 + pi)], yi = pts2d[2 * (o0 + pi) + 1];
                const double xj = pts2d[2 * (o0 + pj)], yj = pts2d[2 * (o0 + pj) + 1];
                const double* ca = cams + (size_t)ci * cam_stride;
                const double* cb = cams + (size_t)cj * cam_stride;
                double X, Y, Z;
                if (rpc_model) {
                    double lon, lat, h, e;   // the reference's binding casts the keypoints to float32 (s2p/triangulation.py)
                    rpc_triangulate_one(ca, cb, (double)(float)xi, (double)(float)yi, (double)(float)xj, (double)(float)yj, lon, lat, h, e);
                    geodetic_to_ecef(lat, lon, h, X, Y, Z);
                } else {
                    dlt_triangulate_one(ca, cb, xi, yi, xj, yj, X, Y, Z);
                }
                cnt = __fadd_rn(cnt, 1.f);
                const float c1 = __fsub_rn(cnt, 1.f);
                ax = __fdiv_rn(__fadd_rn(__fmul_rn(c1, ax), (float)X), cnt);
                ay = __fdiv_rn(__fadd_rn(__fmul_rn(c1, ay), (float)Y), cnt);
                az = __fdiv_rn(__fadd_rn(__fmul_rn(c1, az), (float)Z), cnt);
            }
        }
        if (live) { out[3 * t] = ax; out[3 * t + 1] = ay; out[3 * t + 2] = az; }
    }
}

// The same for many cameras in ONE launch: blockIdx.y = camera, whose coefficients the block stages in shared memory; the
// points are shared by all cameras (in_stride == 0, e.g. one lon/lat/alt grid) or given per camera (in_stride == n).
// kind 0: (a, b, c) = (lon, lat, alt) -> (col, row);  kind 1: (a, b, c) = (col, row, alt) -> (lon, lat)
template <int KIND>
__global__ void __launch_bounds__(256)
k_rpc_batch(const double* __restrict__ rs, const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ c,
            long long n, long long in_stride, double* __restrict__ o0, double* __restrict__ o1)
{
    __shared__ double r[R_STRUCT_DOUBLES];
    load_struct(r, rs + (size_t)blockIdx.y * R_STRUCT_DOUBLES, R_STRUCT_DOUBLES);
    const size_t in0 = (size_t)blockIdx.y * in_stride, out0 = (size_t)blockIdx.y * n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (KIND == 0) s_project(r, a[in0 + i], b[in0 + i], c[in0 + i], o0[out0 + i], o1[out0 + i]);
        else s_localize(r, a[in0 + i], b[in0 + i], c[in0 + i], o0[out0 + i], o1[out0 + i]);
    }
}

// c/rpc.c:480-514 (rpc_height) + c/disp_to_h.c:40-65, one match per thread
__global__ void __launch_bounds__(128)
k_rpc_triangulate(const double* __restrict__ rsa, const double* __restrict__ rsb, const float2* __restrict__ kp_a,
                  const float2* __restrict__ kp_b, long long n, double* __restrict__ lonlatalt, float* __restrict__ err)
{
    __shared__ double ra[R_STRUCT_DOUBLES], rb[R_STRUCT_DOUBLES];
    for (int k = threadIdx.x; k < R_STRUCT_DOUBLES; k += blockDim.x) { ra[k] = rsa[k]; rb[k] = rsb[k]; }
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float2 a = kp_a[i], b = kp_b[i];
        const double xa = a.x, ya = a.y, xb = b.x, yb = b.y;
        double lo, la, h, e;
        rpc_triangulate_one(ra, rb, xa, ya, xb, yb, lo, la, h, e);
        lonlatalt[3 * i] = lo;
        lonlatalt[3 * i + 1] = la;
        lonlatalt[3 * i + 2] = h;
        err[i] = (float)e;
    }
}

// 90-double table (sba_b200.h layout) -> 181-double struct with the direct model marked absent
static void table_to_struct(const double* t, double delta, double* s)
{
    for (int k = 0; k < R_STRUCT_DOUBLES; ++k) s[k] = 0.0;
    for (int k = 0; k < 80; ++k) s[k] = NAN;
    s[R_OFFSET] = t[1]; s[R_OFFSET + 1] = t[0]; s[R_OFFSET + 2] = t[4];
    s[R_SCALE] = t[6]; s[R_SCALE + 1] = t[5]; s[R_SCALE + 2] = t[9];
    s[R_IOFFSET] = t[3]; s[R_IOFFSET + 1] = t[2]; s[R_IOFFSET + 2] = t[4];
    s[R_ISCALE] = t[8]; s[R_ISCALE + 1] = t[7]; s[R_ISCALE + 2] = t[9];
    for (int k = 0; k < 20; ++k) {
        s[R_INUMX + k] = t[50 + k]; s[R_IDENX + k] = t[70 + k];
        s[R_INUMY + k] = t[10 + k]; s[R_IDENY + k] = t[30 + k];
    }
    s[R_DELTA] = delta;
}

struct DevBuf {   // small RAII helper for the per-call staging buffers
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) { SBA_CUDA(cudaMalloc(&p, bytes ? bytes : 1)); return SBA_OK; }
    template <typename T> T* as() { return (T*)p; }
};

static int require_device()
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("no CUDA device: sat_bundleadjust_b200 has no CPU fallback");
        return SBA_E_CUDA;
    }
    return SBA_OK;
}

static inline int grid_n(long long n, int threads)
{
    long long b = (n + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > NUM_SMS * 16) b = NUM_SMS * 16;
    return (int)b;
}

int launch_cholesky_solve(double* A_dev, double* b_dev, double* x_dev, int n, double* fail_dev, double* work_dev,
                          cudaStream_t stream, bool write_factor);
int chol_stage_clocks(long long* out16);

}  // namespace sba

using namespace sba;

extern "C" int sba_rpc_projection(const double* rpc, const double* lon, const double* lat, const double* alt, int64_t n,
                                  double* col, double* row)
{
    if (!rpc || !lon || !lat || !alt || !col || !row || n < 0) { set_error("bad argument"); return SBA_E_INVALID; }
    if (n == 0) return SBA_OK;
    SBA_TRY(require_device());
    double s[R_STRUCT_DOUBLES];
    table_to_struct(rpc, 1.0, s);
    DevBuf ds, in, out;
    SBA_TRY(ds.alloc(sizeof(s))); SBA_TRY(in.alloc(3 * n * sizeof(double))); SBA_TRY(out.alloc(2 * n * sizeof(double)));
    SBA_CUDA(cudaMemcpy(ds.p, s, sizeof(s), cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemcpy(in.as<double>(), lon, n * sizeof(double), cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemcpy(in.as<double>() + n, lat, n * sizeof(double), cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemcpy(in.as<double>() + 2 * n, alt, n * sizeof(double), cudaMemcpyHostToDevice));
    k_rpc_projection<<<grid_n(n, 256), 256>>>(ds.as<double>(), in.as<double>(), in.as<double>() + n, in.as<double>() + 2 * n,
                                             n, out.as<double>(), out.as<double>() + n);
    SBA_CUDA(cudaGetLastError());
    SBA_CUDA(cudaMemcpy(col, out.as<double>(), n * sizeof(double), cudaMemcpyDeviceToHost));
    SBA_CUDA(cudaMemcpy(row, out.as<double>() + n, n * sizeof(double), cudaMemcpyDeviceToHost));
    return SBA_OK;
}

extern "C" int sba_rpc_projection_ecef(const double* rpc, const double* xyz, int64_t n, double* colrow)
{
    if (!rpc || !xyz || !colrow || n < 0) { set_error("bad argument"); return SBA_E_INVALID; }
    if (n == 0) return SBA_OK;
    SBA_TRY(require_device());
    double s[R_STRUCT_DOUBLES];
    table_to_struct(rpc, 1.0, s);
    DevBuf ds, in, out;
    SBA_TRY(ds.alloc(sizeof(s))); SBA_TRY(in.alloc(3 * n * sizeof(double))); SBA_TRY(out.alloc(2 * n * sizeof(double)));
    SBA_CUDA(cudaMemcpy(ds.p, s, sizeof(s), cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemcpy(in.p, xyz, 3 * n * sizeof(double), cudaMemcpyHostToDevice));
    k_rpc_projection_ecef<<<grid_n(n, 256), 256>>>(ds.as<double>(), in.as<double>(), n, out.as<double>());
    SBA_CUDA(cudaGetLastError());
    SBA_CUDA(cudaMemcpy(colrow, out.p, 2 * n * sizeof(double), cudaMemcpyDeviceToHost));
    return SBA_OK;
}

extern "C" int sba_rpc_localization(const double* rpc, const double* col, const double* row, const double* alt, int64_t n,
                                    double delta, double* lon, double* lat)
{
    if (!rpc || !col || !row || !alt || !lon || !lat || n < 0) { set_error("bad argument"); return SBA_E_INVALID; }
    if (n == 0) return SBA_OK;
    SBA_TRY(require_device());
    double s[R_STRUCT_DOUBLES];
    table_to_struct(rpc, delta, s);
    DevBuf ds, in, out;
    SBA_TRY(ds.alloc(sizeof(s))); SBA_TRY(in.alloc(3 * n * sizeof(double))); SBA_TRY(out.alloc(2 * n * sizeof(double)));
    SBA_CUDA(cudaMemcpy(ds.p, s, sizeof(s), cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemcpy(in.as<double>(), col, n * sizeof(double), cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemcpy(in.as<double>() + n, row, n * sizeof(double), cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemcpy(in.as<double>() + 2 * n, alt, n * sizeof(double), cudaMemcpyHostToDevice));
    k_rpc_localization<<<grid_n(n, 256), 256>>>(ds.as<double>(), in.as<double>(), in.as<double>() + n,
                                               in.as<double>() + 2 * n, n, out.as<double>(), out.as<double>() + n);
    SBA_CUDA(cudaGetLastError());
    SBA_CUDA(cudaMemcpy(lon, out.as<double>(), n * sizeof(double), cudaMemcpyDeviceToHost));
    SBA_CUDA(cudaMemcpy(lat, out.as<double>() + n, n * sizeof(double), cudaMemcpyDeviceToHost));
    return SBA_OK;
}

extern "C" int sba_stereo_corresp_to_lonlatalt(double* lonlatalt, float* err, const float* kp_a, const float* kp_b,
                                               int64_t n, const void* rpc_a, const void* rpc_b)
{
    if (!lonlatalt || !err || !kp_a || !kp_b || !rpc_a || !rpc_b || n < 0) { set_error("bad argument"); return SBA_E_INVALID; }
    if (n == 0) return SBA_OK;
    SBA_TRY(require_device());
    DevBuf sa, sb, ka, kb, out, e;
    const size_t sbytes = R_STRUCT_DOUBLES * sizeof(double);
    SBA_TRY(sa.alloc(sbytes)); SBA_TRY(sb.alloc(sbytes));
    SBA_TRY(ka.alloc(2 * n * sizeof(float))); SBA_TRY(kb.alloc(2 * n * sizeof(float)));
    SBA_TRY(out.alloc(3 * n * sizeof(double))); SBA_TRY(e.alloc(n * sizeof(float)));
    SBA_CUDA(cudaMemcpy(sa.p, rpc_a, sbytes, cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemcpy(sb.p, rpc_b, sbytes, cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemcpy(ka.p, kp_a, 2 * n * sizeof(float), cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemcpy(kb.p, kp_b, 2 * n * sizeof(float), cudaMemcpyHostToDevice));
    k_rpc_triangulate<<<grid_n(n, 128), 128>>>(sa.as<double>(), sb.as<double>(), ka.as<float2>(), kb.as<float2>(), n,
                                              out.as<double>(), e.as<float>());
    SBA_CUDA(cudaGetLastError());
    SBA_CUDA(cudaMemcpy(lonlatalt, out.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost));
    SBA_CUDA(cudaMemcpy(err, e.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    return SBA_OK;
}

// the reference's symbol, same signature (returns void; errors are reported on stderr, never exit())
extern "C" void stereo_corresp_to_lonlatalt(double* lonlatalt, float* err, float* kp_a, float* kp_b, int n_kp,
                                            void* rpc_a, void* rpc_b)
{
    const int rc = sba_stereo_corresp_to_lonlatalt(lonlatalt, err, kp_a, kp_b, n_kp, rpc_a, rpc_b);
    if (rc != SBA_OK) {
        // the reference's signature has no way to report failure: never leave plausible-looking zeros behind
        fprintf(stderr, "stereo_corresp_to_lonlatalt (sba_b200): %s\n", sba_last_error());
        if (lonlatalt && err)
            for (int i = 0; i < n_kp; ++i) { lonlatalt[3 * i] = lonlatalt[3 * i + 1] = lonlatalt[3 * i + 2] = NAN; err[i] = NAN; }
    }
}

// Linear triangulation of n matches between two 3x4 matrices (ft_triangulate.py:18-34, the reference calls cv2.triangulatePoints)
extern "C" int sba_linear_triangulation(const double* P1, const double* P2, const double* pts1, const double* pts2, int64_t n,
                                        double* pts3d)
{
    if (!P1 || !P2 || !pts1 || !pts2 || !pts3d || n < 0) { set_error("bad argument"); return SBA_E_INVALID; }
    if (n == 0) return SBA_OK;
    SBA_TRY(require_device());
    DevBuf dp, in, out;
    SBA_TRY(dp.alloc(24 * sizeof(double))); SBA_TRY(in.alloc(4 * n * sizeof(double))); SBA_TRY(out.alloc(3 * n * sizeof(double)));
    SBA_CUDA(cudaMemcpy(dp.as<double>(), P1, 12 * sizeof(double), cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemcpy(dp.as<double>() + 12, P2, 12 * sizeof(double), cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemcpy(in.as<double>(), pts1, 2 * n * sizeof(double), cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemcpy(in.as<double>() + 2 * n, pts2, 2 * n * sizeof(double), cudaMemcpyHostToDevice));
    k_dlt_pairs<<<grid_n(n, 128), 128>>>(dp.as<double>(), in.as<double>(), in.as<double>() + 2 * n, n, out.as<double>());
    SBA_CUDA(cudaGetLastError());
    SBA_CUDA(cudaMemcpy(pts3d, out.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost));
    return SBA_OK;
}

// init_pts3d (ft_triangulate.py:57-127) for all tracks and all triangulation pairs in one launch.  Tracks in CSR form
// (track_ptr, cam_idx ascending within a track, pts2d); cameras: n_cam x 12 doubles (row-major 3x4) for cam_model
// affine / perspective, n_cam x 181 doubles (`struct rpc`, c/rpc.h:14-32) for cam_model rpc.
extern "C" int sba_init_pts3d(int32_t cam_model, const double* cams, int32_t n_cam, const int64_t* track_ptr, const int32_t* cam_idx,
                              const double* pts2d, int64_t n_tracks, const int32_t* pairs, int32_t n_pairs, float* pts3d)
{
    if (!cams || !track_ptr || !pts3d || n_tracks < 0 || n_cam < 1 || n_pairs < 0 || (n_pairs > 0 && !pairs) ||
        (cam_model != MODEL_AFFINE && cam_model != MODEL_PERSPECTIVE && cam_model != MODEL_RPC)) {
        set_error("bad argument"); return SBA_E_INVALID;
    }
    if (n_cam > 64 * TRI_MASK_WORDS) { set_error("unsupported size: init_pts3d handles up to 1024 cameras"); return SBA_E_INVALID; }
    if (n_tracks == 0) return SBA_OK;
    const int64_t K = track_ptr[n_tracks];
    if (K < 0 || (K > 0 && (!cam_idx || !pts2d))) { set_error("bad argument"); return SBA_E_INVALID; }
    for (int64_t t = 0; t < n_tracks; ++t)
        for (int64_t o = track_ptr[t]; o < track_ptr[t + 1]; ++o)
            if (cam_idx[o] < 0 || cam_idx[o] >= n_cam || (o > track_ptr[t] && cam_idx[o] <= cam_idx[o - 1])) {
                set_error("init_pts3d: cameras of a track must be ascending and below n_cam"); return SBA_E_INVALID;
            }
    SBA_TRY(require_device());
    const int rpc_model = cam_model == MODEL_RPC;
    const size_t cam_doubles = (size_t)n_cam * (rpc_model ? R_STRUCT_DOUBLES : 12);
    DevBuf dc, dt, di, dp, dq, out;
    SBA_TRY(dc.alloc(cam_doubles * sizeof(double))); SBA_TRY(dt.alloc((n_tracks + 1) * sizeof(int64_t)));
    SBA_TRY(di.alloc(K * sizeof(int32_t))); SBA_TRY(dp.alloc(2 * K * sizeof(double)));
    SBA_TRY(dq.alloc((size_t)n_pairs * 2 * sizeof(int32_t))); SBA_TRY(out.alloc(3 * n_tracks * sizeof(float)));
    SBA_CUDA(cudaMemcpy(dc.p, cams, cam_doubles * sizeof(double), cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemcpy(dt.p, track_ptr, (n_tracks + 1) * sizeof(int64_t), cudaMemcpyHostToDevice));
    if (K > 0) {
        SBA_CUDA(cudaMemcpy(di.p, cam_idx, K * sizeof(int32_t), cudaMemcpyHostToDevice));
        SBA_CUDA(cudaMemcpy(dp.p, pts2d, 2 * K * sizeof(double), cudaMemcpyHostToDevice));
    }
    if (n_pairs > 0) SBA_CUDA(cudaMemcpy(dq.p, pairs, (size_t)n_pairs * 2 * sizeof(int32_t), cudaMemcpyHostToDevice));
    k_init_pts3d<<<grid_n(n_tracks, 128), 128>>>(rpc_model, dc.as<double>(), n_cam, dt.as<long long>(), di.as<int>(), dp.as<double>(),
                                                n_tracks, dq.as<int2>(), n_pairs, out.as<float>());
    SBA_CUDA(cudaGetLastError());
    SBA_CUDA(cudaMemcpy(pts3d, out.p, 3 * n_tracks * sizeof(float), cudaMemcpyDeviceToHost));
    return SBA_OK;
}

// Projection / localisation for n_cam cameras in one launch (BASELINE config 5: a 100 x 100 x 10 grid for each of 300 cameras).
// tables: (n_cam, 90) in the layout of sba_b200.h; a, b, c: n values each when shared != 0 (the same points for every camera),
// else (n_cam, n); outputs (n_cam, n) each.  kind 0: (lon, lat, alt) -> (col, row); kind 1: (col, row, alt) -> (lon, lat).
static int rpc_batch(int kind, const double* tables, int32_t n_cam, const double* a, const double* b, const double* c, int64_t n,
                     int32_t shared, double delta, double* o0, double* o1)
{
    if (!tables || !a || !b || !c || !o0 || !o1 || n < 0 || n_cam < 1 || n_cam > 65535) { set_error("bad argument"); return SBA_E_INVALID; }
    if (n == 0) return SBA_OK;
    SBA_TRY(require_device());
    std::vector<double> hs((size_t)n_cam * R_STRUCT_DOUBLES);
    for (int j = 0; j < n_cam; ++j) table_to_struct(tables + (size_t)j * 90, delta, hs.data() + (size_t)j * R_STRUCT_DOUBLES);
    const size_t n_in = (size_t)(shared ? 1 : n_cam) * n, n_out = (size_t)n_cam * n;
    DevBuf ds, in, out;
    SBA_TRY(ds.alloc(hs.size() * sizeof(double))); SBA_TRY(in.alloc(3 * n_in * sizeof(double))); SBA_TRY(out.alloc(2 * n_out * sizeof(double)));
    SBA_CUDA(cudaMemcpy(ds.p, hs.data(), hs.size() * sizeof(double), cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemcpy(in.as<double>(), a, n_in * sizeof(double), cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemcpy(in.as<double>() + n_in, b, n_in * sizeof(double), cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemcpy(in.as<double>() + 2 * n_in, c, n_in * sizeof(double), cudaMemcpyHostToDevice));
    const dim3 grid((unsigned)std::min<long long>((n + 255) / 256, 4 * NUM_SMS), (unsigned)n_cam);
    const long long stride = shared ? 0 : n;
    if (kind == 0)
        k_rpc_batch<0><<<grid, 256>>>(ds.as<double>(), in.as<double>(), in.as<double>() + n_in, in.as<double>() + 2 * n_in, n, stride,
                                      out.as<double>(), out.as<double>() + n_out);
    else
        k_rpc_batch<1><<<grid, 256>>>(ds.as<double>(), in.as<double>(), in.as<double>() + n_in, in.as<double>() + 2 * n_in, n, stride,
                                      out.as<double>(), out.as<double>() + n_out);
    SBA_CUDA(cudaGetLastError());
    SBA_CUDA(cudaMemcpy(o0, out.as<double>(), n_out * sizeof(double), cudaMemcpyDeviceToHost));
    SBA_CUDA(cudaMemcpy(o1, out.as<double>() + n_out, n_out * sizeof(double), cudaMemcpyDeviceToHost));
    return SBA_OK;
}

extern "C" int sba_rpc_projection_batch(const double* tables, int32_t n_cam, const double* lon, const double* lat, const double* alt,
                                        int64_t n, int32_t shared_points, double* col, double* row)
{
    return rpc_batch(0, tables, n_cam, lon, lat, alt, n, shared_points, 1.0, col, row);
}

extern "C" int sba_rpc_localization_batch(const double* tables, int32_t n_cam, const double* col, const double* row, const double* alt,
                                          int64_t n, int32_t shared_points, double delta, double* lon, double* lat)
{
    return rpc_batch(1, tables, n_cam, col, row, alt, n, shared_points, delta, lon, lat);
}

// Measurement entry point (bench.py --workload rpc): device-resident throughput of the batched RPC kernels over n_cam cameras
// x n points, inputs uploaded once, `reps` timed repetitions of one launch per camera (CUDA events).
//   kind 0: projection  (a, b, c) = (lon, lat, alt) -> (col, row)
//   kind 1: localisation (a, b, c) = (col, row, alt) -> (lon, lat)
//   kind 2: triangulation of n matches between tables 2j and 2j+1 (kp_a = (a, b), kp_b = (c, d) as doubles, cast to float)
// out (2n or 3n doubles, may be NULL) receives the result of the last camera / pair for a spot check.
extern "C" int sba_rpc_throughput(int32_t kind, const double* tables, int32_t n_cam, const double* a, const double* b,
                                  const double* c, const double* d, int64_t n, double delta, int32_t reps, double* out, double* ms)
{
    if (!tables || !a || !b || !c || n < 1 || n_cam < 1 || reps < 1 || !ms || kind < 0 || kind > 2) { set_error("bad argument"); return SBA_E_INVALID; }
    SBA_TRY(require_device());
    const int n_str = kind == 2 ? 2 * n_cam : n_cam;
    std::vector<double> hs((size_t)n_str * R_STRUCT_DOUBLES);
    for (int j = 0; j < n_str; ++j) table_to_struct(tables + (size_t)j * 90, delta, hs.data() + (size_t)j * R_STRUCT_DOUBLES);
    DevBuf ds, in, o, kf, ef;
    SBA_TRY(ds.alloc(hs.size() * sizeof(double))); SBA_TRY(in.alloc(4 * n * sizeof(double))); SBA_TRY(o.alloc(3 * n * sizeof(double)));
    SBA_CUDA(cudaMemcpy(ds.p, hs.data(), hs.size() * sizeof(double), cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemcpy(in.as<double>(), a, n * sizeof(double), cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemcpy(in.as<double>() + n, b, n * sizeof(double), cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemcpy(in.as<double>() + 2 * n, c, n * sizeof(double), cudaMemcpyHostToDevice));
    if (kind == 2) {
        if (!d) { set_error("bad argument"); return SBA_E_INVALID; }
        std::vector<float> kp(4 * (size_t)n);
        for (int64_t i = 0; i < n; ++i) { kp[2 * i] = (float)a[i]; kp[2 * i + 1] = (float)b[i]; kp[2 * n + 2 * i] = (float)c[i]; kp[2 * n + 2 * i + 1] = (float)d[i]; }
        SBA_TRY(kf.alloc(kp.size() * sizeof(float))); SBA_TRY(ef.alloc(n * sizeof(float)));
        SBA_CUDA(cudaMemcpy(kf.p, kp.data(), kp.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    cudaEvent_t e0, e1;
    SBA_CUDA(cudaEventCreate(&e0)); SBA_CUDA(cudaEventCreate(&e1));
    // kinds 0 and 1: ONE launch for all cameras (blockIdx.y = camera, the same points for every camera); the output of the last
    // camera is what `out` receives.  kind 2: one launch per pair.
    DevBuf big;
    if (kind != 2) SBA_TRY(big.alloc((size_t)2 * n_cam * n * sizeof(double)));
    auto pass = [&]() {
        double* x = in.as<double>();
        const dim3 grid((unsigned)std::min<long long>((n + 255) / 256, 4 * NUM_SMS), (unsigned)n_cam);
        if (kind == 0) k_rpc_batch<0><<<grid, 256>>>(ds.as<double>(), x, x + n, x + 2 * n, n, 0, big.as<double>(), big.as<double>() + (size_t)n_cam * n);
        else if (kind == 1) k_rpc_batch<1><<<grid, 256>>>(ds.as<double>(), x, x + n, x + 2 * n, n, 0, big.as<double>(), big.as<double>() + (size_t)n_cam * n);
        else
            for (int j = 0; j < n_cam; ++j) {
                const double* sj = ds.as<double>() + (size_t)2 * j * R_STRUCT_DOUBLES;
                k_rpc_triangulate<<<grid_n(n, 128), 128>>>(sj, sj + R_STRUCT_DOUBLES, kf.as<float2>(), kf.as<float2>() + n, n, o.as<double>(), ef.as<float>());
            }
    };
    pass();                                     // warm-up
    SBA_CUDA(cudaDeviceSynchronize());
    SBA_CUDA(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) pass();
    SBA_CUDA(cudaEventRecord(e1));
    SBA_CUDA(cudaEventSynchronize(e1));
    float t = 0.f;
    SBA_CUDA(cudaEventElapsedTime(&t, e0, e1));
    SBA_CUDA(cudaGetLastError());
    *ms = (double)t / reps;
    if (out && kind == 2) SBA_CUDA(cudaMemcpy(out, o.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost));
    if (out && kind != 2) {      // last camera: [first output (n) | second output (n)]
        SBA_CUDA(cudaMemcpy(out, big.as<double>() + (size_t)(n_cam - 1) * n, n * sizeof(double), cudaMemcpyDeviceToHost));
        SBA_CUDA(cudaMemcpy(out + n, big.as<double>() + (size_t)n_cam * n + (size_t)(n_cam - 1) * n, n * sizeof(double), cudaMemcpyDeviceToHost));
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return SBA_OK;
}

extern "C" int sba_cholesky_solve(double* A, double* b, int32_t n, int32_t* info)
{
    if (!A || !b || n < 1) { set_error("bad argument"); return SBA_E_INVALID; }
    SBA_TRY(require_device());
    DevBuf dA, db, dx, df, dw;
    SBA_TRY(dw.alloc((size_t)34 * (n + 32) * sizeof(double)));
    SBA_TRY(dA.alloc((size_t)n * n * sizeof(double))); SBA_TRY(db.alloc(n * sizeof(double)));
    SBA_TRY(dx.alloc(n * sizeof(double))); SBA_TRY(df.alloc(sizeof(double)));
    SBA_CUDA(cudaMemcpy(dA.p, A, (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemcpy(db.p, b, n * sizeof(double), cudaMemcpyHostToDevice));
    SBA_TRY(launch_cholesky_solve(dA.as<double>(), db.as<double>(), dx.as<double>(), n, df.as<double>(), dw.as<double>(), 0, true));
    double fail = 0.0;
    SBA_CUDA(cudaMemcpy(&fail, df.p, sizeof(double), cudaMemcpyDeviceToHost));
    SBA_CUDA(cudaMemcpy(A, dA.p, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost));
    SBA_CUDA(cudaMemcpy(b, dx.p, n * sizeof(double), cudaMemcpyDeviceToHost));
    if (info) *info = (int32_t)fail;
    return SBA_OK;
}

// the same solve `reps` times on device-resident data (A restored from a pristine copy before each repetition,
// outside the event pair); *ms receives the mean device time of one factor + solve
extern "C" int sba_cholesky_solve_timed(const double* A, const double* b, int32_t n, int32_t reps, double* x, double* ms)
{
    if (!A || !b || !x || !ms || n < 1 || reps < 1) { set_error("bad argument"); return SBA_E_INVALID; }
    SBA_TRY(require_device());
    DevBuf dA0, dA, db, dx, df, dw;
    const size_t bytes = (size_t)n * n * sizeof(double);
    SBA_TRY(dw.alloc((size_t)34 * (n + 32) * sizeof(double)));
    SBA_TRY(dA0.alloc(bytes)); SBA_TRY(dA.alloc(bytes)); SBA_TRY(db.alloc(n * sizeof(double)));
    SBA_TRY(dx.alloc(n * sizeof(double))); SBA_TRY(df.alloc(sizeof(double)));
    SBA_CUDA(cudaMemcpy(dA0.p, A, bytes, cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemcpy(db.p, b, n * sizeof(double), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    SBA_CUDA(cudaEventCreate(&e0)); SBA_CUDA(cudaEventCreate(&e1));
    double total = 0.0;
    for (int r = -2; r < reps; ++r) {            // two warm-up repetitions
        SBA_CUDA(cudaMemcpyAsync(dA.p, dA0.p, bytes, cudaMemcpyDeviceToDevice, 0));
        SBA_CUDA(cudaEventRecord(e0, 0));
        SBA_TRY(launch_cholesky_solve(dA.as<double>(), db.as<double>(), dx.as<double>(), n, df.as<double>(), dw.as<double>(), 0, false));
        SBA_CUDA(cudaEventRecord(e1, 0));
        SBA_CUDA(cudaEventSynchronize(e1));
        float t = 0.f;
        SBA_CUDA(cudaEventElapsedTime(&t, e0, e1));
        if (r >= 0) total += t;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    SBA_CUDA(cudaMemcpy(x, dx.p, n * sizeof(double), cudaMemcpyDeviceToHost));
    if (getenv("SBA_CHOL_CLK") && n <= 127) {      // stage clocks of the one-CTA kernel: load | per panel: block, rows, trailing | back-substitution
        long long c[16];
        SBA_TRY(chol_stage_clocks(c));
        fprintf(stderr, "chol n=%d cycles: load %lld | p0 %lld %lld %lld | p1 %lld %lld %lld | backsub %lld | total %lld\n", n, c[1] - c[0],
                c[2] - c[1], c[3] - c[2], c[4] - c[3], c[5] - c[4], c[6] - c[5], c[7] - c[6], c[8] - c[7], c[8] - c[0]);
    }
    *ms = total / reps;
    return SBA_OK;
}
