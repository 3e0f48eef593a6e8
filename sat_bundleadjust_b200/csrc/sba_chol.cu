// FP64 dense Cholesky factorisation + triangular solves of the reduced camera system S dc = rhs
// (n = n_cam * n_params <= ~1800).  There is no reference counterpart: the reference never forms S,
// it hands the full sparse Jacobian to LSMR (scipy/optimize/_lsq/trf.py:485-495).
//
// One CTA, left-looking, one barrier per column (see the kernel).  The right-hand side rides along as an extra
// row of the matrix, so the forward substitution L y = rhs falls out of the factorisation itself and only the
// backward substitution remains.  S is staged in shared memory when it fits (n <= 160: 161 x 161 doubles =
// 207 KB of the 227 KB a CTA may use), otherwise worked on in a global scratch buffer, which is L2-resident.
#include "sba_internal.cuh"

namespace sba {

constexpr int CHOL_THREADS = 256;
constexpr int CHOL_SMEM_MAX_N = 160;

// Left-looking, one thread per row, ONE barrier per column: thread i forms
//     L[i,k] = (A[i,k] - sum_{m<k} L[i,m] L[k,m]) * rsqrt(A[k,k] - sum_{m<k} L[k,m]^2)
// from columns that are already final (every thread recomputes the pivot instead of waiting for it).
// Rows are stored with an odd leading dimension, so a column access is bank-conflict free and the pivot row is
// a broadcast.  The right-hand side is row n: after the factorisation it holds y = L^-1 rhs.
// A: n x n column-major, lower triangle used, overwritten by L.  b: rhs (n).  x: solution (n).
// fail: (k+1) when pivot k is not positive / finite, else 0.
template <bool SMEM>
__global__ void __launch_bounds__(CHOL_THREADS)
k_cholesky_solve(double* Ag, double* b, double* x, int n, double* fail, double* work)
{
    extern __shared__ double sh[];
    const int ld = n | 1;                          // odd row stride
    double* L = SMEM ? sh : work;                  // (n+1) rows x ld, then n diagonal entries of the factor
    double* dg = L + (size_t)(n + 1) * ld;
    const int tid = threadIdx.x;
    __shared__ int s_fail;
    if (tid == 0) s_fail = 0;
    for (int e = tid; e < n * n; e += CHOL_THREADS) {
        const int j = e / n, i = e - j * n;        // column-major source
        if (i >= j) L[(size_t)i * ld + j] = Ag[e];
    }
    for (int j = tid; j < n; j += CHOL_THREADS) L[(size_t)n * ld + j] = b[j];
    __syncthreads();
    for (int k = 0; k < n; ++k) {
        const double* rk = L + (size_t)k * ld;
        // A[k,k] is never overwritten: the factor's diagonal lives in dg[].  Four independent chains per dot product:
        // a single dependent DFMA chain costs ~10 cycles per term on this part.
        double d0 = rk[k], d1 = 0.0, d2 = 0.0, d3 = 0.0;
        {
            int m = 0;
            for (; m + 4 <= k; m += 4) {
                d0 -= rk[m] * rk[m]; d1 -= rk[m + 1] * rk[m + 1]; d2 -= rk[m + 2] * rk[m + 2]; d3 -= rk[m + 3] * rk[m + 3];
            }
            for (; m < k; ++m) d0 -= rk[m] * rk[m];
        }
        const double dkk = (d0 + d1) + (d2 + d3);
        const bool bad = !(dkk > 0.0) || !isfinite(dkk);
        const double ipiv = bad ? 0.0 : rsqrt(dkk);
        if (tid == 0) {
            dg[k] = dkk * ipiv;
            if (bad && s_fail == 0) s_fail = k + 1;
        }
        for (int i = k + 1 + tid; i <= n; i += CHOL_THREADS) {
            double* ri = L + (size_t)i * ld;
            double v0 = ri[k], v1 = 0.0, v2 = 0.0, v3 = 0.0;
            int m = 0;
            for (; m + 4 <= k; m += 4) {
                v0 -= ri[m] * rk[m]; v1 -= ri[m + 1] * rk[m + 1]; v2 -= ri[m + 2] * rk[m + 2]; v3 -= ri[m + 3] * rk[m + 3];
            }
            for (; m < k; ++m) v0 -= ri[m] * rk[m];
            ri[k] = ((v0 + v1) + (v2 + v3)) * ipiv;   // only thread i ever touches L[i,k] during this step
        }
        __syncthreads();
        if (s_fail) break;
    }
    if (s_fail) {
        if (tid == 0) *fail = (double)s_fail;
        return;
    }
    // backward substitution L^T x = y (y = row n); the solution overwrites row n
    double* y = L + (size_t)n * ld;
    for (int k = n - 1; k >= 0; --k) {
        const double* rk = L + (size_t)k * ld;
        if (tid == 0) y[k] = y[k] / dg[k];
        __syncthreads();
        const double xk = y[k];
        for (int i = tid; i < k; i += CHOL_THREADS) y[i] -= rk[i] * xk;
        __syncthreads();
    }
    for (int i = tid; i < n; i += CHOL_THREADS) x[i] = y[i];
    for (int e = tid; e < n * n; e += CHOL_THREADS) {
        const int j = e / n, i = e - j * n;
        Ag[e] = (i > j) ? L[(size_t)i * ld + j] : (i == j ? dg[j] : 0.0);
    }
    if (tid == 0) *fail = 0.0;
}

constexpr int CHOL_RL_MAX_N = 100;     // two (n+1) x (n|1) buffers must fit in shared memory

// Small systems (n <= 100): right-looking with ONE barrier per column and no serial section.
// W holds the running Schur complement (never scaled), Lf receives the factor.  In step k every thread reads the
// pivot W[k,k] itself, forms 1/pivot, and applies W[i,j] -= W[i,k] W[j,k] / W[k,k] to its share of the trailing
// triangle (16 x 16 thread tile, rows strided by 16 over ty, columns over tx); column k of W is only read in step k,
// so nothing it needs is overwritten.  The threads with tx == 0 also emit L[i,k] = W[i,k] / sqrt(W[k,k]).
// The right-hand side is row n of W, so row n of Lf ends up as y = L^-1 rhs.
constexpr int CHOL_SMALL_THREADS = 1024;     // 32 x 32 thread tile: 8 warps per scheduler hide the FP64 / LDS latencies

template <int NT>     // NT x NT register tile per thread: rows/columns k+1+t+32a, a < NT  (32*NT >= n+1)
__global__ void __launch_bounds__(CHOL_SMALL_THREADS)
k_cholesky_solve_small(double* Ag, double* b, double* x, int n, double* fail)
{
    extern __shared__ double sh[];
    const int ld = n | 1;
    double* W = sh;                                  // running Schur complement, (n+1) x ld, row n = rhs
    double* Lf = sh + (n + 1) * ld;                  // factor, same shape; row n = y = L^-1 rhs
    __shared__ double s_ip[CHOL_RL_MAX_N + 1];       // 1 / L[k,k]
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    for (int j = ty; j < n; j += 32)
        for (int i = j + tx; i < n; i += 32) W[i * ld + j] = Ag[(size_t)j * n + i];     // column-major source
    for (int j = tid; j < n; j += CHOL_SMALL_THREADS) W[n * ld + j] = b[j];
    __syncthreads();
    int failed = 0;
    for (int k = 0; k < n; ++k) {
        const double dkk = W[k * ld + k];
        if (!(dkk > 0.0) || !isfinite(dkk)) { failed = k + 1; break; }      // uniform: every thread reads the same value
        const double inv = fast_rcp(dkk);
        // column k of the running complement for this thread's rows and columns (read-only during this step)
        double ci[NT], cj[NT];
#pragma unroll
        for (int a = 0; a < NT; ++a) {
            const int i = k + 1 + ty + 32 * a, j = k + 1 + tx + 32 * a;
            ci[a] = i <= n ? W[i * ld + k] * inv : 0.0;
            cj[a] = j < n ? W[j * ld + k] : 0.0;
        }
        // factor column k (rows k..n), one element per thread of the first warps, from the same unmodified column
        if (k + tid <= n) {
            const double ip = fast_rsqrt(dkk);
            Lf[(k + tid) * ld + k] = W[(k + tid) * ld + k] * ip;
            if (tid == 0) s_ip[k] = ip;
        }
        // all loads of the tile, then all FMAs, then all stores: shared-memory stores would otherwise serialise
        // the loop (the compiler must assume they alias the next loads)
        double w[NT][NT];
#pragma unroll
        for (int a = 0; a < NT; ++a)
#pragma unroll
            for (int c = 0; c < NT; ++c) {
                const int i = k + 1 + ty + 32 * a, j = k + 1 + tx + 32 * c;
                w[a][c] = (i <= n && j <= i && j < n) ? W[i * ld + j] : 0.0;
            }
#pragma unroll
        for (int a = 0; a < NT; ++a)
#pragma unroll
            for (int c = 0; c < NT; ++c) w[a][c] -= ci[a] * cj[c];
#pragma unroll
        for (int a = 0; a < NT; ++a)
#pragma unroll
            for (int c = 0; c < NT; ++c) {
                const int i = k + 1 + ty + 32 * a, j = k + 1 + tx + 32 * c;
                if (i <= n && j <= i && j < n) W[i * ld + j] = w[a][c];
            }
        __syncthreads();
    }
    if (failed) {
        if (tid == 0) *fail = (double)failed;
        return;
    }
    // backward substitution L^T x = y by ONE warp, y in registers (lane l holds y[l + 32 m]), no block barriers
    if (ty == 0) {
        double y0 = tx < n ? Lf[n * ld + tx] : 0.0;
        double y1 = tx + 32 < n ? Lf[n * ld + tx + 32] : 0.0;
        double y2 = tx + 64 < n ? Lf[n * ld + tx + 64] : 0.0;
        double y3 = tx + 96 < n ? Lf[n * ld + tx + 96] : 0.0;
        for (int k = n - 1; k >= 0; --k) {
            const int m = k >> 5;
            const double yk = m == 0 ? y0 : (m == 1 ? y1 : (m == 2 ? y2 : y3));
            const double xk = __shfl_sync(0xffffffffu, yk, k & 31) * s_ip[k];
            const double* rk = Lf + k * ld;
            if (tx < k) y0 -= rk[tx] * xk; else if (tx == k) y0 = xk;
            if (k >= 32) { if (tx + 32 < k) y1 -= rk[tx + 32] * xk; else if (tx + 32 == k) y1 = xk; }
            if (k >= 64) { if (tx + 64 < k) y2 -= rk[tx + 64] * xk; else if (tx + 64 == k) y2 = xk; }
            if (k >= 96) { if (tx + 96 < k) y3 -= rk[tx + 96] * xk; else if (tx + 96 == k) y3 = xk; }
        }
        if (tx < n) x[tx] = y0;
        if (tx + 32 < n) x[tx + 32] = y1;
        if (tx + 64 < n) x[tx + 64] = y2;
        if (tx + 96 < n) x[tx + 96] = y3;
    }
    for (int j = ty; j < n; j += 32)
        for (int i = tx; i < n; i += 32) Ag[(size_t)j * n + i] = (i >= j) ? Lf[i * ld + j] : 0.0;
    if (tid == 0) *fail = 0.0;
}

// `work` must hold (n+1)*(n|1)+n doubles when n > CHOL_SMEM_MAX_N (ignored otherwise)
int launch_cholesky_solve(double* A_dev, double* b_dev, double* x_dev, int n, double* fail_dev, double* work_dev,
                          cudaStream_t stream)
{
    if (n <= CHOL_RL_MAX_N) {
        const size_t bytes = 2 * (size_t)(n + 1) * (n | 1) * sizeof(double);
        static bool attr_small = false;
        const int max_bytes = 2 * (CHOL_RL_MAX_N + 1) * (CHOL_RL_MAX_N | 1) * (int)sizeof(double);
        if (!attr_small) {
            SBA_CUDA(cudaFuncSetAttribute(k_cholesky_solve_small<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_bytes));
            SBA_CUDA(cudaFuncSetAttribute(k_cholesky_solve_small<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_bytes));
            SBA_CUDA(cudaFuncSetAttribute(k_cholesky_solve_small<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_bytes));
            attr_small = true;
        }
        if (n + 1 <= 32) k_cholesky_solve_small<1><<<1, CHOL_SMALL_THREADS, bytes, stream>>>(A_dev, b_dev, x_dev, n, fail_dev);
        else if (n + 1 <= 64) k_cholesky_solve_small<2><<<1, CHOL_SMALL_THREADS, bytes, stream>>>(A_dev, b_dev, x_dev, n, fail_dev);
        else k_cholesky_solve_small<4><<<1, CHOL_SMALL_THREADS, bytes, stream>>>(A_dev, b_dev, x_dev, n, fail_dev);
    } else if (n <= CHOL_SMEM_MAX_N) {
        const size_t bytes = ((size_t)(n + 1) * (n | 1) + n) * sizeof(double);
        static bool attr_set = false;
        if (!attr_set) {
            SBA_CUDA(cudaFuncSetAttribute(k_cholesky_solve<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          ((CHOL_SMEM_MAX_N + 1) * (CHOL_SMEM_MAX_N | 1) + CHOL_SMEM_MAX_N) * (int)sizeof(double)));
            attr_set = true;
        }
        k_cholesky_solve<true><<<1, CHOL_THREADS, bytes, stream>>>(A_dev, b_dev, x_dev, n, fail_dev, nullptr);
    } else {
        if (!work_dev) { set_error("cholesky: workspace required for n > 160"); return SBA_E_INVALID; }
        k_cholesky_solve<false><<<1, CHOL_THREADS, 0, stream>>>(A_dev, b_dev, x_dev, n, fail_dev, work_dev);
    }
    SBA_CUDA(cudaGetLastError());
    return SBA_OK;
}

}  // namespace sba
