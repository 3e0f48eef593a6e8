// FP64 dense Cholesky factorisation + triangular solves of the reduced camera system S dc = rhs
// (n = n_cam * n_params <= ~1800).  There is no reference counterpart: the reference never forms S,
// it hands the full sparse Jacobian to LSMR (scipy/optimize/_lsq/trf.py:485-495).
//
// One CTA, right-looking.  The right-hand side rides along as an extra row of the matrix, so the forward
// substitution L y = rhs falls out of the factorisation itself (row n of L is y^T) and only the backward
// substitution remains.  Columns of the trailing update are dealt to warps, rows to lanes (coalesced /
// conflict-free in the column-major layout).  S is staged in shared memory when it fits
// (n <= 160: 161 x 160 doubles = 206 KB of the 227 KB a CTA may use), otherwise worked on in place in
// global memory, where the whole matrix is L2-resident.
#include "sba_internal.cuh"

namespace sba {

constexpr int CHOL_THREADS = 256;
constexpr int CHOL_SMEM_MAX_N = 160;

// A: n x n column-major, lower triangle used, overwritten by L (full n x n written back).
// b: rhs (n).  x: solution (n).  fail: (k+1) when pivot k is not positive / finite, else 0.
template <bool SMEM>
__global__ void __launch_bounds__(CHOL_THREADS)
k_cholesky_solve(double* Ag, double* b, double* x, int n, double* fail, double* work)
{
    extern __shared__ double sh[];
    // augmented storage: (n+1) rows x n columns, leading dimension ld = n+1, row n = rhs^T
    const int ld = n + 1;
    double* A = SMEM ? sh : work;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = CHOL_THREADS / 32;
    __shared__ double s_ipiv;
    __shared__ int s_fail;
    if (tid == 0) s_fail = 0;
    for (int e = tid; e < n * n; e += CHOL_THREADS) {
        const int j = e / n, i = e - j * n;
        A[i + (size_t)j * ld] = Ag[e];
    }
    for (int j = tid; j < n; j += CHOL_THREADS) A[n + (size_t)j * ld] = b[j];
    __syncthreads();
    for (int k = 0; k < n; ++k) {
        if (tid == 0) {
            const double d = A[k + (size_t)k * ld];
            if (!(d > 0.0) || !isfinite(d)) { s_fail = k + 1; s_ipiv = 0.0; }
            else { const double p = sqrt(d); A[k + (size_t)k * ld] = p; s_ipiv = 1.0 / p; }
        }
        __syncthreads();
        if (s_fail) break;
        const double ipiv = s_ipiv;
        double* colk = A + (size_t)k * ld;
        for (int i = k + 1 + tid; i <= n; i += CHOL_THREADS) colk[i] *= ipiv;
        __syncthreads();
        // trailing update: A[i,j] -= L[i,k] L[j,k] for k < j < n, j <= i <= n
        for (int j = k + 1 + warp; j < n; j += NW) {
            const double ljk = colk[j];
            double* colj = A + (size_t)j * ld;
            for (int i = j + lane; i <= n; i += 32) colj[i] -= colk[i] * ljk;
        }
        __syncthreads();
    }
    if (s_fail) {
        if (tid == 0) *fail = (double)s_fail;
        return;
    }
    // backward substitution L^T x = y, y = row n of the factor; solution accumulates in row n
    for (int k = n - 1; k >= 0; --k) {
        if (tid == 0) A[n + (size_t)k * ld] /= A[k + (size_t)k * ld];
        __syncthreads();
        const double xk = A[n + (size_t)k * ld];
        for (int i = tid; i < k; i += CHOL_THREADS) A[n + (size_t)i * ld] -= A[k + (size_t)i * ld] * xk;
        __syncthreads();
    }
    for (int i = tid; i < n; i += CHOL_THREADS) x[i] = A[n + (size_t)i * ld];
    for (int e = tid; e < n * n; e += CHOL_THREADS) {
        const int j = e / n, i = e - j * n;
        Ag[e] = A[i + (size_t)j * ld];
    }
    if (tid == 0) *fail = 0.0;
}

// `work` must hold (n+1)*n doubles when n > CHOL_SMEM_MAX_N (ignored otherwise)
int launch_cholesky_solve(double* A_dev, double* b_dev, double* x_dev, int n, double* fail_dev, double* work_dev,
                          cudaStream_t stream)
{
    if (n <= CHOL_SMEM_MAX_N) {
        const size_t bytes = (size_t)(n + 1) * n * sizeof(double);
        static bool attr_set = false;
        if (!attr_set) {
            SBA_CUDA(cudaFuncSetAttribute(k_cholesky_solve<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (CHOL_SMEM_MAX_N + 1) * CHOL_SMEM_MAX_N * (int)sizeof(double)));
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
