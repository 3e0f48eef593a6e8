// FP64 dense Cholesky factorisation + triangular solves of the reduced camera system S dc = rhs
// (n = n_cam * n_params <= ~1800).  There is no reference counterpart: the reference never forms S,
// it hands the full sparse Jacobian to LSMR (scipy/optimize/_lsq/trf.py:485-495).
//
// Round-1 implementation: one CTA, right-looking, column by column, S staged in shared memory when it
// fits (n <= 160: 200 KB of the 227 KB a CTA may use), otherwise worked on in place (L2-resident).
#include "sba_internal.cuh"

namespace sba {

constexpr int CHOL_THREADS = 512;
constexpr int CHOL_SMEM_MAX_N = 160;

// A: n x n column-major (lower triangle used, overwritten by L); b: rhs (overwritten by y); x: solution.
// fail: set to (k+1) when pivot k is not positive / finite.
template <bool SMEM>
__global__ void __launch_bounds__(CHOL_THREADS)
k_cholesky_solve(double* Ag, double* b, double* x, int n, double* fail)
{
    extern __shared__ double sh[];
    double* A = SMEM ? sh : Ag;
    const int tid = threadIdx.x;
    __shared__ double s_piv;
    __shared__ int s_fail;
    if (tid == 0) s_fail = 0;
    if (SMEM) {
        for (int e = tid; e < n * n; e += CHOL_THREADS) A[e] = Ag[e];
    }
    __syncthreads();
    for (int k = 0; k < n; ++k) {
        if (tid == 0) {
            const double d = A[k + (size_t)k * n];
            if (!(d > 0.0) || !isfinite(d)) { s_fail = k + 1; s_piv = 1.0; }
            else s_piv = sqrt(d);
        }
        __syncthreads();
        if (s_fail) break;
        const double piv = s_piv, ipiv = 1.0 / piv;
        if (tid == 0) A[k + (size_t)k * n] = piv;
        for (int i = k + 1 + tid; i < n; i += CHOL_THREADS) A[i + (size_t)k * n] *= ipiv;
        __syncthreads();
        // trailing update of the lower triangle: A[i,j] -= L[i,k] L[j,k], k < j <= i < n
        const int m = n - k - 1;
        for (int e = tid; e < m * m; e += CHOL_THREADS) {
            const int jj = e / m, ii = e - jj * m;
            if (ii >= jj) {
                const int i = k + 1 + ii, j = k + 1 + jj;
                A[i + (size_t)j * n] -= A[i + (size_t)k * n] * A[j + (size_t)k * n];
            }
        }
        __syncthreads();
    }
    if (s_fail) {
        if (tid == 0) *fail = (double)s_fail;
        return;
    }
    if (tid == 0) *fail = 0.0;
    // forward substitution L y = b (b overwritten)
    for (int k = 0; k < n; ++k) {
        if (tid == 0) b[k] = b[k] / A[k + (size_t)k * n];
        __syncthreads();
        const double yk = b[k];
        for (int i = k + 1 + tid; i < n; i += CHOL_THREADS) b[i] -= A[i + (size_t)k * n] * yk;
        __syncthreads();
    }
    // backward substitution L^T x = y
    for (int k = n - 1; k >= 0; --k) {
        if (tid == 0) b[k] = b[k] / A[k + (size_t)k * n];
        __syncthreads();
        const double xk = b[k];
        for (int i = tid; i < k; i += CHOL_THREADS) b[i] -= A[k + (size_t)i * n] * xk;
        __syncthreads();
    }
    for (int i = tid; i < n; i += CHOL_THREADS) x[i] = b[i];
    if (SMEM) {
        for (int e = tid; e < n * n; e += CHOL_THREADS) Ag[e] = A[e];
    }
}

int launch_cholesky_solve(double* A_dev, double* b_dev, double* x_dev, int n, double* fail_dev, cudaStream_t stream)
{
    if (n <= CHOL_SMEM_MAX_N) {
        const size_t bytes = (size_t)n * n * sizeof(double);
        static bool attr_set = false;
        if (!attr_set) {
            SBA_CUDA(cudaFuncSetAttribute(k_cholesky_solve<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          CHOL_SMEM_MAX_N * CHOL_SMEM_MAX_N * (int)sizeof(double)));
            attr_set = true;
        }
        k_cholesky_solve<true><<<1, CHOL_THREADS, bytes, stream>>>(A_dev, b_dev, x_dev, n, fail_dev);
    } else {
        k_cholesky_solve<false><<<1, CHOL_THREADS, 0, stream>>>(A_dev, b_dev, x_dev, n, fail_dev);
    }
    SBA_CUDA(cudaGetLastError());
    return SBA_OK;
}

}  // namespace sba
