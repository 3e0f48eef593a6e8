// FP64 dense Cholesky factorisation + triangular solves of the reduced camera system S dc = rhs
// (n = n_cam * n_params <= ~1800).  There is no reference counterpart: the reference never forms S,
// it hands the full sparse Jacobian to LSMR (scipy/optimize/_lsq/trf.py:485-495).
//
// Two size classes, both blocked by 32 columns with the diagonal block factored by one warp in registers:
// n <= 127 one CTA with the matrix in shared memory (k_chol_fused); larger n a multi-CTA kernel sequence working in
// place in the L2-resident matrix (k_chol_panel / k_chol_update / k_chol_backsolve).  In both the right-hand side
// rides along as an extra row, so the forward substitution L y = rhs falls out of the factorisation and only the
// backward substitution remains.
#include "sba_internal.cuh"
#include <math_constants.h>

namespace sba {

constexpr int CB = 32, CB_LD = CB + 1, CU_TILE = 64;
constexpr int CHOL_FUSED_MAX_N = 127;        // one CTA, matrix in shared memory, 4 x 4 register tiles of the trailing update
__device__ long long g_chol_clk[16];         // stage clocks of the last k_chol_fused launch (diagnostics, tools/chol_time.py)
#define CHOL_CLK(i) do { if (threadIdx.x == 0) g_chol_clk[i] = clock64(); } while (0)
constexpr int CHOL_FUSED_WARPS = 16, CHOL_FUSED_THREADS = 32 * CHOL_FUSED_WARPS;    // 128 registers per thread: warp_chol32 needs ~110

// Cholesky factorisation of a 32 x 32 block held by ONE warp in registers: lane r owns row r (a[c], c <= r; identity
// padding for unused rows), right-looking, the pivot column travels by shuffles; entries above the diagonal are scratch
// and never leave their lane.  ~150 dependent cycles per column instead of the ~1500 of a shared-memory/barrier version.
// Returns 0 or (c+1) for the first non-positive pivot (uniform across the warp); inv receives 1 / L[lane, lane].
// 1 / sqrt(x) for the pivots: hardware seed (rsqrt.approx.ftz.f64, ~2^-22) and ONE cubically convergent step, r (1 + e/2 + 3 e^2 / 8) with e = 1 - x r^2 --
// four dependent FP64 operations instead of the six of two Newton steps; the remainder 5 e^3 / 16 is below 2^-64.
__device__ __forceinline__ double pivot_rsqrt(double x)
{
    double r;                                  // MUFU.RSQ64H on the high word: no conversions to and from FP32 on the pivot chain
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-(x * r), r, 1.0);
    return fma(r * e, fma(0.375, e, 0.5), r);
}

__device__ __forceinline__ int warp_chol32(double (&a)[CB], int lane, double& inv)
{
    inv = 1.0;
    int bad = 0;
    // No branch inside the column loop: a failed pivot is recorded and replaced by 1, so the 32 columns are ONE basic block and
    // the pending updates of column c overlap the dependent chain (shuffle, rsqrt, scale) of column c + 1.
#pragma unroll
    for (int c = 0; c < CB; ++c) {
        const double d = __shfl_sync(0xffffffffu, a[c], c);
        const bool ok = d > 0.0 && d < CUDART_INF;
        bad = (bad == 0 && !ok) ? c + 1 : bad;
        const double ip = pivot_rsqrt(ok ? d : 1.0);
        const double l = a[c] * ip;
        a[c] = l;
        if (lane == c) inv = ip;
#pragma unroll
        for (int k = c + 1; k < CB; ++k) a[k] = fma(-l, __shfl_sync(0xffffffffu, l, k), a[k]);
    }
    return bad;
}

// n <= 127: the whole solve in one CTA.  W = lower triangle of A plus the right-hand side as row n, in shared memory
// with an odd row stride.  Per panel of 32 columns: warp 0 factors the diagonal block in registers (warp_chol32); every
// warp solves its rows of the panel against it (2 NT independent rows per warp, interleaved); all threads apply the
// panel to the trailing triangle in register tiles (row warp + 16 a, column lane + 32 q per thread).  Row n ends up as y = L^-1 rhs, and warp 0 finishes with
// the backward substitution, y distributed over its lanes.
// A: n x n column-major, lower triangle used, overwritten by L when write_factor is set (the solver only needs x).
// fail: (k+1) when pivot k is not positive, else 0.
template <int NT>     // 32 * NT >= n + 1
__global__ void __launch_bounds__(CHOL_FUSED_THREADS)
k_chol_fused(double* Ag, const double* b, double* x, int n, double* fail, bool write_factor)
{
    extern __shared__ double W[];                    // (n+1) x ld
    CHOL_CLK(0);
    __shared__ double s_inv[CB * NT];
    __shared__ int s_bad;
    const int ld = n | 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int FW = CHOL_FUSED_WARPS, RPW = 2 * NT;      // rows per warp
    for (int e0 = 0; e0 < n * n; e0 += 8 * CHOL_FUSED_THREADS) {        // 8 loads in flight per thread, then the stores
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = e0 + tid + CHOL_FUSED_THREADS * u;
            v[u] = e < n * n ? Ag[e] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = e0 + tid + CHOL_FUSED_THREADS * u, j = e / n, i = e - j * n;     // column-major source
            if (e < n * n && i >= j) W[i * ld + j] = v[u];
        }
    }
    for (int j = tid; j < n; j += CHOL_FUSED_THREADS) W[n * ld + j] = b[j];
    if (tid == 0) s_bad = 0;
    __syncthreads();
    CHOL_CLK(1);
    for (int j0 = 0; j0 < n; j0 += CB) {
        const int nb = min(CB, n - j0), m0 = j0 + nb;
        if (warp == 0) {
            double a[CB], inv;
#pragma unroll
            for (int c = 0; c < CB; ++c)
                a[c] = (lane < nb && c <= lane) ? W[(j0 + lane) * ld + j0 + c] : (c == lane ? 1.0 : 0.0);
            const int bad = warp_chol32(a, lane, inv);
            if (bad) { if (lane == 0) s_bad = j0 + bad; }
            else if (lane < nb) {
#pragma unroll
                for (int c = 0; c < CB; ++c) if (c <= lane) W[(j0 + lane) * ld + j0 + c] = a[c];
                s_inv[j0 + lane] = inv;
            }
        }
        __syncthreads();
        if (j0 < 2 * CB) CHOL_CLK(2 + 3 * (j0 / CB));
        if (s_bad) break;
        // rows m0 + warp + 16 k of the panel (row n = right-hand side): X L_jj^T = rows
        {
            double xv[RPW];
#pragma unroll
            for (int k = 0; k < RPW; ++k) {
                const int r = m0 + warp + FW * k;
                xv[k] = (r <= n && lane < nb) ? W[r * ld + j0 + lane] : 0.0;
            }
            if (m0 + warp <= n) {
                // row `lane` of the factored block and the inverse pivots in registers: the substitution below is then a pure
                // shuffle -> multiply -> fma chain (no shared-memory latency inside it)
                double lrow[CB];
#pragma unroll
                for (int c = 0; c < CB; ++c) lrow[c] = (lane < nb && c < nb && c <= lane) ? W[(j0 + lane) * ld + j0 + c] : 0.0;
                const double invl = lane < nb ? s_inv[j0 + lane] : 1.0;
#pragma unroll
                for (int c = 0; c < CB; ++c) {
                    if (c < nb) {                                   // uniform
                        const double ic = __shfl_sync(0xffffffffu, invl, c);
#pragma unroll
                        for (int k = 0; k < RPW; ++k) {
                            const double xc = __shfl_sync(0xffffffffu, xv[k], c) * ic;
                            xv[k] = lane == c ? xc : fma(-xc, lrow[c], xv[k]);        // lrow[c] = 0 for lanes < c: they are done
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < RPW; ++k) {
                    const int r = m0 + warp + FW * k;
                    if (r <= n && lane < nb) W[r * ld + j0 + lane] = xv[k];
                }
            }
        }
        __syncthreads();
        if (j0 < 2 * CB) CHOL_CLK(3 + 3 * (j0 / CB));
        // trailing triangle (rows <= n, columns < n): W[i,j] -= sum_c W[i,j0+c] W[j,j0+c]; i = m0 + warp + 16 a, j = m0 + lane + 32 q
        if (m0 < n) {
            double acc[RPW][NT];
#pragma unroll
            for (int a = 0; a < RPW; ++a)
#pragma unroll
                for (int q = 0; q < NT; ++q) acc[a][q] = 0.0;
            for (int c = 0; c < nb; ++c) {
                double ri[RPW], rj[NT];
#pragma unroll
                for (int a = 0; a < RPW; ++a) {
                    const int i = m0 + warp + FW * a;
                    ri[a] = i <= n ? W[i * ld + j0 + c] : 0.0;
                }
#pragma unroll
                for (int q = 0; q < NT; ++q) {
                    const int j = m0 + lane + 32 * q;
                    rj[q] = j < n ? W[j * ld + j0 + c] : 0.0;
                }
#pragma unroll
                for (int a = 0; a < RPW; ++a)
#pragma unroll
                    for (int q = 0; q <= a / 2; ++q) acc[a][q] += ri[a] * rj[q];      // tiles with q > a/2 lie above the diagonal
            }
#pragma unroll
            for (int a = 0; a < RPW; ++a)
#pragma unroll
                for (int q = 0; q <= a / 2; ++q) {
                    const int i = m0 + warp + FW * a, j = m0 + lane + 32 * q;
                    if (i <= n && j < n && j <= i) W[i * ld + j] -= acc[a][q];
                }
        }
        __syncthreads();
        if (j0 < 2 * CB) CHOL_CLK(4 + 3 * (j0 / CB));
    }
    if (s_bad) {
        if (tid == 0) *fail = (double)s_bad;
        return;
    }
    // backward substitution L^T x = y by ONE warp, y in registers (lane l holds y[l + 32 m]), no block barriers
    if (warp == 0) {
        double yv[NT];
#pragma unroll
        for (int m = 0; m < NT; ++m) yv[m] = lane + 32 * m < n ? W[n * ld + lane + 32 * m] : 0.0;
        double invv[NT];
#pragma unroll
        for (int m = 0; m < NT; ++m) invv[m] = lane + 32 * m < n ? s_inv[lane + 32 * m] : 1.0;
#pragma unroll
        for (int m = NT - 1; m >= 0; --m) {
            if (32 * m >= n) continue;                                       // uniform
#pragma unroll 8
            for (int kk = 31; kk >= 0; --kk) {
                const int k = 32 * m + kk;
                if (k < n) {                                                 // uniform
                    // the row of L is fetched before the value it multiplies is known: only shuffle -> fma stays on the chain
                    double lk[NT];
#pragma unroll
                    for (int m2 = 0; m2 <= m; ++m2) lk[m2] = lane + 32 * m2 < k ? W[k * ld + lane + 32 * m2] : 0.0;
                    const double xk = __shfl_sync(0xffffffffu, yv[m] * invv[m], kk);
#pragma unroll
                    for (int m2 = 0; m2 <= m; ++m2) yv[m2] = (lane + 32 * m2 == k) ? xk : fma(-lk[m2], xk, yv[m2]);
                }
            }
        }
#pragma unroll
        for (int m = 0; m < NT; ++m) if (lane + 32 * m < n) x[lane + 32 * m] = yv[m];
        CHOL_CLK(8);
    }
    if (write_factor)
        for (int j = warp; j < n; j += FW)
            for (int i = lane; i < n; i += 32) Ag[(size_t)j * n + i] = (i >= j) ? W[i * ld + j] : 0.0;
    if (tid == 0) *fail = 0.0;
}

// ---------------------------------------------------------------------------------------------------------------
// Blocked right-looking factorisation for n > CHOL_FUSED_MAX_N (cfg3: n = 300, cfg4: n = 1800), many CTAs, in place
// in the column-major matrix (which stays L2-resident: 26 MB at n = 1800).  Per panel of CB = 32 columns:
//   k_chol_panel   every CTA factors the 32 x 32 diagonal block redundantly (same arithmetic, same bits) in shared
//                  memory with one warp, then each warp solves one row of the panel below it against that block;
//                  CTA 0 also carries the right-hand side along as one more row (forward substitution for free);
//   k_chol_update  trailing lower triangle -= panel panel^T in 64 x 64 tiles, 4 x 4 outputs per thread, and the
//                  right-hand side below the panel -= panel y_panel.
// then k_chol_backsolve (one CTA) solves L^T x = y block by block.  A non-positive pivot writes (k+1) to *fail and
// every later kernel of the sequence returns at once; there is no host round trip inside the sequence.

__global__ void __launch_bounds__(256)
k_chol_panel(double* A, double* y, double* D, double* Dinv, int n, int j0, double* fail)
{
    __shared__ double Ls[CB * CB_LD];
    __shared__ double s_inv[CB];
    __shared__ int s_bad;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nb = min(CB, n - j0), m0 = j0 + nb;
    const double failed = *fail;
    // this warp's row of the panel (or the right-hand side): requested before the factorisation, used after it
    const int row = m0 + blockIdx.x * 8 + warp;
    const bool is_rhs = (row == n);
    double xv = 0.0;
    if (row <= n && lane < nb) xv = is_rhs ? y[j0 + lane] : A[(size_t)(j0 + lane) * n + row];
    // 32 x 32 diagonal block, right-looking in registers of warp 0: lane r holds row r (identity padding beyond nb), the
    // pivot column travels by shuffles; entries above the diagonal are scratch and never leave their lane
    double a[CB];
    if (warp == 0) {
#pragma unroll
        for (int c = 0; c < CB; ++c)
            a[c] = (lane < nb && c <= lane) ? A[(size_t)(j0 + c) * n + j0 + lane] : (c == lane ? 1.0 : 0.0);
    }
    if (failed != 0.0) return;
    if (tid == 0) s_bad = 0;
    __syncthreads();
    if (warp == 0) {
        double my_inv;
        const int bad = warp_chol32(a, lane, my_inv);
        if (bad) { if (lane == 0) s_bad = bad; }
        else {
#pragma unroll
            for (int c = 0; c < CB; ++c) Ls[lane * CB_LD + c] = c <= lane ? a[c] : 0.0;
            s_inv[lane] = my_inv;
            // the factored block goes to the scratch D ([c * 32 + r]), NOT into A: other CTAs of this launch may still
            // be reading the unfactored block from A
            if (blockIdx.x == 0) {
#pragma unroll
                for (int c = 0; c < CB; ++c) D[(size_t)(j0 / CB) * CB * CB + c * CB + lane] = (c <= lane && lane < nb) ? a[c] : 0.0;
                if (lane < nb) Dinv[j0 + lane] = my_inv;
            }
        }
    }
    __syncthreads();
    if (s_bad) {
        if (blockIdx.x == 0 && tid == 0) *fail = (double)(j0 + s_bad);
        return;
    }
    // one row of the panel per warp: X L_jj^T = A_row
    if (row > n) return;
    for (int c = 0; c < nb; ++c) {
        const double xc = __shfl_sync(0xffffffffu, xv, c) * s_inv[c];
        if (lane == c) xv = xc;
        else if (lane > c) xv -= xc * Ls[lane * CB_LD + c];
    }
    if (lane < nb) {
        if (is_rhs) y[j0 + lane] = xv; else A[(size_t)(j0 + lane) * n + row] = xv;
    }
}

__global__ void __launch_bounds__(256)
k_chol_update(double* A, double* y, const double* D, int n, int j0, const double* fail, bool export_diag)
{
    const int ti = blockIdx.y, tj = blockIdx.x;
    if (tj > ti) return;
    __shared__ double Pi[CB][CU_TILE], Pj[CB][CU_TILE];
    __shared__ double ys[CB];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int nb = min(CB, n - j0), m0 = j0 + nb;
    const int i0 = m0 + ti * CU_TILE, jj0 = m0 + tj * CU_TILE;
    // every global read of this CTA is requested before the first one is consumed (one memory round trip, not 25)
    const double failed = *fail;
    constexpr int NLD = CB * CU_TILE / 256;
    double pi[NLD], pj[NLD], old[4][4];
#pragma unroll
    for (int u = 0; u < NLD; ++u) {
        const int e = tid + 256 * u, c = e / CU_TILE, r = e - c * CU_TILE;
        pi[u] = (c < nb && i0 + r < n) ? A[(size_t)(j0 + c) * n + i0 + r] : 0.0;
        pj[u] = (c < nb && jj0 + r < n) ? A[(size_t)(j0 + c) * n + jj0 + r] : 0.0;
    }
#pragma unroll
    for (int l = 0; l < 4; ++l)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = i0 + tx + 16 * k, j = jj0 + ty + 16 * l;
            old[k][l] = (i < n && j <= i) ? A[(size_t)j * n + i] : 0.0;
        }
    const double yv = (tid < nb) ? y[j0 + tid] : 0.0;
    double yold = 0.0, dblk[CB * CB / 256];
    if (tj == 0 && tid < CU_TILE && i0 + tid < n) yold = y[i0 + tid];
    if (ti == 0 && export_diag)                        // tile (0,0) also moves the factored diagonal block into A
#pragma unroll
        for (int u = 0; u < CB * CB / 256; ++u) dblk[u] = D[(size_t)(j0 / CB) * CB * CB + tid + 256 * u];
    if (failed != 0.0) return;
#pragma unroll
    for (int u = 0; u < NLD; ++u) {
        const int e = tid + 256 * u, c = e / CU_TILE, r = e - c * CU_TILE;
        Pi[c][r] = pi[u]; Pj[c][r] = pj[u];
    }
    if (tid < CB) ys[tid] = yv;
    if (ti == 0 && export_diag)
#pragma unroll
        for (int u = 0; u < CB * CB / 256; ++u) {
            const int e = tid + 256 * u, c = e / CB, r = e - c * CB;
            if (r < nb && c < nb) A[(size_t)(j0 + c) * n + j0 + r] = dblk[u];
        }
    __syncthreads();
    double acc[4][4] = {};
#pragma unroll 4
    for (int c = 0; c < CB; ++c) {
        double a[4], b[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { a[k] = Pi[c][tx + 16 * k]; b[k] = Pj[c][ty + 16 * k]; }
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int l = 0; l < 4; ++l) acc[k][l] += a[k] * b[l];
    }
#pragma unroll
    for (int l = 0; l < 4; ++l)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = i0 + tx + 16 * k, j = jj0 + ty + 16 * l;
            if (i < n && j <= i) A[(size_t)j * n + i] = old[k][l] - acc[k][l];
        }
    // right-hand side below the panel (the column-0 tiles cover every row once)
    if (tj == 0 && tid < CU_TILE && i0 + tid < n) {
        double s0 = 0.0;
#pragma unroll 8
        for (int c = 0; c < CB; ++c) s0 += Pi[c][tid] * ys[c];
        y[i0 + tid] = yold - s0;
    }
}

// L^T x = y, one CTA, 32-column blocks from the bottom: warp 0 solves the diagonal block (staged in shared memory, the
// next one already in flight), then groups of 4 threads remove that block's contribution from one earlier unknown,
// two unknowns per thread in flight.
__global__ void __launch_bounds__(1024)
k_chol_backsolve(const double* __restrict__ A, const double* __restrict__ y, const double* __restrict__ D,
                 const double* __restrict__ Dinv, double* __restrict__ x, int n, const double* __restrict__ fail)
{
    if (*fail != 0.0) return;
    extern __shared__ double xs[];                 // x (n), then 1 / L[k,k] (n)
    double* xinv = xs + n;
    __shared__ double Ls[CB * CB_LD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nblk = (n + CB - 1) / CB;
    double pre = D[(size_t)(nblk - 1) * CB * CB + tid];
    for (int i = tid; i < n; i += 1024) { xs[i] = y[i]; xinv[i] = Dinv[i]; }
    __syncthreads();
    for (int J = nblk - 1; J >= 0; --J) {
        const int j0 = J * CB, nb = min(CB, n - j0);
        {
            const int c = tid / CB, r = tid - c * CB;
            Ls[r * CB_LD + c] = pre;
            if (J > 0) pre = D[(size_t)(J - 1) * CB * CB + tid];
        }
        __syncthreads();
        if (warp == 0) {
            double v = lane < nb ? xs[j0 + lane] : 0.0;
            for (int c = nb - 1; c >= 0; --c) {
                const double xc = __shfl_sync(0xffffffffu, v, c) * xinv[j0 + c];
                if (lane == c) v = xc;
                else if (lane < c) v -= xc * Ls[c * CB_LD + lane];
            }
            if (lane < nb) xs[j0 + lane] = v;
        }
        __syncthreads();
        const int q = tid & 3, nw = j0 * 4;             // nw is a multiple of 128: whole warps enter the loop body
        double xr[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) xr[c] = (q * 8 + c < nb) ? xs[j0 + q * 8 + c] : 0.0;
        for (int w0 = tid; w0 < nw; w0 += 2048) {
            double cv[2][8];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int w = w0 + 1024 * u;
                const double* col = A + (size_t)(w >> 2) * n + j0 + q * 8;      // L[j0 + c, i], contiguous in c
#pragma unroll
                for (int c = 0; c < 8; ++c) cv[u][c] = (w < nw && q * 8 + c < nb) ? col[c] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int w = w0 + 1024 * u;
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int c = 0; c < 8; c += 2) { s0 += cv[u][c] * xr[c]; s1 += cv[u][c + 1] * xr[c + 1]; }
                double sv = s0 + s1;
                sv += __shfl_xor_sync(0xffffffffu, sv, 1);
                sv += __shfl_xor_sync(0xffffffffu, sv, 2);
                if (q == 0 && w < nw) xs[w >> 2] -= sv;
            }
        }
        __syncthreads();
    }
    for (int i = tid; i < n; i += 1024) x[i] = xs[i];
}

// `work` must hold 34 * (n + 32) doubles when n > CHOL_FUSED_MAX_N (ignored otherwise)
int launch_cholesky_solve(double* A_dev, double* b_dev, double* x_dev, int n, double* fail_dev, double* work_dev,
                          cudaStream_t stream, bool write_factor)
{
    if (n <= CHOL_FUSED_MAX_N) {
        const size_t bytes = (size_t)(n + 1) * (n | 1) * sizeof(double);
        static bool attr_set = false;
        if (!attr_set) {
            const int max_bytes = (CHOL_FUSED_MAX_N + 1) * (CHOL_FUSED_MAX_N | 1) * (int)sizeof(double);
            SBA_CUDA(cudaFuncSetAttribute(k_chol_fused<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_bytes));
            SBA_CUDA(cudaFuncSetAttribute(k_chol_fused<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_bytes));
            SBA_CUDA(cudaFuncSetAttribute(k_chol_fused<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_bytes));
            SBA_CUDA(cudaFuncSetAttribute(k_chol_fused<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_bytes));
            attr_set = true;
        }
        if (n + 1 <= 32) k_chol_fused<1><<<1, CHOL_FUSED_THREADS, bytes, stream>>>(A_dev, b_dev, x_dev, n, fail_dev, write_factor);
        else if (n + 1 <= 64) k_chol_fused<2><<<1, CHOL_FUSED_THREADS, bytes, stream>>>(A_dev, b_dev, x_dev, n, fail_dev, write_factor);
        else if (n + 1 <= 96) k_chol_fused<3><<<1, CHOL_FUSED_THREADS, bytes, stream>>>(A_dev, b_dev, x_dev, n, fail_dev, write_factor);
        else k_chol_fused<4><<<1, CHOL_FUSED_THREADS, bytes, stream>>>(A_dev, b_dev, x_dev, n, fail_dev, write_factor);
    } else {
        // blocked multi-CTA path; the right-hand side is transformed in place in `work_dev` (n doubles)
        if (!work_dev) { set_error("cholesky: workspace required for the blocked factorisation"); return SBA_E_INVALID; }
        if (2 * (size_t)n * sizeof(double) > 200 * 1024) { set_error("cholesky: n too large for the back-substitution kernel"); return SBA_E_INVALID; }
        static bool attr_bs = false;
        if (!attr_bs) {
            SBA_CUDA(cudaFuncSetAttribute(k_chol_backsolve, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_bs = true;
        }
        const size_t n_up = (((size_t)n + CB - 1) / CB) * CB;
        double* Dinv = work_dev + n_up;                                     // 1 / L[k,k]
        double* Dblk = work_dev + 2 * n_up;                                 // factored diagonal blocks, 32 x 32 each
        SBA_CUDA(cudaMemsetAsync(fail_dev, 0, sizeof(double), stream));
        SBA_CUDA(cudaMemcpyAsync(work_dev, b_dev, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, stream));
        for (int j0 = 0; j0 < n; j0 += CB) {
            const int nb = n - j0 < CB ? n - j0 : CB;
            const int rows = n - j0 - nb + 1;                         // panel rows below the block + the right-hand side
            k_chol_panel<<<(rows + 7) / 8, 256, 0, stream>>>(A_dev, work_dev, Dblk, Dinv, n, j0, fail_dev);
            const int m = n - j0 - nb;
            if (m == 0 && !write_factor) break;
            const int T = m > 0 ? (m + CU_TILE - 1) / CU_TILE : 1;       // with write_factor tile (0,0) always runs: it exports the diagonal block
            k_chol_update<<<dim3(T, T), 256, 0, stream>>>(A_dev, work_dev, Dblk, n, j0, fail_dev, write_factor);
        }
        k_chol_backsolve<<<1, 1024, 2 * (size_t)n * sizeof(double), stream>>>(A_dev, work_dev, Dblk, Dinv, x_dev, n, fail_dev);
    }
    SBA_CUDA(cudaGetLastError());
    return SBA_OK;
}

int chol_stage_clocks(long long* out16)
{
    SBA_CUDA(cudaMemcpyFromSymbol(out16, g_chol_clk, 16 * sizeof(long long)));
    return SBA_OK;
}

}  // namespace sba
