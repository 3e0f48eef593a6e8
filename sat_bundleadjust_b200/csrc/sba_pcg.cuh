// G5: block-Jacobi preconditioned conjugate gradients on the reduced camera system, matrix-free -- the path for
// time-series scale problems (hundreds of cameras: BASELINE config 4, the reference bounds this scale with ba_global /
// ba_sequential, bundle_adjust/ba_timeseries.py:516-550, and solves it with LSMR, scipy/optimize/_lsq/trf.py:485-500).
//
//   S = U' - W V'^-1 W^T  is never formed.  With Z_a = (Jc_a^T Jp_a) G_i^T per observation (k_point_prep), G_i the inverse
//   Cholesky factor of the damped point block:      S v = U' v - sum_tracks sum_a Z_a ( sum_b Z_b^T v_cam(b) )
//   k_pcg_tracks    track-major (warp tiles): s_i = sum_b Z_b^T v_cam(b)                         (3 doubles per track)
//   k_pcg_cameras   camera-major (block-uniform camera, register accumulators, no atomics): partial sums of Z_a s_i
//   k_pcg_sum       per camera: fixed-order sum of the chunk partials
//   k_pcg_update    one CTA: the CG recurrences on the (M n_params)-vectors, preconditioner solves, stopping test --
//                   all scalars stay on the device; the host looks at the `done` flag every few iterations
//   preconditioner  M_j = S_jj = U'_j - sum_{a in cam j} Z_a Z_a^T  (k_pcg_diag + k_pcg_sum), factored per camera
// Multi-GPU: tracks are sharded, so the sums over observations are partial -> one all-reduce of an (M n_params)-vector
// per CG iteration (+ one of the diagonal blocks and the right-hand side per trust-region iteration).
// No obs_of table and no pair lists are needed on this path.  Included by sba_ba.cu only.
#pragma once
#include "sba_kernels.cuh"

namespace sba {

// scalar slots of the CG state (device, after the vectors): [0] rz, [1] rz0, [2] done, [3] iterations, [4] |r|_M / |r0|_M
constexpr int PCG_SCAL = 8;

// ---- s_i = sum_b Z_b^T v_cam(b), warp tiles as in k_backsub ----------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(TPB)
k_pcg_tracks(ObsArrays o, const double* __restrict__ Zin, const double* __restrict__ v, const double* __restrict__ cg,
             double* __restrict__ s_out)
{
    constexpr int ZS = NC * 3, ZP = ZS + 1;
    __shared__ double sv[WPB][3][32];
    __shared__ double sZ[WPB][32 * ZP];
    if (cg[2] != 0.0) return;                         // converged: the remaining launches of the batch are no-ops
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int ob, nobs;
    warp_tile(o, ob, nobs);
    if (nobs == 0) return;
    if (nobs <= 32) {
        const double* src = Zin + (size_t)ob * ZS;
        for (int t = lane; t < nobs * ZS; t += 32) sZ[warp][(t / ZS) * ZP + (t % ZS)] = src[t];
        __syncwarp();
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        int i = -1, a = -1;
        if (lane < nobs) {
            a = ob + lane;
            i = o.pts_ind[a];
            const int j = o.cam_ind[a];
            const double* z = &sZ[warp][lane * ZP];
#pragma unroll
            for (int r = 0; r < NC; ++r) {
                const double dc = v[(size_t)j * NC + r];
                s0 += z[3 * r] * dc; s1 += z[3 * r + 1] * dc; s2 += z[3 * r + 2] * dc;
            }
        }
        sv[warp][0][lane] = s0; sv[warp][1][lane] = s1; sv[warp][2][lane] = s2;
        __syncwarp();
        if (lane < nobs) {
            const int beg = o.track_ptr[i];
            if (a == beg) {
                const int L = o.track_ptr[i + 1] - beg;
                double t0 = 0.0, t1 = 0.0, t2 = 0.0;
                for (int m = 0; m < L; ++m) { t0 += sv[warp][0][lane + m]; t1 += sv[warp][1][lane + m]; t2 += sv[warp][2][lane + m]; }
                s_out[3 * (size_t)i] = t0; s_out[3 * (size_t)i + 1] = t1; s_out[3 * (size_t)i + 2] = t2;
            }
        }
    } else {
        const int i = o.pts_ind[ob];
        double s[3] = {0.0, 0.0, 0.0};
        for (int a = ob + lane; a < ob + nobs; a += 32) {
            const int j = o.cam_ind[a];
            const double* z = Zin + (size_t)a * ZS;
#pragma unroll
            for (int r = 0; r < NC; ++r) {
                const double dc = v[(size_t)j * NC + r];
                s[0] += z[3 * r] * dc; s[1] += z[3 * r + 1] * dc; s[2] += z[3 * r + 2] * dc;
            }
        }
        warp_allreduce_sum<3>(s);
        if (lane == 0) { s_out[3 * (size_t)i] = s[0]; s_out[3 * (size_t)i + 1] = s[1]; s_out[3 * (size_t)i + 2] = s[2]; }
    }
}

// ---- camera-major: per chunk of one camera, sum_a Z_a s_i(a)  (NC values) ---------------------------------------------
template <int NC>
__global__ void __launch_bounds__(TPB)
k_pcg_cameras(const int* __restrict__ chunk_beg, const int* __restrict__ chunk_end, const int* __restrict__ cm_obs,
              const int* __restrict__ cm_pts, const double* __restrict__ Zin, const double* __restrict__ s_in,
              const double* __restrict__ cg, double* __restrict__ partials)
{
    __shared__ double sm[NC * (TPB / 32)];
    if (cg[2] != 0.0) return;
    const int ch = blockIdx.x;
    double acc[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) acc[k] = 0.0;
    for (int t = chunk_beg[ch] + threadIdx.x; t < chunk_end[ch]; t += TPB) {
        const int a = cm_obs[t], i = cm_pts[t];
        const double s0 = s_in[3 * (size_t)i], s1 = s_in[3 * (size_t)i + 1], s2 = s_in[3 * (size_t)i + 2];
        const double* z = Zin + (size_t)a * NC * 3;
#pragma unroll
        for (int r = 0; r < NC; ++r) acc[r] += z[3 * r] * s0 + z[3 * r + 1] * s1 + z[3 * r + 2] * s2;
    }
    const double tot = block_reduce_sum<NC, TPB>(acc, sm);
    if (threadIdx.x < NC) partials[(size_t)ch * NC + threadIdx.x] = tot;
}

// ---- camera-major: per chunk, the diagonal Schur block sum_a Z_a Z_a^T (lower triangle) and sum_a Z_a q_i -------------
template <int NC>
__global__ void __launch_bounds__(TPB)
k_pcg_diag(const int* __restrict__ chunk_beg, const int* __restrict__ chunk_end, const int* __restrict__ cm_obs,
           const int* __restrict__ cm_pts, const double* __restrict__ Zin, const double* __restrict__ q,
           double* __restrict__ partials)
{
    constexpr int NU = NC * (NC + 1) / 2, NV = NU + NC;
    __shared__ double sm[NV * (TPB / 32)];
    const int ch = blockIdx.x;
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0.0;
    for (int t = chunk_beg[ch] + threadIdx.x; t < chunk_end[ch]; t += TPB) {
        const int a = cm_obs[t], i = cm_pts[t];
        const double q0 = q[3 * (size_t)i], q1 = q[3 * (size_t)i + 1], q2 = q[3 * (size_t)i + 2];
        double z[NC * 3];
        const double* zp = Zin + (size_t)a * NC * 3;
#pragma unroll
        for (int k = 0; k < NC * 3; ++k) z[k] = zp[k];
        int k = 0;
#pragma unroll
        for (int r = 0; r < NC; ++r) {
#pragma unroll
            for (int c = 0; c <= r; ++c) { acc[k] += z[3 * r] * z[3 * c] + z[3 * r + 1] * z[3 * c + 1] + z[3 * r + 2] * z[3 * c + 2]; ++k; }
        }
#pragma unroll
        for (int r = 0; r < NC; ++r) acc[NU + r] += z[3 * r] * q0 + z[3 * r + 1] * q1 + z[3 * r + 2] * q2;
    }
    const double tot = block_reduce_sum<NV, TPB>(acc, sm);
    if (threadIdx.x < NV) partials[(size_t)ch * NV + threadIdx.x] = tot;
}

// per camera: fixed-order sum of its chunks' partials (NV values per chunk) -> out[j * NV + k]
__global__ void __launch_bounds__(128)
k_pcg_sum(const double* __restrict__ partials, const int* __restrict__ first_chunk, int NV, const double* __restrict__ cg,
          int check_done, double* __restrict__ out)
{
    if (check_done && cg[2] != 0.0) return;
    const int j = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = first_chunk[j], c1 = first_chunk[j + 1];
    for (int k = warp; k < NV; k += 4) {
        double s = 0.0;
        for (int ch = c0 + lane; ch < c1; ch += 32) s += partials[(size_t)ch * NV + k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) out[(size_t)j * NV + k] = s;
    }
}

// ---- preconditioner and right-hand side from the (all-reduced) diagonal sums; CG start --------------------------------
//   Mblk_j = U_j + reg D_j^2 - sum Z Z^T  (Cholesky factor L_j stored, lower, row-major NC x NC), identity for frozen cameras
//   rhs_j  = -g_j + sum Z q ;  x = 0, r = rhs, z = M^-1 r, p = z, rz = r.z
// vectors in `vec`: x | r | z | p | Ap  (ns each), then PCG_SCAL scalars.  One CTA.
template <int NC>
__device__ __forceinline__ void pcg_block_solve(const double* __restrict__ Lf, const double* rr, double* zz)
{
    double y[NC];
#pragma unroll
    for (int r = 0; r < NC; ++r) {
        double t = rr[r];
#pragma unroll
        for (int c = 0; c < r; ++c) t -= Lf[r * NC + c] * y[c];
        y[r] = t / Lf[r * NC + r];
    }
#pragma unroll
    for (int r = NC - 1; r >= 0; --r) {
        double t = y[r];
#pragma unroll
        for (int c = r + 1; c < NC; ++c) t -= Lf[c * NC + r] * zz[c];
        zz[r] = t / Lf[r * NC + r];
    }
}

template <int NC>
__global__ void __launch_bounds__(1024)
k_pcg_init(const double* __restrict__ diag_sums, const double* __restrict__ camsys, const double* __restrict__ sinv,
           const double* __restrict__ scal, int M, int n_cam_fix, double* __restrict__ Lfac, double* __restrict__ vec, double* fail)
{
    constexpr int NU = NC * (NC + 1) / 2, NV = NU + NC;
    const int ns = M * NC;
    double *x = vec, *r = vec + ns, *z = vec + 2 * ns, *pv = vec + 3 * ns, *cg = vec + 5 * (size_t)ns;
    const double reg = scal[SC_REG];
    __shared__ double s_red[32];
    double rz = 0.0;
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
        double A[NC * NC], rr[NC], zz[NC];
        const double* ds = diag_sums + (size_t)j * NV;
        if (j < n_cam_fix) {
#pragma unroll
            for (int e = 0; e < NC * NC; ++e) A[e] = (e / NC == e % NC) ? 1.0 : 0.0;
#pragma unroll
            for (int e = 0; e < NC; ++e) rr[e] = 0.0;
        } else {
            int k = 0;
#pragma unroll
            for (int rw = 0; rw < NC; ++rw) {
#pragma unroll
                for (int c = 0; c <= rw; ++c) {
                    double v = camsys[(size_t)j * NC * NC + rw * NC + c] - ds[k];
                    if (rw == c) { const double si = sinv[(size_t)j * NC + rw]; v += reg * si * si; }
                    A[rw * NC + c] = v;
                    ++k;
                }
            }
#pragma unroll
            for (int e = 0; e < NC; ++e) rr[e] = -camsys[(size_t)M * NC * NC + (size_t)j * NC + e] + ds[NU + e];
            // in-place Cholesky of the lower triangle
            bool ok = true;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                double d = A[c * NC + c];
#pragma unroll
                for (int k2 = 0; k2 < c; ++k2) d -= A[c * NC + k2] * A[c * NC + k2];
                ok = ok && d > 0.0;
                const double l = sqrt(d > 0.0 ? d : 1.0);
                A[c * NC + c] = l;
#pragma unroll
                for (int rw = c + 1; rw < NC; ++rw) {
                    double t = A[rw * NC + c];
#pragma unroll
                    for (int k2 = 0; k2 < c; ++k2) t -= A[rw * NC + k2] * A[c * NC + k2];
                    A[rw * NC + c] = t / l;
                }
            }
            if (!ok) *fail = (double)(j * NC + 1);
        }
#pragma unroll
        for (int e = 0; e < NC * NC; ++e) Lfac[(size_t)j * NC * NC + e] = A[e];
        pcg_block_solve<NC>(A, rr, zz);
#pragma unroll
        for (int e = 0; e < NC; ++e) {
            const size_t idx = (size_t)j * NC + e;
            x[idx] = 0.0; r[idx] = rr[e]; z[idx] = zz[e]; pv[idx] = zz[e];
            rz += rr[e] * zz[e];
        }
    }
    // block sum of rz (fixed order)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rz += __shfl_down_sync(0xffffffffu, rz, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = rz;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int wq = 0; wq < (int)(blockDim.x >> 5); ++wq) t += s_red[wq];
        cg[0] = t; cg[1] = t; cg[2] = (t > 0.0) ? 0.0 : 1.0; cg[3] = 0.0; cg[4] = 1.0;
    }
}

// ---- one CG step on the device: Ap = U' p - (all-reduced) W-term; alpha, x, r, z, beta, p; stopping test ---------------
template <int NC>
__global__ void __launch_bounds__(1024)
k_pcg_update(const double* __restrict__ wterm, const double* __restrict__ camsys, const double* __restrict__ sinv,
             const double* __restrict__ scal, const double* __restrict__ Lfac, int M, int n_cam_fix, double tol, int max_it,
             double* __restrict__ vec)
{
    const int ns = M * NC;
    double *x = vec, *r = vec + ns, *z = vec + 2 * ns, *pv = vec + 3 * ns, *Ap = vec + 4 * (size_t)ns, *cg = vec + 5 * (size_t)ns;
    if (cg[2] != 0.0) return;
    const double reg = scal[SC_REG];
    __shared__ double s_red[32];
    __shared__ double s_bc[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    // Ap and p.Ap (one camera per thread)
    double pap = 0.0;
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
        double pj[NC];
#pragma unroll
        for (int e = 0; e < NC; ++e) pj[e] = pv[(size_t)j * NC + e];
#pragma unroll
        for (int rw = 0; rw < NC; ++rw) {
            double t;
            if (j < n_cam_fix) {
                t = pj[rw];
            } else {
                t = 0.0;
#pragma unroll
                for (int c = 0; c < NC; ++c) t += camsys[(size_t)j * NC * NC + rw * NC + c] * pj[c];
                const double si = sinv[(size_t)j * NC + rw];
                t += reg * si * si * pj[rw] - wterm[(size_t)j * NC + rw];
            }
            Ap[(size_t)j * NC + rw] = t;
            pap += pj[rw] * t;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pap += __shfl_down_sync(0xffffffffu, pap, o);
    if (lane == 0) s_red[warp] = pap;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int wq = 0; wq < nw; ++wq) t += s_red[wq];
        s_bc[0] = t;
    }
    __syncthreads();
    const double rz = cg[0];
    const double alpha = s_bc[0] > 0.0 ? rz / s_bc[0] : 0.0;
    double rz_new = 0.0;
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
        double rr[NC], zz[NC];
#pragma unroll
        for (int e = 0; e < NC; ++e) {
            const size_t idx = (size_t)j * NC + e;
            x[idx] += alpha * pv[idx];
            rr[e] = r[idx] - alpha * Ap[idx];
            r[idx] = rr[e];
        }
        pcg_block_solve<NC>(Lfac + (size_t)j * NC * NC, rr, zz);
#pragma unroll
        for (int e = 0; e < NC; ++e) { z[(size_t)j * NC + e] = zz[e]; rz_new += rr[e] * zz[e]; }
    }
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rz_new += __shfl_down_sync(0xffffffffu, rz_new, o);
    if (lane == 0) s_red[warp] = rz_new;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int wq = 0; wq < nw; ++wq) t += s_red[wq];
        s_bc[1] = t;
    }
    __syncthreads();
    const double beta = rz > 0.0 ? s_bc[1] / rz : 0.0;
    for (int e = threadIdx.x; e < ns; e += blockDim.x) pv[e] = z[e] + beta * pv[e];
    if (threadIdx.x == 0) {
        const double it = cg[3] + 1.0;
        const double rel = cg[1] > 0.0 ? sqrt(fmax(s_bc[1], 0.0) / cg[1]) : 0.0;
        cg[0] = s_bc[1]; cg[3] = it; cg[4] = rel;
        if (rel <= tol || it >= (double)max_it || !(s_bc[0] > 0.0)) cg[2] = 1.0;
    }
}

}  // namespace sba
