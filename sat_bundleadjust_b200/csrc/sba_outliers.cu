// Outlier detection between the two bundle-adjustment passes: the device part of
// bundle_adjust/ba_outliers.py:14-58 (get_elbow_value) and :112-153 (compute_obs_to_remove).
//
// The reference sorts the reprojection errors of every camera (np.sort, one camera at a time, inside a Python loop
// over boolean masks of all K observations) and takes the "elbow" of the sorted curve: the sample furthest from
// the chord between its first and last point.  Here ONE stable LSD radix sort orders all K observations by
// (camera, error) -- non-negative IEEE doubles order like their bit patterns -- and one CTA per camera finds the
// elbow.  The distance formula is evaluated in the reference's operation order with explicitly rounded
// multiplications and additions (no FMA contraction), so the arg-max and hence the thresholds are bit-identical;
// the O(n_cam) scalar logic that follows (np.percentile interpolation, max, np.round) stays in numpy on the host
// (sat_bundleadjust_b200/ba_outliers.py).
#include <cstdint>
#include <vector>

#include "sba_internal.cuh"

namespace sba {

namespace {

constexpr int RS_THREADS = 256, RS_WARPS = RS_THREADS / 32, RS_ITEMS = 8;     // 2048 keys per CTA
constexpr int RS_TILE = RS_THREADS * RS_ITEMS, RS_BINS = 256;

struct DevBufO {
    void* p = nullptr;
    ~DevBufO() { if (p) cudaFree(p); }
    int alloc(size_t bytes) { SBA_CUDA(cudaMalloc(&p, bytes ? bytes : 1)); return SBA_OK; }
    template <typename T> T* as() { return (T*)p; }
};

// sort key of observation k: high word = camera, low 64 bits = error bits (sorted as two 64-bit words: the error
// digits first, then the camera digits)
__global__ void k_outlier_keys(const double* __restrict__ err, const int* __restrict__ cam_ind, long long K,
                               unsigned long long* __restrict__ key_err, unsigned int* __restrict__ key_cam,
                               unsigned int* __restrict__ idx, unsigned int* __restrict__ digit_used)
{
    // digit_used[d * 256 + v] != 0 when some key has value v in byte d (d < 8: error bytes, 8..9: camera bytes);
    // a byte in which all keys agree needs no pass
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < K; k += (long long)gridDim.x * blockDim.x) {
        double e = err[k];
        if (!(e == e)) e = __longlong_as_double(0x7ff8000000000000LL);      // canonical NaN: sorts last, like np.sort
        else if (e == 0.0) e = 0.0;                                         // -0.0 -> +0.0
        const unsigned long long b = (unsigned long long)__double_as_longlong(e);
        const unsigned int c = (unsigned int)cam_ind[k];
        key_err[k] = b; key_cam[k] = c; idx[k] = (unsigned int)k;
#pragma unroll
        for (int d = 0; d < 8; ++d) digit_used[d * 256 + (int)((b >> (8 * d)) & 255ull)] = 1u;
        digit_used[8 * 256 + (int)(c & 255u)] = 1u;
        digit_used[9 * 256 + (int)((c >> 8) & 255u)] = 1u;
    }
}

__device__ __forceinline__ unsigned int digit_of(unsigned long long ke, unsigned int kc, int d)
{
    return d < 8 ? (unsigned int)((ke >> (8 * d)) & 255ull) : ((kc >> (8 * (d - 8))) & 255u);
}

// One radix pass, two launches of the same body.  Warp w of a CTA owns the 256 consecutive keys
// [tile + 256 w, tile + 256 (w+1)), visited 32 at a time, so "earlier in the input" = (warp, iteration, lane)
// lexicographic order and the pass is stable.  SCATTER == false: per-CTA digit histogram -> hist[bin * nblk + blk].
// SCATTER == true: hist holds the exclusive scan of that table; every key moves to
//     scan[bin][blk] + (keys of the same bin in earlier warps of the CTA) + (rank inside the warp's chunk).
template <bool SCATTER>
__global__ void __launch_bounds__(RS_THREADS)
k_radix_pass(const unsigned long long* __restrict__ ke_in, const unsigned int* __restrict__ kc_in,
             const unsigned int* __restrict__ idx_in, long long K, int d, unsigned int* __restrict__ hist, int nblk,
             unsigned long long* __restrict__ ke_out, unsigned int* __restrict__ kc_out, unsigned int* __restrict__ idx_out)
{
    __shared__ unsigned int cnt[RS_WARPS][RS_BINS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int b = tid; b < RS_WARPS * RS_BINS; b += RS_THREADS) (&cnt[0][0])[b] = 0u;
    __syncthreads();
    const long long base = (long long)blockIdx.x * RS_TILE + warp * (RS_TILE / RS_WARPS);
    unsigned long long ke[RS_ITEMS];
    unsigned int kc[RS_ITEMS], id[RS_ITEMS], rank[RS_ITEMS], dg[RS_ITEMS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const long long k = base + i * 32 + lane;
        const bool ok = k < K;
        ke[i] = ok ? ke_in[k] : 0ull; kc[i] = ok ? kc_in[k] : 0u; id[i] = ok ? idx_in[k] : 0u;
    }
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const long long k = base + i * 32 + lane;
        const bool ok = k < K;
        dg[i] = ok ? digit_of(ke[i], kc[i], d) : 0xffffffffu;               // out-of-range lanes form their own group
        const unsigned int peers = __match_any_sync(0xffffffffu, dg[i]);
        const unsigned int before = __popc(peers & ((1u << lane) - 1u));
        unsigned int start = 0u;
        if (ok) start = cnt[warp][dg[i]];
        __syncwarp();
        if (ok && before == 0u) cnt[warp][dg[i]] = start + __popc(peers);  // the group's first lane publishes the new count
        __syncwarp();
        rank[i] = start + before;
    }
    __syncthreads();
    // per bin: exclusive prefix over the warps of this CTA (thread b owns bin b)
    unsigned int total = 0u;
    for (int w = 0; w < RS_WARPS; ++w) {
        const unsigned int c = cnt[w][tid];
        cnt[w][tid] = total;
        total += c;
    }
    if (!SCATTER) {
        hist[(size_t)tid * nblk + blockIdx.x] = total;
        return;
    }
    __shared__ unsigned int gbase[RS_BINS];
    gbase[tid] = hist[(size_t)tid * nblk + blockIdx.x];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const long long k = base + i * 32 + lane;
        if (k < K) {
            const unsigned int pos = gbase[dg[i]] + cnt[warp][dg[i]] + rank[i];
            ke_out[pos] = ke[i]; kc_out[pos] = kc[i]; idx_out[pos] = id[i];
        }
    }
}

// exclusive scan of `n` counters in place, one CTA (n = 256 * number of tiles: 62 k entries at K = 5e5)
__global__ void __launch_bounds__(1024) k_scan_exclusive(unsigned int* __restrict__ a, int n)
{
    __shared__ unsigned int warp_sum[32];
    __shared__ unsigned int carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0u;
    __syncthreads();
    for (int base = 0; base < n; base += 1024 * 4) {
        unsigned int v[4], s = 0u;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = base + tid * 4 + u;
            v[u] = i < n ? a[i] : 0u;
            s += v[u];
        }
        unsigned int incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned int w = warp_sum[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warp_sum[lane] = wi - w;                       // exclusive over warps
        }
        __syncthreads();
        unsigned int run = carry + warp_sum[warp] + (incl - s);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = base + tid * 4 + u;
            if (i < n) a[i] = run;
            run += v[u];
        }
        __syncthreads();
        if (tid == 1023) carry = run;
        __syncthreads();
    }
}

// per camera: segment bounds of the sorted keys (cameras are the most significant digits)
__global__ void k_segment_bounds(const unsigned int* __restrict__ kc_sorted, long long K, int M, long long* __restrict__ seg)
{
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k <= K; k += (long long)gridDim.x * blockDim.x) {
        const long long prev = k == 0 ? -1 : (long long)kc_sorted[k - 1];
        const long long cur = k == K ? (long long)M : (long long)kc_sorted[k];
        for (long long c = prev + 1; c <= cur && c <= M; ++c) seg[c] = k;        // seg[c] = first position with camera >= c
    }
}

// One CTA per camera: elbow of the sorted error curve v[0..n) exactly as ba_outliers.py:32-48 computes it
// (operation order kept, every product and sum individually rounded), first arg-max like np.argmax (a NaN distance
// wins over any number, as in numpy).  Also returns v[q_lo], v[q_hi] (for np.percentile on the host) and v[n-1].
__global__ void __launch_bounds__(256)
k_elbow(const unsigned long long* __restrict__ ke_sorted, const long long* __restrict__ seg, const long long* __restrict__ q_lo,
        const long long* __restrict__ q_hi, double* __restrict__ out /* M x 5: elbow, v[q_lo], v[q_hi], v[n-1], argmax */)
{
    const int cam = blockIdx.x, tid = threadIdx.x;
    const long long s0 = seg[cam], n = seg[cam + 1] - s0;
    double* o = out + (size_t)cam * 5;
    if (n <= 0) {
        if (tid == 0) { o[0] = o[1] = o[2] = o[3] = __longlong_as_double(0x7ff8000000000000LL); o[4] = -1.0; }
        return;
    }
    const unsigned long long* v = ke_sorted + s0;
    const double v0 = __longlong_as_double((long long)v[0]), vl = __longlong_as_double((long long)v[n - 1]);
    // line_vec = all_coord[-1] - all_coord[0]; line_vec_norm = line_vec / sqrt(sum(line_vec ** 2))
    const double lx = (double)(n - 1) - 0.0, ly = __dsub_rn(vl, v0);
    const double nrm = __dsqrt_rn(__dadd_rn(__dmul_rn(lx, lx), __dmul_rn(ly, ly)));
    const double ux = __ddiv_rn(lx, nrm), uy = __ddiv_rn(ly, nrm);
    double best = 0.0;
    long long best_k = -1;
    bool best_nan = false;
    for (long long k = tid; k < n; k += 256) {
        const double fx = (double)k - 0.0, fy = __dsub_rn(__longlong_as_double((long long)v[k]), v0);
        const double sp = __dadd_rn(__dmul_rn(fx, ux), __dmul_rn(fy, uy));
        const double dx = __dsub_rn(fx, __dmul_rn(sp, ux)), dy = __dsub_rn(fy, __dmul_rn(sp, uy));
        const double dist = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        const bool is_nan = !(dist == dist);
        // this thread visits k in increasing order: a later k wins only if strictly better
        if (best_k < 0 || (!best_nan && (is_nan || dist > best))) { best = dist; best_k = k; best_nan = is_nan; }
    }
    __shared__ double s_best[256];
    __shared__ long long s_k[256];
    __shared__ int s_nan[256];
    s_best[tid] = best; s_k[tid] = best_k; s_nan[tid] = best_nan ? 1 : 0;
    __syncthreads();
    for (int o2 = 128; o2 > 0; o2 >>= 1) {
        if (tid < o2) {
            const double b2 = s_best[tid + o2];
            const long long k2 = s_k[tid + o2];
            const int n2 = s_nan[tid + o2];
            const long long k1 = s_k[tid];
            bool take = false;
            if (k2 >= 0) {
                if (k1 < 0) take = true;
                else if (n2 != s_nan[tid]) take = n2 != 0;                          // NaN beats a number
                else if (n2) take = k2 < k1;                                        // first NaN
                else take = (b2 > s_best[tid]) || (b2 == s_best[tid] && k2 < k1);   // first maximum
            }
            if (take) { s_best[tid] = b2; s_k[tid] = k2; s_nan[tid] = n2; }
        }
        __syncthreads();
    }
    if (tid == 0) {
        o[0] = __longlong_as_double((long long)v[s_k[0]]);
        o[1] = __longlong_as_double((long long)v[q_lo[cam] < n ? q_lo[cam] : n - 1]);
        o[2] = __longlong_as_double((long long)v[q_hi[cam] < n ? q_hi[cam] : n - 1]);
        o[3] = vl;
        o[4] = (double)s_k[0];
    }
}

__global__ void k_mark_outliers(const double* __restrict__ err, const int* __restrict__ cam_ind, long long K,
                                const double* __restrict__ thr, unsigned char* __restrict__ remove)
{
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < K; k += (long long)gridDim.x * blockDim.x)
        remove[k] = err[k] > thr[cam_ind[k]] ? 1 : 0;
}

int require_device_o()
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("no CUDA device: sat_bundleadjust_b200 has no CPU fallback");
        return SBA_E_CUDA;
    }
    return SBA_OK;
}

}  // namespace

}  // namespace sba

using namespace sba;

// err[K], cam_ind[K] (host) -> per camera: count, elbow value, the two order statistics np.percentile interpolates
// between (positions q_lo[c], q_hi[c] of the camera's sorted errors, chosen by the caller from the counts it gets from
// sba_outlier_counts or from its own bincount), the maximum, and the arg-max position.  stats is n_cam x 5 doubles.
extern "C" int sba_outlier_elbow(const double* err, const int32_t* cam_ind, int64_t K, int32_t n_cam, const int64_t* q_lo,
                                 const int64_t* q_hi, double* stats, int64_t* counts)
{
    if (!err || !cam_ind || !q_lo || !q_hi || !stats || K < 1 || n_cam < 1 || n_cam > 65535 || K > 0x7fffffffLL) {
        set_error("sba_outlier_elbow: bad argument");
        return SBA_E_INVALID;
    }
    SBA_TRY(require_device_o());
    const int nblk = (int)((K + RS_TILE - 1) / RS_TILE);
    DevBufO d_err, d_cam, ke[2], kc[2], id[2], d_used, d_hist, d_seg, d_qlo, d_qhi, d_out;
    SBA_TRY(d_err.alloc(K * sizeof(double))); SBA_TRY(d_cam.alloc(K * sizeof(int)));
    for (int b = 0; b < 2; ++b) {
        SBA_TRY(ke[b].alloc(K * sizeof(unsigned long long)));
        SBA_TRY(kc[b].alloc(K * sizeof(unsigned int)));
        SBA_TRY(id[b].alloc(K * sizeof(unsigned int)));
    }
    SBA_TRY(d_used.alloc(10 * 256 * sizeof(unsigned int)));
    SBA_TRY(d_hist.alloc((size_t)RS_BINS * nblk * sizeof(unsigned int)));
    SBA_TRY(d_seg.alloc((size_t)(n_cam + 1) * sizeof(long long)));
    SBA_TRY(d_qlo.alloc(n_cam * sizeof(long long))); SBA_TRY(d_qhi.alloc(n_cam * sizeof(long long)));
    SBA_TRY(d_out.alloc((size_t)n_cam * 5 * sizeof(double)));
    SBA_CUDA(cudaMemcpy(d_err.p, err, K * sizeof(double), cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemcpy(d_cam.p, cam_ind, K * sizeof(int), cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemcpy(d_qlo.p, q_lo, n_cam * sizeof(long long), cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemcpy(d_qhi.p, q_hi, n_cam * sizeof(long long), cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemset(d_used.p, 0, 10 * 256 * sizeof(unsigned int)));
    const int g = (int)((K + 255) / 256 < NUM_SMS * 8 ? (K + 255) / 256 : NUM_SMS * 8);
    k_outlier_keys<<<g, 256>>>(d_err.as<double>(), d_cam.as<int>(), K, ke[0].as<unsigned long long>(), kc[0].as<unsigned int>(),
                               id[0].as<unsigned int>(), d_used.as<unsigned int>());
    SBA_CUDA(cudaGetLastError());
    unsigned int used[10 * 256];
    SBA_CUDA(cudaMemcpy(used, d_used.p, sizeof(used), cudaMemcpyDeviceToHost));
    int cur = 0;
    for (int d = 0; d < 10; ++d) {
        int distinct = 0;
        for (int v = 0; v < 256; ++v) distinct += used[d * 256 + v] ? 1 : 0;
        if (distinct <= 1) continue;                     // all keys agree in this byte: the pass would be the identity
        k_radix_pass<false><<<nblk, RS_THREADS>>>(ke[cur].as<unsigned long long>(), kc[cur].as<unsigned int>(),
                                                  id[cur].as<unsigned int>(), K, d, d_hist.as<unsigned int>(), nblk, nullptr,
                                                  nullptr, nullptr);
        k_scan_exclusive<<<1, 1024>>>(d_hist.as<unsigned int>(), RS_BINS * nblk);
        k_radix_pass<true><<<nblk, RS_THREADS>>>(ke[cur].as<unsigned long long>(), kc[cur].as<unsigned int>(),
                                                 id[cur].as<unsigned int>(), K, d, d_hist.as<unsigned int>(), nblk,
                                                 ke[cur ^ 1].as<unsigned long long>(), kc[cur ^ 1].as<unsigned int>(),
                                                 id[cur ^ 1].as<unsigned int>());
        SBA_CUDA(cudaGetLastError());
        cur ^= 1;
    }
    k_segment_bounds<<<g, 256>>>(kc[cur].as<unsigned int>(), K, n_cam, d_seg.as<long long>());
    k_elbow<<<n_cam, 256>>>(ke[cur].as<unsigned long long>(), d_seg.as<long long>(), d_qlo.as<long long>(), d_qhi.as<long long>(),
                            d_out.as<double>());
    SBA_CUDA(cudaGetLastError());
    SBA_CUDA(cudaMemcpy(stats, d_out.p, (size_t)n_cam * 5 * sizeof(double), cudaMemcpyDeviceToHost));
    if (counts) {
        std::vector<long long> seg(n_cam + 1);
        SBA_CUDA(cudaMemcpy(seg.data(), d_seg.p, (size_t)(n_cam + 1) * sizeof(long long), cudaMemcpyDeviceToHost));
        for (int c = 0; c < n_cam; ++c) counts[c] = seg[c + 1] - seg[c];
    }
    return SBA_OK;
}

// remove[k] = err[k] > thr[cam_ind[k]]  (ba_outliers.py:140-146)
extern "C" int sba_outlier_mark(const double* err, const int32_t* cam_ind, int64_t K, int32_t n_cam, const double* thr,
                                uint8_t* remove)
{
    if (!err || !cam_ind || !thr || !remove || K < 1 || n_cam < 1) { set_error("sba_outlier_mark: bad argument"); return SBA_E_INVALID; }
    SBA_TRY(require_device_o());
    DevBufO d_err, d_cam, d_thr, d_rm;
    SBA_TRY(d_err.alloc(K * sizeof(double))); SBA_TRY(d_cam.alloc(K * sizeof(int)));
    SBA_TRY(d_thr.alloc(n_cam * sizeof(double))); SBA_TRY(d_rm.alloc(K));
    SBA_CUDA(cudaMemcpy(d_err.p, err, K * sizeof(double), cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemcpy(d_cam.p, cam_ind, K * sizeof(int), cudaMemcpyHostToDevice));
    SBA_CUDA(cudaMemcpy(d_thr.p, thr, n_cam * sizeof(double), cudaMemcpyHostToDevice));
    const int g = (int)((K + 255) / 256 < NUM_SMS * 8 ? (K + 255) / 256 : NUM_SMS * 8);
    k_mark_outliers<<<g, 256>>>(d_err.as<double>(), d_cam.as<int>(), K, d_thr.as<double>(), d_rm.as<unsigned char>());
    SBA_CUDA(cudaGetLastError());
    SBA_CUDA(cudaMemcpy(remove, d_rm.p, K, cudaMemcpyDeviceToHost));
    return SBA_OK;
}
