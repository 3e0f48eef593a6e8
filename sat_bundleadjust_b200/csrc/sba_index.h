// Host-side index construction of a problem (pure C++, no CUDA): int64 -> int32 indices, track offsets, the
// camera-major permutation, camera-major work chunks and the warp tiles of the track-major kernels.
// Input contract = the reference's observation layout (bundle_adjust/ba_params.py:138-149): observations sorted by
// track (pts_ind non-decreasing), cameras ascending inside a track.  Compiled into libsba_b200.so (sba_ba.cu) and,
// for the CPU tests, into tests/host_harness.
#pragma once
#include <algorithm>
#include <cstdint>
#include <thread>
#include <vector>

namespace sba {

struct HostIndex {
    std::vector<int> cam, pts;          // (K) int32 copies of cam_ind / pts_ind
    std::vector<int> track_ptr;         // (N+1) first observation of every track (empty tracks: offset of the next one)
    std::vector<int> cam_cnt;           // (M+1) prefix of the observations per camera
    std::vector<int> cm_obs;            // (K) camera-major permutation: observations of camera 0 in input order, then camera 1, ...
    std::vector<int> ch_cam, ch_beg, ch_end, first_chunk;   // camera-major work items of <= chunk observations; (M+1) first item per camera
    std::vector<int> tile_obs;          // (T+1) warp tiles: runs of whole tracks with <= 32 observations; a longer track is its own tile
};

// returns 0, 1 (index out of range) or 2 (pts_ind decreasing)
// tracks_only: stop after the int32 copies and the track offsets (the pattern engine needs nothing else)
inline int build_host_index(const int64_t* cam_ind, const int64_t* pts_ind, int64_t K, int M, int N, int chunk, int max_threads,
                            HostIndex& h, bool tracks_only = false)
{
    h.cam.assign(K, 0); h.pts.assign(K, 0); h.track_ptr.assign(N + 1, 0); h.cam_cnt.assign(M + 1, 0);
    const int hw = (int)std::thread::hardware_concurrency();
    const int nthr = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)max_threads, (int64_t)(hw > 0 ? hw : 1), K / 65536 + 1}));
    std::vector<std::vector<int>> thr_cnt(nthr, std::vector<int>(M, 0));
    std::vector<int> thr_err(nthr, 0);
    std::vector<char> seen(N + 1, 0);
    auto range_of = [&](int t, int64_t& a0, int64_t& a1) { a0 = K * t / nthr; a1 = K * (t + 1) / nthr; };
    auto run = [&](auto&& body) {
        if (nthr == 1) { body(0); return; }
        std::vector<std::thread> th;
        for (int t = 0; t < nthr; ++t) th.emplace_back(body, t);
        for (auto& t : th) t.join();
    };
    run([&](int t) {
        int64_t a0, a1;
        range_of(t, a0, a1);
        std::vector<int>& cnt = thr_cnt[t];
        for (int64_t a = a0; a < a1; ++a) {
            const int64_t c = cam_ind[a], tr = pts_ind[a];
            if (c < 0 || c >= M || tr < 0 || tr >= N) { thr_err[t] = 1; return; }
            if (a > 0 && tr < pts_ind[a - 1]) { thr_err[t] = 2; return; }
            h.cam[a] = (int)c; h.pts[a] = (int)tr;
            cnt[c]++;
            // first observation of a track: its offset (tracks are contiguous runs)
            if (a == 0 || tr != pts_ind[a - 1]) { h.track_ptr[tr] = (int)a; seen[tr] = 1; }
        }
    });
    for (int t = 0; t < nthr; ++t) if (thr_err[t]) return thr_err[t];
    h.track_ptr[N] = (int)K;
    for (int i = N - 1; i >= 0; --i) if (!seen[i]) h.track_ptr[i] = h.track_ptr[i + 1];
    if (tracks_only) return 0;
    for (int j = 0; j < M; ++j) {
        int tot = 0;
        for (int t = 0; t < nthr; ++t) { const int c = thr_cnt[t][j]; thr_cnt[t][j] = tot; tot += c; }   // thread t's start inside camera j
        h.cam_cnt[j + 1] = h.cam_cnt[j] + tot;
    }
    h.cm_obs.assign(K, 0);
    run([&](int t) {
        int64_t a0, a1;
        range_of(t, a0, a1);
        std::vector<int> fill(M);
        for (int j = 0; j < M; ++j) fill[j] = h.cam_cnt[j] + thr_cnt[t][j];
        for (int64_t a = a0; a < a1; ++a) h.cm_obs[fill[h.cam[a]]++] = (int)a;
    });
    h.ch_cam.clear(); h.ch_beg.clear(); h.ch_end.clear(); h.first_chunk.assign(M + 1, 0);
    for (int j = 0; j < M; ++j) {
        h.first_chunk[j] = (int)h.ch_cam.size();
        for (int b = h.cam_cnt[j]; b < h.cam_cnt[j + 1]; b += chunk) {
            h.ch_cam.push_back(j); h.ch_beg.push_back(b); h.ch_end.push_back(std::min(b + chunk, h.cam_cnt[j + 1]));
        }
    }
    h.first_chunk[M] = (int)h.ch_cam.size();
    h.tile_obs.assign(1, 0);
    int cur = 0;   // observations in the open tile
    for (int i = 0; i < N; ++i) {
        const int L = h.track_ptr[i + 1] - h.track_ptr[i];
        if (L == 0) continue;
        if (cur > 0 && cur + L > 32) { h.tile_obs.push_back(h.track_ptr[i]); cur = 0; }
        cur += L;
        if (cur >= 32) { h.tile_obs.push_back(h.track_ptr[i + 1]); cur = 0; }
    }
    if (cur > 0) h.tile_obs.push_back(h.track_ptr[N]);
    return 0;
}

}  // namespace sba
