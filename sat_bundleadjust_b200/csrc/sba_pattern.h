// Pattern-major layout of a bundle-adjustment problem (host side, pure C++, no CUDA).
//
// Satellite bundle adjustment has few cameras (the reference adjusts a handful of views per run: ba_bruteforce on all
// images of a date, ba_sequential / ba_global on short windows of dates, bundle_adjust/ba_timeseries.py:516-550) and
// many tracks, so the SET of cameras that sees a track -- its visibility pattern -- repeats thousands of times.  The
// device layout exploits that: tracks are reordered so that tracks with the same pattern are adjacent.  Inside such
// a run every track has the same length L and position k of every track is the same camera, hence
//   * cam_ind / pts_ind are not read in the hot loops at all (lane -> (track slot, position) is arithmetic),
//   * a lane works for ONE camera during a whole run: camera blocks U_j, g_j and the Schur products
//     Z_a Z_b^T of a fixed camera pair accumulate in registers and are flushed once per work unit,
//   * the observations of a run are one contiguous, coalesced span.
// The permutation is internal: every C-ABI entry point takes and returns the reference's order
// (bundle_adjust/ba_params.py:138-149: by track, camera ascending inside a track).
//
//   run   = maximal set of tracks with identical (frozen flag, camera set); frozen tracks come first, tracks without
//           observations last
//   tile  = floor(32 / L) tracks of a run = one pass of a warp, lane = slot * L + position
//   unit  = up to `unit_tiles` consecutive tiles of a run = the work item of one warp (static assignment, so that
//           every reduction has a fixed order and results are reproducible bit for bit)
//   CTA c owns units [cta_unit0[c], cta_unit0[c+1]); its warps take them round-robin.
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

namespace sba {

struct PUnit {            // 32 bytes, read by every lane of the warp that owns the unit
    int trk0;             // first internal track
    int ntrk;             // tracks in the unit
    int obs0;             // first internal observation (= internal track_ptr[trk0])
    int L;                // observations per track (1..32)
    int pat;              // offset of the camera list in pat_cams
    int pts_free;         // 0: the points of this run are frozen (the caller's first n_pts_fix tracks)
    int pad0, pad1;
};

struct PatternLayout {
    bool ok = false;
    std::string why;                   // why the layout does not apply (generic engine is used instead)
    std::vector<int> trk_new2old;      // (N) internal track -> caller's track (tracks without observations last)
    std::vector<int> obs_new2old;      // (K) internal observation -> caller's observation
    std::vector<int> track_ptr;        // (N+1) internal track offsets
    std::vector<PUnit> units;
    std::vector<int> pat_cams;
    std::vector<int> cta_unit0;        // (n_cta + 1)
    int n_cta = 0, Lmax = 0, n_runs = 0, unit_tiles = 0;
    int n_frozen_tracks = 0;           // frozen tracks with observations = internal tracks [0, n_frozen_tracks)
    long long n_tiles = 0;
};

// cam: (K) camera of every observation, caller's order; track_ptr_old: (N+1) first observation of every track.
// n_pts_fix: the caller's first n_pts_fix tracks are frozen.
inline void build_pattern_layout(const int* cam, const int* track_ptr_old, long long K, int M, int N, int n_pts_fix,
                                 int n_cta, int warps_per_cta, PatternLayout& out)
{
    out = PatternLayout();
    if (M > 64) { out.why = "more than 64 cameras"; return; }
    if (N < 1 || K < 1) { out.why = "empty"; return; }
    // key = camera bit set (+ frozen flag as bit 64 handled by a separate byte)
    std::vector<uint64_t> key(N);
    std::vector<unsigned char> aux(N);          // bit 0: free (frozen tracks come first), bit 1: no observations (sorted last)
    int Lmax = 0;
    for (int i = 0; i < N; ++i) {
        const int a0 = track_ptr_old[i], a1 = track_ptr_old[i + 1];
        uint64_t m = 0;
        int prev = -1;
        for (int a = a0; a < a1; ++a) {
            if (cam[a] <= prev) { out.why = "cameras not strictly ascending inside a track"; return; }
            prev = cam[a];
            m |= (uint64_t)1 << cam[a];
        }
        key[i] = m;
        aux[i] = (unsigned char)((i < n_pts_fix ? 0 : 1) | (a1 == a0 ? 2 : 0));
        if (i < n_pts_fix && a1 > a0) out.n_frozen_tracks++;
        Lmax = std::max(Lmax, a1 - a0);
    }
    if (Lmax > 32) { out.why = "a track has more than 32 observations"; return; }
    out.Lmax = Lmax;
    // stable LSD radix sort of the tracks by (aux, key): byte passes over the key bytes in use, then aux
    std::vector<int> order(N), tmp(N);
    for (int i = 0; i < N; ++i) order[i] = i;
    const int key_bytes = (M + 7) / 8;
    for (int pass = 0; pass <= key_bytes; ++pass) {
        size_t hist[257] = {0};
        auto digit = [&](int i) -> unsigned { return pass < key_bytes ? (unsigned)((key[i] >> (8 * pass)) & 0xff) : aux[i]; };
        for (int i = 0; i < N; ++i) hist[digit(i) + 1]++;
        bool single = false;
        for (int d = 0; d < 256; ++d) if (hist[d + 1] == (size_t)N) single = true;
        if (single) continue;
        for (int d = 0; d < 256; ++d) hist[d + 1] += hist[d];
        for (int t = 0; t < N; ++t) { const int i = order[t]; tmp[hist[digit(i)]++] = i; }
        order.swap(tmp);
    }
    out.trk_new2old = order;
    out.track_ptr.assign(N + 1, 0);
    for (int t = 0; t < N; ++t) out.track_ptr[t + 1] = out.track_ptr[t] + (track_ptr_old[order[t] + 1] - track_ptr_old[order[t]]);
    out.obs_new2old.resize((size_t)K);
    for (int t = 0; t < N; ++t) {
        const int o = order[t], a0 = track_ptr_old[o], L = track_ptr_old[o + 1] - a0, b0 = out.track_ptr[t];
        for (int k = 0; k < L; ++k) out.obs_new2old[b0 + k] = a0 + k;
    }
    // runs -> tiles
    struct Run { int trk0, ntrk, L, pts_free; };
    std::vector<Run> runs;
    long long tiles = 0;
    for (int t = 0; t < N;) {
        const int o = order[t];
        if (aux[o] & 2) break;                       // tracks without observations: no work
        int e = t + 1;
        while (e < N && key[order[e]] == key[o] && aux[order[e]] == aux[o]) ++e;
        const int L = track_ptr_old[o + 1] - track_ptr_old[o];
        runs.push_back({t, e - t, L, (aux[o] & 1) ? 1 : 0});
        const int T = 32 / L;
        tiles += (e - t + T - 1) / T;
        t = e;
    }
    out.n_runs = (int)runs.size();
    out.n_tiles = tiles;
    // unit size: ~6 units per warp for balance, at most 16 tiles (flush cost amortised), at least 1
    long long ut = tiles / ((long long)n_cta * warps_per_cta * 6);
    out.unit_tiles = (int)std::max<long long>(1, std::min<long long>(16, ut));
    std::vector<long long> unit_cost;
    for (const Run& r : runs) {
        const int pat = (int)out.pat_cams.size();
        const int o = order[r.trk0];
        for (int a = track_ptr_old[o]; a < track_ptr_old[o + 1]; ++a) out.pat_cams.push_back(cam[a]);
        const int T = 32 / r.L, per_unit = T * out.unit_tiles;
        for (int s = 0; s < r.ntrk; s += per_unit) {
            PUnit u;
            u.trk0 = r.trk0 + s; u.ntrk = std::min(per_unit, r.ntrk - s); u.obs0 = out.track_ptr[u.trk0];
            u.L = r.L; u.pat = pat; u.pts_free = r.pts_free; u.pad0 = u.pad1 = 0;
            out.units.push_back(u);
            unit_cost.push_back((u.ntrk + T - 1) / T);
        }
    }
    // contiguous split of the units over the CTAs by cumulative tile count
    out.n_cta = n_cta;
    out.cta_unit0.assign(n_cta + 1, 0);
    long long acc = 0;
    size_t u = 0;
    for (int c = 0; c < n_cta; ++c) {
        out.cta_unit0[c] = (int)u;
        const long long target = tiles * (c + 1) / n_cta;
        while (u < out.units.size() && (acc + unit_cost[u] <= target || c == n_cta - 1)) { acc += unit_cost[u]; ++u; }
    }
    out.cta_unit0[n_cta] = (int)out.units.size();
    out.ok = true;
}

}  // namespace sba
