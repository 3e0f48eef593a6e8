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
//   tile  = min(floor(32 / L), 16) tracks of a run = one pass of a warp, lane = slot * L + position
//   unit  = the tiles of ONE run inside the contiguous tile range of one warp (every warp of the persistent grid gets
//           the same number of tiles; a range usually touches one or two runs).  Register accumulators live for a
//           whole unit.  The assignment is static, so every reduction has a fixed order and results are reproducible
//           bit for bit.
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <thread>
#include <vector>

namespace sba {

struct PUnit {            // 32 bytes, read by every lane of the warp that owns the unit
    int trk0;             // first internal track
    int ntrk;             // tracks in the unit
    int obs0;             // first internal observation (= internal track_ptr[trk0])
    int L;                // observations per track (1..32)
    unsigned mask_lo;     // camera set of the run, bits 0..31 (position k of a track = k-th set bit) ...
    unsigned mask_hi;     // ... and bits 32..63
    int pts_free;         // 0: the points of this run are frozen (the caller's first n_pts_fix tracks)
    int rec;              // first record of this unit in the Schur record buffer (one record per pass over the unit)
};

// static work assignment for kernels launched with `warps` warps per CTA: warp w of CTA c owns the units
// [warp_unit0[c * warps + w], warp_unit0[c * warps + w + 1])
struct PatternAssignment {
    int warps = 0;
    int n_records = 0;                 // Schur records (sum over units of the passes a unit needs)
    std::vector<PUnit> units;
    std::vector<int> warp_unit0;       // (n_cta * warps + 1)
};

struct PatternLayout {
    bool ok = false;
    std::string why;                   // why the layout does not apply (generic engine is used instead)
    std::vector<int> trk_new2old;      // (N) internal track -> caller's track (tracks without observations last)
    std::vector<int> track_ptr;        // (N+1) internal track offsets
    PatternAssignment light, wide, narrow;   // assignments for the CTA shapes in use (24 / 16 / 12 warps per CTA)
    int n_cta = 0, Lmax = 0, n_runs = 0;
    int n_frozen_tracks = 0;           // frozen tracks with observations = internal tracks [0, n_frozen_tracks)
    long long n_tiles = 0;
    double fill = 0.0;                 // observations per lane slot of the tiles (K / (32 n_tiles)): how well tracks share their camera sets
};

constexpr int PT_MAX_T = 16;           // track slots per tile (bounds the per-warp staging buffers)
inline int pattern_tile_tracks(int L) { return std::min(32 / L, PT_MAX_T); }

struct PatternRun { int trk0, ntrk, L, pts_free; uint64_t mask; long long tile0; };

// passes of the Schur kernel over a unit: a lane holds at most two (camera pair, row chunk) tasks at a time.
// rows_per_task == 0 selects the tensor-core kernel: the (L nc) x (L nc) Gram matrix of a track's Z rows is cut into 8 x 8
// tiles (upper triangle, enumerated column by column: tile (i, j <= ...) has number j (j + 1) / 2 + i), PT_MMA_TILES per pass.
constexpr int PT_MMA_TILES = 21;       // tiles of a pass = accumulator fragments of a lane (2 doubles each): R <= 6 row blocks in one pass
inline int pattern_mma_row_blocks(int L, int nc) { return (L * nc + 7) / 8; }
inline int pattern_schur_passes(int L, int nc, int rows_per_task)
{
    if (rows_per_task == 0) {
        const int R = pattern_mma_row_blocks(L, nc);
        return (R * (R + 1) / 2 + PT_MMA_TILES - 1) / PT_MMA_TILES;
    }
    const int ntask = L * (L + 1) / 2 * ((nc + rows_per_task - 1) / rows_per_task);
    return (ntask + 63) / 64;
}

// Cost of one tile (all tiles of a run have the same track length L) and of starting a unit, per kernel shape, in clocks.  The
// numbers are a least-squares fit of the clocks every warp spent in its unit loop (SBA_PT_CYCLES_FILE, 2368 warps, B200,
// perspective, 6 unknowns per camera, the 1e6-observation bench scene) against the warp's tile counts per track length and its
// number of units (tools/fit_tile_cost.py).  With a uniform tile cost the slowest warp of a kernel took 1.2-1.3x the average, and
// a kernel ends with its slowest warp.  kind: 0 = light (K2 + K4 share the assignment: sum of both), 1 = assembly (K1), 2 = Schur
// (K3, DFMA kernel; the tensor-core variant keeps its formula).
inline int pattern_unit_cost(int kind, int rows_per_task)
{
    if (kind == 2 && rows_per_task == 0) return 0;
    return kind == 0 ? 2100 : (kind == 1 ? 3300 : 3100);
}
inline int pattern_tile_cost(int L, int nc, int rows_per_task, int kind)
{
    static const int light[11] = {0, 0, 5515, 6283, 6405, 7048, 7400, 7734, 7762, 8120, 8422};
    static const int wide[11] = {0, 0, 5872, 5040, 4930, 4756, 4953, 4980, 4811, 4458, 4662};
    static const int narrow[11] = {0, 0, 4312, 4832, 6161, 5406, 6803, 6125, 10793, 9247, 10266};
    const int T = pattern_tile_tracks(L);
    if (kind == 2 && rows_per_task == 0) {          // evaluation once per pass + per track R fragment loads and one MMA per tile
        const int R = pattern_mma_row_blocks(L, nc), nt = R * (R + 1) / 2, np = (nt + PT_MMA_TILES - 1) / PT_MMA_TILES;
        return np * 24 + (T * (np == 1 ? R + nt : 3 * nt)) / 4;
    }
    const int l = L < 2 ? 2 : L;
    if (kind == 0) return l <= 10 ? light[l] : 8422 + 300 * (l - 10);
    if (kind == 1) return l <= 10 ? wide[l] : 4660;
    if (l <= 10) return narrow[l];
    // longer tracks (more than 10 cameras): one evaluation per pass plus the product rounds, scaled to meet the table at L = 10
    const int ntask = L * (L + 1) / 2 * ((nc + rows_per_task - 1) / rows_per_task);
    int cost = 0;
    for (int left = ntask; left > 0; left -= 64) cost += 8 + T * (left > 32 ? 2 : 1) * 3;
    return cost * (10266 / 52);
}

// Cut the tile sequence of the runs into one contiguous range per warp (equal cost); a unit is the part of a warp's
// range inside one run.
inline void assign_pattern_units(const std::vector<PatternRun>& runs, const std::vector<int>& track_ptr, int n_cta, int warps,
                                 int nc, int rows_per_task, int kind, PatternAssignment& A)
{
    A.warps = warps;
    A.units.clear();
    A.n_records = 0;
    const long long nw = (long long)n_cta * warps;
    A.warp_unit0.assign(nw + 1, 0);
    // cumulative cost at the start of every run
    std::vector<long long> cost0(runs.size() + 1, 0);
    const long long ucost = pattern_unit_cost(kind, rows_per_task);       // starting a unit (first tile not prefetched, flush of the camera blocks)
    for (size_t r = 0; r < runs.size(); ++r) {
        const int T = pattern_tile_tracks(runs[r].L);
        cost0[r + 1] = cost0[r] + ucost + (long long)((runs[r].ntrk + T - 1) / T) * pattern_tile_cost(runs[r].L, nc, rows_per_task, kind);
    }
    const long long total = cost0.back();
    size_t r = 0;
    long long tile = 0;          // next unassigned tile of run r
    for (long long g = 0; g < nw; ++g) {
        A.warp_unit0[g] = (int)A.units.size();
        const long long target = total * (g + 1) / nw;       // this warp takes tiles while the cumulative cost stays <= target
        while (r < runs.size()) {
            const PatternRun& R = runs[r];
            const int T = pattern_tile_tracks(R.L), c = pattern_tile_cost(R.L, nc, rows_per_task, kind);
            const long long run_tiles = (R.ntrk + T - 1) / T;
            const long long done = cost0[r] + ucost + tile * c;
            long long take = g == nw - 1 ? run_tiles - tile : std::min(run_tiles - tile, (target - done) / c);
            if (take <= 0) break;
            PUnit u;
            const int s0 = (int)tile * T;
            u.trk0 = R.trk0 + s0; u.ntrk = (int)std::min<long long>(take * T, R.ntrk - s0); u.obs0 = track_ptr[u.trk0];
            u.L = R.L; u.mask_lo = (unsigned)(R.mask & 0xffffffffu); u.mask_hi = (unsigned)(R.mask >> 32);
            u.pts_free = R.pts_free; u.rec = A.n_records;
            A.n_records += pattern_schur_passes(R.L, nc, rows_per_task);
            A.units.push_back(u);
            tile += take;
            if (tile == run_tiles) { ++r; tile = 0; } else break;
        }
    }
    A.warp_unit0[nw] = (int)A.units.size();
}

// cam: (K) camera of every observation, caller's order; track_ptr_old: (N+1) first observation of every track.
// n_pts_fix: the caller's first n_pts_fix tracks are frozen.
inline void build_pattern_layout(const int* cam, const int* track_ptr_old, long long K, int M, int N, int n_pts_fix,
                                 int n_cta, int warps_light, int warps_wide, int warps_narrow, int nc, int rows_per_task,
                                 PatternLayout& out)
{
    out = PatternLayout();
    if (M > 64) { out.why = "more than 64 cameras"; return; }
    if (N < 1 || K < 1) { out.why = "empty"; return; }
    // key = camera bit set; aux orders frozen tracks first and tracks without observations last (a few threads: O(K))
    std::vector<uint64_t> key(N);
    std::vector<unsigned char> aux(N);          // bit 0: free (frozen tracks come first), bit 1: no observations (sorted last)
    const int hw = (int)std::thread::hardware_concurrency();
    const int nthr = std::max(1, std::min({8, hw > 0 ? hw : 1, N / 16384 + 1}));
    std::vector<int> t_lmax(nthr, 0), t_frozen(nthr, 0), t_bad(nthr, 0);
    auto scan = [&](int th) {
        const int i0 = (int)((long long)N * th / nthr), i1 = (int)((long long)N * (th + 1) / nthr);
        int lmax = 0, frozen = 0;                       // thread-local: the per-thread slots share cache lines
        for (int i = i0; i < i1; ++i) {
            const int a0 = track_ptr_old[i], a1 = track_ptr_old[i + 1];
            uint64_t m = 0;
            int prev = -1;
            for (int a = a0; a < a1; ++a) {
                if (cam[a] <= prev) { t_bad[th] = 1; return; }
                prev = cam[a];
                m |= (uint64_t)1 << cam[a];
            }
            key[i] = m;
            aux[i] = (unsigned char)((i < n_pts_fix ? 0 : 1) | (a1 == a0 ? 2 : 0));
            if (i < n_pts_fix && a1 > a0) ++frozen;
            lmax = std::max(lmax, a1 - a0);
        }
        t_lmax[th] = lmax;
        t_frozen[th] = frozen;
    };
    if (nthr == 1) scan(0);
    else {
        std::vector<std::thread> pool;
        for (int th = 0; th < nthr; ++th) pool.emplace_back(scan, th);
        for (auto& t : pool) t.join();
    }
    int Lmax = 0;
    for (int th = 0; th < nthr; ++th) {
        if (t_bad[th]) { out.why = "cameras not strictly ascending inside a track"; return; }
        Lmax = std::max(Lmax, t_lmax[th]);
        out.n_frozen_tracks += t_frozen[th];
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
    // runs and their tiles
    std::vector<PatternRun> runs;
    long long tiles = 0;
    for (int t = 0; t < N;) {
        const int o = order[t];
        if (aux[o] & 2) break;                       // tracks without observations: no work
        int e = t + 1;
        while (e < N && key[order[e]] == key[o] && aux[order[e]] == aux[o]) ++e;
        const int L = track_ptr_old[o + 1] - track_ptr_old[o];
        runs.push_back({t, e - t, L, (aux[o] & 1) ? 1 : 0, key[o], tiles});
        const int T = pattern_tile_tracks(L);
        tiles += (e - t + T - 1) / T;
        t = e;
    }
    out.n_runs = (int)runs.size();
    out.n_tiles = tiles;
    out.fill = tiles > 0 ? (double)K / (32.0 * (double)tiles) : 0.0;
    out.n_cta = n_cta;
    if (runs.empty()) { out.why = "no observations"; return; }
    assign_pattern_units(runs, out.track_ptr, n_cta, warps_light, nc, rows_per_task, 0, out.light);
    assign_pattern_units(runs, out.track_ptr, n_cta, warps_wide, nc, rows_per_task, 1, out.wide);
    assign_pattern_units(runs, out.track_ptr, n_cta, warps_narrow, nc, rows_per_task, 2, out.narrow);
    out.ok = true;
}

}  // namespace sba
