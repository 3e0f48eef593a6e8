"""
Fits the cost model of the pattern engine's static assignment (csrc/sba_pattern.h: pattern_tile_cost, pattern_unit_cost):
    SBA_PT_CYCLES=1 SBA_PT_CYCLES_FILE=gpurun_out/cyc.bin python tools/profile_iter.py 1m 6      (on the GPU box)
    python tools/fit_tile_cost.py gpurun_out/cyc.bin                                             (anywhere)
Least squares of the clocks every warp spent in its unit loop against its number of units and its tile counts per track length.
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402

p = bench.build_problem("1m", 1)
lib = ctypes.CDLL(os.path.join(ROOT, "tests", "host_harness", "libmodel_harness.so"))
ip = ctypes.POINTER(ctypes.c_int)
cam = np.ascontiguousarray(p.cam_ind, dtype=np.int32)
N = p.n_pts
tp = np.ascontiguousarray(np.searchsorted(p.pts_ind, np.arange(N + 1), side="left"), dtype=np.int32)
h = np.fromfile(sys.argv[1], dtype=np.int64)
names = {0: "K2 light", 1: "K1 wide", 2: "K3 narrow", 3: "K4 light"}
warps = {0: 16, 1: 16, 2: 12, 3: 16}
which = {0: 0, 1: 1, 2: 2, 3: 0}
for k in range(4):
    nw = warps[k]
    units = np.zeros((40000, 8), np.int32)
    wu0 = np.zeros(148 * nw + 1, np.int32)
    n = lib.hh_pattern_assignment(cam.ctypes.data_as(ip), tp.ctypes.data_as(ip), ctypes.c_longlong(cam.size), int(p.n_cam), N, 0, 148, 16, 16, 12,
                                  int(p.n_params), 3, which[k], units.ctypes.data_as(ip), 40000, wu0.ctypes.data_as(ip))
    units = units[:n]
    cyc = h[k * 148 * 32: k * 148 * 32 + 148 * nw].astype(float)
    A = np.zeros((148 * nw, 11))
    for g in range(148 * nw):
        for u in range(wu0[g], wu0[g + 1]):
            ntrk, L = units[u, 1], units[u, 3]
            T = min(32 // L, 16)
            A[g, L] += (ntrk + T - 1) // T
        A[g, 0] = wu0[g + 1] - wu0[g]
        A[g, 1] = 1.0
    sol, *_ = np.linalg.lstsq(A, cyc, rcond=None)
    pred = A @ sol
    print(names[k], "tiles/warp %.1f, clocks avg %.0f max %.0f (max/avg %.2f)" % (A[:, 2:].sum(axis=1).mean(), cyc.mean(), cyc.max(), cyc.max() / cyc.mean()),
          "| per unit %.0f, constant %.0f, per tile for L = 2..10:" % (sol[0], sol[1]), np.round(sol[2:]).astype(int),
          "| residual %.1f%% of the mean, R2 %.2f" % (100 * (cyc - pred).std() / cyc.mean(), 1 - ((cyc - pred) ** 2).sum() / ((cyc - cyc.mean()) ** 2).sum()))
