"""
CPU tests of the multi-GPU host logic with world_size 2 over gloo: the track sharding, and that the
quantities exchanged per iteration (camera blocks U, g_c and the Schur complement S, rhs) are additive
over track shards while the point blocks stay rank-local -- which is what makes one SUM all-reduce of
the rank's partial camera system sufficient.
"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util
from oracle import ba_oracle
from sat_bundleadjust_b200 import dist as sdist
from sat_bundleadjust_b200 import synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _partial_systems(p, J, f, track_range, lam=1e-3):
    """numpy statement of what one rank contributes: U_r, g_c,r and S_r, rhs_r from its own tracks."""
    c, M = p.n_params, p.n_cam
    nc = M * c
    t0, t1 = track_range
    a0, a1 = np.searchsorted(p.pts_ind, [t0, t1])
    rows = slice(2 * a0, 2 * a1)
    Jr, fr = J[rows], f[rows]
    Jc = Jr[:, :nc]
    U, gc = Jc.T @ Jc, Jc.T @ fr
    S, rhs = U.copy(), -gc.copy()
    for i in range(t0, t1):
        cols = slice(nc + 3 * i, nc + 3 * i + 3)
        Jp = Jr[:, cols]
        V = Jp.T @ Jp + lam * np.eye(3)
        W = Jc.T @ Jp
        Vi = np.linalg.inv(V)
        S -= W @ Vi @ W.T
        rhs += W @ Vi @ (Jp.T @ fr)
    return U, gc, S, rhs


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc = synth.make_scene(n_cam=4, n_tracks=60, p_vis=0.7, cam_model="perspective", seed=21)
        p = synth.scene_to_params(sc, ["R", "T"])
        x = p.params_opt.copy()
        J = ba_oracle.dense_jacobian_fd(x.copy(), p, rel_step=1e-7)
        f = ba_oracle.residuals(x.copy(), p)
        ranges = sdist.shard_ranges(p.pts_ind, p.n_pts, world)
        U, gc, S, rhs = _partial_systems(p, J, f, ranges[rank])
        buf = torch.from_numpy(np.concatenate([U.ravel(), gc, S.ravel(), rhs]))
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        Ug, gcg, Sg, rhsg = _partial_systems(p, J, f, (0, p.n_pts))
        ref = np.concatenate([Ug.ravel(), gcg, Sg.ravel(), rhsg])
        ok = np.allclose(buf.numpy(), ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())
        # local / merged variable vectors
        ncv = p.n_cam * p.n_params
        xl = sdist.local_vars(x, ncv, ranges[rank])
        gathered = [None] * world
        dist.all_gather_object(gathered, xl)
        ok = ok and np.array_equal(sdist.merge_vars(gathered, ncv), x)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_shard_ranges_cover_and_balance():
    sc = synth.make_scene(n_cam=8, n_tracks=5000, p_vis=0.4, seed=4)
    for world in (1, 2, 3, 8):
        rg = sdist.shard_ranges(sc.pts_ind, sc.n_pts, world)
        assert rg[0][0] == 0 and rg[-1][1] == sc.n_pts
        assert all(rg[r][1] == rg[r + 1][0] for r in range(world - 1))
        counts = [int(np.sum((sc.pts_ind >= a) & (sc.pts_ind < b))) for a, b in rg]
        assert sum(counts) == sc.n_obs
        assert max(counts) - min(counts) <= 2 * 8      # within two maximal tracks of each other
    # degenerate: more ranks than tracks
    rg = sdist.shard_ranges(np.array([0, 0, 1, 1]), 2, 4)
    assert rg[0][0] == 0 and rg[-1][1] == 2 and all(a <= b for a, b in rg)


def test_partial_camera_systems_sum_over_ranks_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]
