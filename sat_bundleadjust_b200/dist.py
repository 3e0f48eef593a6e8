"""
Multi-GPU bundle adjustment: observations shard by track across the ranks of one box
(SURVEY.md section 8e).

Observations are sorted by track (bundle_adjust/ba_params.py:138-149), so rank r owns a contiguous
track range [t_r, t_{r+1}) and the contiguous observation range it induces.  Point blocks (V, g_p,
the point elimination and the back-substitution) are rank-local; the cameras are replicated.  The
exchange step is one SUM all-reduce of the rank's partial camera system per trust-region iteration
([U | g_c] after the assembly, [S | rhs] after the Schur complement) plus a handful of scalars; every
rank then factors the same reduced camera system redundantly, so all ranks take identical steps.

Plumbing only: torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests) moves the
bytes; the arithmetic stays in libsba_b200.so.
"""
import os

import numpy as np


def shard_ranges(pts_ind, n_pts, world_size):
    """
    Cut the track range [0, n_pts) into `world_size` contiguous pieces holding ~equal numbers of
    observations.  Returns a list of (t0, t1).  Deterministic; every rank computes the same cuts.
    """
    pts_ind = np.asarray(pts_ind)
    K = pts_ind.size
    track_ptr = np.searchsorted(pts_ind, np.arange(n_pts + 1), side="left")      # first observation of each track
    cuts = [0]
    for r in range(1, world_size):
        target = (K * r) // world_size
        t = int(np.searchsorted(track_ptr, target, side="left"))
        t = min(max(t, cuts[-1]), n_pts)
        cuts.append(t)
    cuts.append(n_pts)
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


def local_vars(x, n_cam_vars, track_range):
    """[cameras | all points] -> [cameras | this rank's points]."""
    t0, t1 = track_range
    return np.concatenate([x[:n_cam_vars], x[n_cam_vars + 3 * t0: n_cam_vars + 3 * t1]])


def merge_vars(x_locals, n_cam_vars):
    """Per-rank solutions -> the global variable vector (cameras are identical on all ranks)."""
    return np.concatenate([x_locals[0][:n_cam_vars]] + [xl[n_cam_vars:] for xl in x_locals])


class _CudaView:
    """Zero-copy view of `count` doubles at a raw device pointer, importable by torch.as_tensor."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def run_ba_optimization_distributed(p, ls_params=None, group=None):
    """
    Multi-GPU counterpart of ba_core.run_ba_optimization, to be called by every rank of an initialised
    torch.distributed NCCL process group (one process per GPU).  Every rank passes the same `p`.
    Returns (vars_init, vars_ba, err_init, err_ba, nfev, info) with the global vectors on every rank.
    """
    from .solver import n_common_params
    if n_common_params(p):
        raise NotImplementedError("COMMON_K is not supported by the distributed driver")
    import torch
    import torch.distributed as dist

    from . import ba_core
    from .solver import DeviceProblem, initial_vars

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    cfg = ba_core.init_optimization_config(ls_params)
    ranges = shard_ranges(p.pts_ind, p.n_pts, world)
    ncv = p.n_cam * p.n_params
    x0 = initial_vars(p)
    stream = torch.cuda.current_stream().cuda_stream

    def hook(ptr, count):
        t = torch.as_tensor(_CudaView(ptr, count), device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)

    def gather_obj(obj):
        out = [None] * world
        dist.all_gather_object(out, obj, group=group)
        return out

    with DeviceProblem(p, stream=stream, rank=rank, world_size=world, track_range=ranges[rank]) as prob:
        prob.set_allreduce(hook)                       # NCCL fall-back for exchanges that do not fit the peer buffer
        if os.environ.get("SBA_COMM", "peer") == "peer":
            prob.connect_peers(gather_obj)
        xl0 = local_vars(x0, ncv, ranges[rank])
        xl, e0, e1, info = prob.solve_with_errors(xl0, loss=cfg["loss"], f_scale=cfg["f_scale"], ftol=cfg["ftol"],
                                                  xtol=cfg["xtol"], max_nfev=cfg["max_iter"])
        dist.barrier(group=group)                      # peers keep reading each other's buffers until everybody is done
    # one all-gather of [x_local | err_init | err] per rank (variable length -> padded to the longest shard)
    a = [np.searchsorted(p.pts_ind, [t0, t1]) for t0, t1 in ranges]
    n_loc = [ncv + 3 * (t1 - t0) for t0, t1 in ranges]
    k_loc = [int(a1 - a0) for a0, a1 in a]
    longest = max(n + 2 * k for n, k in zip(n_loc, k_loc))
    buf = torch.zeros(longest, dtype=torch.float64, device="cuda")
    buf[: n_loc[rank] + 2 * k_loc[rank]] = torch.from_numpy(np.concatenate([xl, e0, e1])).cuda()
    out = torch.empty((world, longest), dtype=torch.float64, device="cuda")
    dist.all_gather_into_tensor(out, buf, group=group)
    out = out.cpu().numpy()
    x = merge_vars([out[r, : n_loc[r]] for r in range(world)], ncv)
    err0 = np.concatenate([out[r, n_loc[r]: n_loc[r] + k_loc[r]] for r in range(world)])
    err1 = np.concatenate([out[r, n_loc[r] + k_loc[r]: n_loc[r] + 2 * k_loc[r]] for r in range(world)])
    return x0, x, err0, err1, info["nfev"], info
