"""Partitioning of the hot path across GPUs.

Two axes, both without any data-path collective except one sum of the tiny stitched system:
  * independent windows / sequences: `assign_windows` deals whole windows to ranks;
  * one oversized window: `shard_window` gives each rank a contiguous range of the points sorted by
    (host frame, index) together with ALL residuals of those points (the Schur fold needs every residual
    of a point on one device, reference Src/AccumulatedSCHessian.cpp:38-52).  Frames and the n^2 tables
    are replicated.  Every rank stitches its own fp64 [H_A|b_A|H_L|b_L|H_SC|b_SC]; the stitch is linear
    in the accumulators, so ONE all-reduce(sum) of 3*((8n+4)^2+(8n+4)) doubles yields the window's
    system.  Priors are added by rank 0 only (mirrors "only do this on one thread",
    Src/AccumulatedTopHessian.cpp:268-279).
"""
from __future__ import annotations

import copy
import numpy as np


def assign_windows(n_windows: int, rank: int, world: int):
    """Contiguous equal-count deal of window ids (fixed boundaries => identical results for any world)."""
    per = (n_windows + world - 1) // world
    return list(range(rank * per, min(n_windows, (rank + 1) * per)))


def point_ranges(host_idx, world: int):
    """Order points by (host, index) and cut into `world` contiguous equal-count ranges.
    Returns (order, bounds) with bounds[r]:bounds[r+1] the slice of `order` owned by rank r."""
    host_idx = np.asarray(host_idx)
    order = np.lexsort((np.arange(len(host_idx)), host_idx))
    P = len(order)
    bounds = [(P * r) // world for r in range(world + 1)]
    return order, bounds


def shard_window(win, rank: int, world: int):
    """The sub-window of `win` owned by `rank`: its points, their residuals, re-indexed; tables shared."""
    order, bounds = point_ranges(win.host_idx, world)
    mine = np.sort(order[bounds[rank]:bounds[rank + 1]])
    sub = copy.copy(win)
    for k in ("u", "v", "idepth", "idepth_zero", "color", "weights", "priorF", "deltaF", "host_idx"):
        setattr(sub, k, np.ascontiguousarray(getattr(win, k)[mine]))
    # residuals of my points, in per-point order
    rb = win.res_begin
    cnt = (rb[mine + 1] - rb[mine]).astype(np.int64)
    new_begin = np.zeros(len(mine) + 1, np.int64)
    np.cumsum(cnt, out=new_begin[1:])
    old_lists = [win.res_index[rb[p]:rb[p + 1]] for p in mine]
    old_res = np.concatenate(old_lists) if len(old_lists) else np.zeros(0, np.int64)
    remap = {int(r): i for i, r in enumerate(old_res)}
    sub.res_begin = new_begin.astype(np.int32)
    sub.res_index = np.arange(len(old_res), dtype=np.int32)
    new_point = np.repeat(np.arange(len(mine), dtype=np.int32), cnt)
    sub.point_idx = new_point
    for k in ("target_idx", "state_state", "is_linearized", "is_active"):
        setattr(sub, k, np.ascontiguousarray(getattr(win, k)[old_res]))
    sub._orig_points = mine
    sub._orig_residuals = old_res
    sub._remap = remap
    return sub
