"""bench.py --mode sharded: ONE oversized window (BASELINE.json configs[4]: "one 2M-residual sharded
window"), points sharded across ranks (h-slam_b200/shard.py), every rank runs the 5-launch GN-iteration hot
path on its shard and the stitched fp64 [H_A|b_A|H_L|b_L|H_SC|b_SC] block (87.8 kB at n=7) is summed with ONE
NCCL all-reduce.  Strong scaling: the total work is fixed as N grows."""
from __future__ import annotations

import importlib
import json
import os
import time

import numpy as np


class _DevArray:
    """Minimal __cuda_array_interface__ holder so torch can view the library's device buffer (no copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3, "strides": None}


def make_big_window(synth, points_per_frame, seed=900):
    kf = synth.render_keyframes(640, 480, 7, seed=4321)
    win, _ = synth.make_window(points_per_frame=points_per_frame, seed=seed, keyframes=kf)
    return win


def run_sharded(args, rank, local_rank, world, capi, synth):
    import torch
    import torch.distributed as dist

    shard = importlib.import_module("h-slam_b200.shard")
    ppf = int(os.environ.get("HSLAM_SHARDED_PPF", "47620"))    # 7 x 47620 points x 6 targets = 2.0 M residuals
    win = make_big_window(synth, ppf)
    sub = shard.shard_window(win, rank, world) if world > 1 else win
    ctx = capi.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))
    frames = []
    keep = []
    for im in win.images:
        t = torch.from_numpy(np.ascontiguousarray(im)).cuda()
        keep.append(t)
        frames.append(capi.Frame(ctx, None, 1, device_ptr=t.data_ptr(), w=win.cam.w, h=win.cam.h))
    ef = capi.EnergyFunctional(ctx)
    ef.set_window(sub, frames)
    ef.set_topology(sub)
    ef.set_add_priors(rank == 0)
    ptr, cnt = ef.iteration_device()
    Hb = torch.as_tensor(_DevArray(ptr, cnt), device=torch.device("cuda", local_rank))

    def step():
        ef.iteration_device()
        if world > 1:
            with torch.cuda.stream(stream):
                dist.all_reduce(Hb, op=dist.ReduceOp.SUM)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # correctness of the sharded sum against the un-sharded window (rank 0 runs the whole window once)
    step(); barrier()
    got = Hb.cpu().numpy().copy()
    check = None
    if rank == 0 and world > 1:
        full = capi.EnergyFunctional(ctx)
        full.set_window(win, frames); full.set_topology(win)
        full.iteration_device()
        ref = np.concatenate([np.concatenate([H.T.reshape(-1), b]) for H, b in full.download_Hb()])   # synchronises
        check = float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300))
        full.close()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launches - l0
    if world > 1:
        tt = torch.tensor([ms], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    px = win.R * 8 * args.steps
    # e2e: per-iteration host state up, stitched system + states down
    def step_e2e():
        ef.begin_async()
        ef.set_window(sub, frames)
        ef.update_points(idepth=sub.idepth, deltaF=sub.deltaF)
        ef.linearizeAll(False, outputs="states")
        ef.applyRes()
        ef.accumulate_all_device()
        if world > 1:
            with torch.cuda.stream(stream):
                dist.all_reduce(Hb, op=dist.ReduceOp.SUM)
        r = ef.download_Hb()
        ef.wait()
        return r
    step_e2e(); barrier()
    k2 = max(3, args.steps // 2)
    t0 = time.perf_counter()
    for _ in range(k2):
        step_e2e()
    barrier()
    wall = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([wall], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        wall = float(tt.item())
    N = 8 * win.n + 4
    n2 = win.n * win.n
    line = {
        "metric": "photometric residual+Hessian throughput (one 2M-residual BA window, point-sharded)",
        "value": px / (ms * 1e-3) / 1e6, "unit": "Mres/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"cfg5b: one window 7KF x {ppf} pts/KF (R={win.R}), points sharded over {world} GPU(s), "
                               "one fp64 NCCL all-reduce of the stitched 3x(H,b) per iteration",
                   "l2_policy": "inputs larger than L2 (no flush)", "allreduce_bytes": int(3 * (N * N + N) * 8)},
        "e2e": {"value": win.R * 8 * k2 / wall / 1e6, "unit": "Mres/s",
                "h2d_bytes_per_step": int(n2 * 28 * 4 + n2 * 8 * 4 + 2 * n2 * 64 * 8 + 2 * win.n * 64 + sub.P * 20),
                "d2h_bytes_per_step": int(8 + sub.R * 9 + 3 * (N * N + N) * 8)},
        "gpu_launches": int(launches), "sharded_check_rel_err": check,
    }
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
