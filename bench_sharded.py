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


def sharded_record(args, rank, local_rank, world, capi, synth, ctx=None):
    """Measure the point-sharded 2M-residual window on `world` GPUs; returns the record (every rank must call)."""
    import torch
    import torch.distributed as dist

    shard = importlib.import_module("h-slam_b200.shard")
    ppf = int(os.environ.get("HSLAM_SHARDED_PPF", "47620"))    # 7 x 47620 points x 6 targets = 2.0 M residuals
    win = make_big_window(synth, ppf)
    sub = shard.shard_window(win, rank, world) if world > 1 else win
    own_ctx = ctx is None
    if own_ctx:
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
    ef.set_jacobian_mode(1)
    ef.set_add_priors(rank == 0)              # exactly one rank adds the priors (hslam_ba_allreduce's contract)
    comm = None
    if world > 1:
        # ncclComm_t through the library's own NCCL handle (the one torch bundles); the id travels over torch.distributed
        capi.use_torch_nccl()
        ids = [capi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = capi.NcclComm(ctx, ids[0], world, rank)

    def step():
        if comm is not None:
            # one CUDA graph (linearise + applyRes + top + Schur + fold + stitch of the shard), then hslam_ba_allreduce's
            # ncclAllReduce(sum, fp64) of 3 x (H,b) + energy + counts on the same stream
            ef.iteration_allreduce(comm)
        else:
            ef.iteration_device()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(k):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ms], device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms / k

    # correctness of the sharded sum against the un-sharded window (rank 0 runs the whole window as well)
    step(); barrier()
    got = np.concatenate([np.concatenate([H.T.reshape(-1), b]) for H, b in ef.download_Hb()])
    check, ms_full = None, None
    if world > 1:
        if rank == 0:
            full = capi.EnergyFunctional(ctx)
            full.set_window(win, frames); full.set_topology(win); full.set_jacobian_mode(1)
            full.iteration_device()
            ref = np.concatenate([np.concatenate([H.T.reshape(-1), b]) for H, b in full.download_Hb()])   # synchronises
            check = float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300))
            for _ in range(3):
                full.iteration_device()
            ctx.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(10):
                full.iteration_device()
            e1.record(stream)
            torch.cuda.synchronize()
            ms_full = e0.elapsed_time(e1) / 10
            full.close()
        dist.barrier()
    for _ in range(max(args.warmup, 3)):
        step()
    l0 = ctx.launches
    k = max(args.steps, 10)
    ms = timed(step, k)
    launches = ctx.launches - l0
    N = 8 * win.n + 4
    rec = {"workload": f"cfg5b: one window 7KF x {ppf} pts/KF (R={win.R}), points sharded over {world} GPU(s), one fp64 "
                       "hslam_ba_allreduce (ncclAllReduce) of the stitched 3x(H,b) + energy + counts per iteration",
           "residuals": int(win.R), "ms_per_iteration": ms, "value": win.R * 8 / (ms * 1e-3) / 1e6, "unit": "Mres/s",
           "n_gpus": world, "scaling": "strong", "allreduce_bytes": int((3 * (N * N + N) + 3) * 8),
           "gpu_launches_per_iteration": launches / k, "sharded_check_rel_err": check,
           "ms_per_iteration_1gpu_same_run": ms_full if world > 1 else ms,
           "speedup_vs_1gpu": (ms_full / ms) if ms_full else (1.0 if world == 1 else None)}
    if comm is not None:
        comm.close()
    ef.close()
    for f in frames:
        f.close()
    if own_ctx:
        ctx.close()
    return rec


def run_sharded(args, rank, local_rank, world, capi, synth):
    """bench.py --mode sharded: the sharded window alone, as its own JSON line."""
    import torch.distributed as dist
    rec = sharded_record(args, rank, local_rank, world, capi, synth)
    line = {
        "metric": "photometric residual+Hessian throughput (one 2M-residual BA window, point-sharded)",
        "value": rec["value"], "unit": "Mres/s", "n_gpus": world, "steps": max(args.steps, 10), "warmup": max(args.warmup, 3),
        "ms_per_step": rec["ms_per_iteration"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": rec["workload"], "l2_policy": "inputs larger than L2 (no flush)"},
        "sharded": rec, "gpu_launches": int(rec["gpu_launches_per_iteration"] * max(args.steps, 10)),
    }
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
