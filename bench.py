#!/usr/bin/env python
"""bench.py — hot-path benchmark (driver contract; see DESIGN.md "Measurement").

Workload (BASELINE.json configs[1] as SURVEY.md §8d specifies it): windowed photometric BA linearisation, 7 keyframes
x 2000 points per keyframe x 8-px pattern (P=14000 points, R=84000 residuals, 672k pattern pixels per window), 30 % of
the residuals marked isLinearized (mode L of the top accumulation), the rest re-linearised every iteration.
A *step* = ONE Gauss-Newton iteration's hot path (linearizeAll + setNewFrameEnergyTH + applyRes + accumulateAF/LF/SCF
with the per-point Schur fold + fp64 stitch) over a batch of `--windows` independent windows per GPU (8 rendered
scenes), each with its own frames/points/Jacobians in HBM (the batch working set is >> the 126 MB L2, so no flush is
needed between iterations).  Metric: pattern-pixel residuals linearised / reduced into H,b per second (Mres/s), whole
job.  The all-active variant (0 % linearised: every residual goes through the lineariser) is reported next to it.

    python bench.py --gpus N --steps K --warmup W           # ours (CUDA through the C ABI)
    python bench.py --impl reference ...                    # the reference's CPU algorithm (oracle port, all host cores)

Multi-GPU: one process per GPU (torchrun), windows are independent => weak scaling, no data-path
collective; `--mode sharded` instead shards ONE oversized window's points across ranks and sums the
stitched fp64 H,b with one NCCL all-reduce.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FRAMES = 7
PTS_PER_FRAME = 2000
METRIC = "photometric residual+Hessian throughput (BA linearise+accumulate+Schur, 7KFx2000pts)"
UNIT = "Mres/s"

# algorithmic bytes per residual of each kernel class (DESIGN.md "Kernels and rooflines")
ALGO_BYTES_PER_RES = {
    # 8 px x 4 texels x 12 B, point record/6, indices/state; written: the 34 geometric + summed floats of RawResidualJacobian
    # (GN-loop mode: the 40 per-pixel floats and projectedTo, 160 + 64 B, are not materialised; 769 B with them)
    "ba_linearize": 384 + 13.3 + 12 + 136,
    "ba_top": 156 + 32,                           # mode A: 39 Jacobian floats read, one 32-byte point-term record written
    "ba_schur": 32 + 2,                           # JpJdF read per residual
}


def _clock_sampler(stop, out, gpu_index):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(gpu_index)],
                               capture_output=True, text=True, timeout=5)
            parts = [p.strip() for p in r.stdout.strip().split(",")]
            if len(parts) >= 6:
                out.append(parts)
        except Exception:
            pass
        stop.wait(0.2)


def _clock_summary(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    sm = sorted(int(s[0]) for s in samples if s[0].isdigit())
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in samples)]
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(samples[0][1]) if samples[0][1].isdigit() else None,
            "reasons": reasons}


FRAC_LINEARIZED = 0.3      # "30 % of residuals marked isLinearized (exercise mode 1)"
N_SCENES = 8


def build_windows(synth, n_windows, seed0, n_scenes=N_SCENES, frac_linearized=FRAC_LINEARIZED):
    """`n_windows` cfg2 windows over `n_scenes` rendered 7-keyframe scenes; returns (windows, per-window isLinearized flags)."""
    kfs = [synth.render_keyframes(640, 480, N_FRAMES, seed=4321 + 17 * s) for s in range(min(n_scenes, n_windows))]
    wins, lins = [], []
    for i in range(n_windows):
        w, is_lin = synth.make_window(points_per_frame=PTS_PER_FRAME, seed=seed0 + i, keyframes=kfs[i % len(kfs)],
                                      frac_linearized=frac_linearized)
        wins.append(w); lins.append(is_lin)
    return wins, lins


def reference_arm(args, rank):
    """The reference's own CPU algorithm for the path (oracle port, -O3 -march=native, IndexThreadReduce
    semantics on all host cores) on the SAME workload as our arm: every step runs one GN iteration's hot path over
    the same `--windows` cfg2 windows (same seeds, 30 % linearised), one after the other."""
    if rank != 0:
        return
    synth = importlib.import_module("h-slam_b200.synth")
    from oracle import oraclelib as O
    cores = os.cpu_count() or 1
    wins, lins = build_windows(synth, args.windows, seed0=100)
    pyr = {}
    bas = []
    for w, is_lin in zip(wins, lins):
        dIs = []
        for im in w.images:
            if id(im) not in pyr:
                pyr[id(im)] = O.pyramid(im, 1)[0][0]
            dIs.append(pyr[id(im)])
        ba = O.BAOracle(w, dIs, prec="f", threads=cores, fast=True)
        ba.linearize(False); ba.apply_res()
        ba.fix_linearization(np.nonzero(is_lin)[0].astype(np.int32))
        bas.append((ba, w.n - 1))

    def step():
        for ba, newest in bas:
            ba.linearize(False); ba.new_frame_energy_th(newest); ba.apply_res()
            ba.accumulate(0); ba.accumulate(1); ba.accumulate(2)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    R = sum(w.R for w in wins)
    val = R * 8 * args.steps / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.windows),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x {args.windows} cfg2 windows (84k residuals each, 30 % linearised), "
                                   f"oracle port -O3 -march=native on {cores} threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(windows):
    return {"workload": f"cfg2 (SURVEY 8d): BA window 7KF x 2000pts/KF x 8px (P=14000, R=84000), 30% of residuals isLinearized, "
                        f"1 GN iteration hot path (linearizeAll incl. setNewFrameEnergyTH + applyRes + accumulateAF/LF/SCF + stitch), {windows} independent windows per GPU per step over {N_SCENES} rendered scenes",
            "windows_per_gpu": windows, "frac_linearized": FRAC_LINEARIZED, "scenes": N_SCENES}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--windows", type=int, default=64,
                    help="independent windows per GPU per step (BASELINE.json configs[4]: 'batched 64 independent synthetic windows')")
    ap.add_argument("--mode", default="batched", choices=["batched", "sharded"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the tracking / single-window latency figures (profiling runs)")
    ap.add_argument("--no-sharded", action="store_true", help="skip the sharded 2M-residual window sub-record")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    capi = importlib.import_module("h-slam_b200.capi")
    synth = importlib.import_module("h-slam_b200.synth")
    if args.mode == "sharded":
        from bench_sharded import run_sharded
        run_sharded(args, rank, local_rank, world, capi, synth)
        return

    ctx = capi.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))
    wins, lins = build_windows(synth, args.windows, seed0=100 + 1000 * rank)
    efs, frames_all, host_imgs = [], [], []
    for w in wins:
        # `value` leg: inputs resident in HBM before the timed region (images uploaded via torch, pyramid built on device);
        # every window owns device copies of its seven keyframes (independent windows share nothing)
        fr = []
        for im in w.images:
            t = torch.from_numpy(np.ascontiguousarray(im)).cuda(non_blocking=False)
            host_imgs.append(t)
            fr.append(capi.Frame(ctx, None, 1, device_ptr=t.data_ptr(), w=w.cam.w, h=w.cam.h))
        ef = capi.EnergyFunctional(ctx)
        ef.set_window(w, fr)
        ef.set_topology(w)
        ef.set_jacobian_mode(1)       # GN-loop mode (hslam_ba_set_jacobian_mode): per-pixel Jacobian rows only for fixLinearizationF
        efs.append(ef); frames_all.append(fr)
    ctx.synchronize()
    px_per_step = sum(w.R for w in wins) * 8
    res_per_step = sum(w.R for w in wins)

    batch = capi.EnergyFunctionalBatch(ctx, efs)
    batch.set_newest_frame(N_FRAMES - 1)   # linearizeAll includes System::setNewFrameEnergyTH (device-side select, side branch of the graph)

    def step_device():
        batch.iteration()          # one CUDA graph of the launch chain for all windows of this GPU

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k, w):
        for _ in range(w):
            fn()
        barrier()
        l0 = ctx.launches
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
        return ms, ctx.launches - l0

    # ---- secondary leg first (nothing linearised yet): every residual goes through the lineariser
    ms_all, _ = timed(step_device, args.steps, args.warmup)
    value_all_active = world * px_per_step * args.steps / (ms_all * 1e-3) / 1e6
    # ---- cfg2 proper: 30 % of the residuals are fixed (fixLinearizationF) and accumulate through mode L from now on
    n_lin = 0
    for ef, is_lin in zip(efs, lins):
        ids = np.nonzero(is_lin)[0].astype(np.int32)
        ef.fixLinearization(ids)
        n_lin += len(ids)
    ctx.synchronize()
    res_lin_per_step = res_per_step - n_lin        # residuals the lineariser actually processes per step
    samples, stop = [], threading.Event()
    th = threading.Thread(target=_clock_sampler, args=(stop, samples, local_rank), daemon=True)
    th.start()
    t_max, launches = timed(step_device, args.steps, args.warmup)
    stop.set(); th.join(timeout=2)
    value = world * px_per_step * args.steps / (t_max * 1e-3) / 1e6

    # ---- per-kernel device timing (CUDA events on the launching stream) for the roofline
    ctx.profile_enable(True)
    for _ in range(args.steps):
        step_device()
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    dom = max(prof.items(), key=lambda kv: kv[1][0])[0]
    dom_ms, dom_cnt = prof[dom]
    launches_per_step = max(1, round(dom_cnt / args.steps))      # ba_top runs twice per step (modes A and L)
    units = {"ba_linearize": res_lin_per_step, "ba_top": res_per_step, "ba_schur": res_per_step}.get(dom, res_per_step)
    algo_bytes = ALGO_BYTES_PER_RES.get(dom, 0.0) * units / (launches_per_step if dom == "ba_top" else 1)
    avg_s = dom_ms * 1e-3 / max(dom_cnt, 1)
    achieved = algo_bytes / avg_s / 1e9 if avg_s > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_kind, "avg_launch_us": avg_s * 1e6,
                "algorithmic_bytes_per_launch": algo_bytes, "units_per_launch": int(units),
                "bytes_per_unit": ALGO_BYTES_PER_RES.get(dom),
                "kernel_share_of_step": {k: v[0] / sum(x[0] for x in prof.values()) for k, v in prof.items()},
                "kernel_avg_us": {k: 1e3 * v[0] / max(v[1], 1) for k, v in prof.items()}}

    # ---- e2e: the same GN iteration through the C ABI with HOST buffers, copies inside the timed region.
    # Default result set = what System::optimize's loop body reads on the host per iteration: energy, resInA/L, the three
    # stitched systems, per-residual newState, the new frameEnergyTH of the newest frame (System::setNewFrameEnergyTH runs
    # on the device: the per-residual energies stay in HBM), per-point step after the back-substitution.
    N = 8 * N_FRAMES + 4
    xs = [np.ascontiguousarray(np.random.default_rng(i).normal(0, 1e-4, N)) for i in range(len(wins))]
    n2 = N_FRAMES * N_FRAMES
    h2d = sum(n2 * 28 * 4 + n2 * 8 * 4 + 2 * n2 * 64 * 8 + 2 * N_FRAMES * 64 + w.P * 4 + N * 8 + 2 * 144 for w in wins)   # tables, idepth, x, scalars

    def e2e_leg(with_energies):
        io = batch.make_io(wins, frames_all, pinned="bound")  # outputs live in the batch's page-locked result block (zero-copy views)
        for o in io["arr"]:
            o.deltaF = None; o.derive_deltaF = 1      # deltaF = idepth - idepth_zero is formed on the device (setDeltaF's point part)
        if not with_energies:
            for o in io["arr"]:
                o.newEnergy = None; o.newEnergyWithOutlier = None

        def step_e2e():
            # phase 1: ONE C-ABI call for the batch — host tables + per-point state up, linearizeAll + setNewFrameEnergyTH
            # + applyRes + accumulateAF/LF/SCF, results down, returns when they are valid
            batch.step_host(io)
            # the caller keeps the threshold in its Frame and hands it back in the next hslam_window (as the reference does)
            batch.feed_back_thresholds(io)
            # (the host solves the 60x60 systems here in the real pipeline) phase 2: back-substitution, x up, steps down
            batch.resubstitute_host(io, xs)

        for _ in range(2):
            step_e2e()
        barrier()
        k = max(3, args.steps // 2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        for _ in range(k):
            step_e2e()
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), 1e3 * wall)
        if world > 1:
            tt = torch.tensor([ms], device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        d2h = sum(8 + 8 + 4 + w.R * (9 if with_energies else 1) + 3 * (N * N + N) * 8 + w.P * 4 for w in wins)
        return world * px_per_step * k / (ms * 1e-3) / 1e6, ms / k, k, int(d2h)

    e2e_val, e2e_ms, k_e2e, d2h = e2e_leg(False)
    full_val, full_ms, _, d2h_full = e2e_leg(True)

    cfg = workload_config(args.windows)
    cfg.update({"l2_policy": f"inputs larger than L2 (no flush): batch working set ~{args.windows * 62} MB vs 126 MB L2",
                "jacobians": "GN-loop mode: the 34 geometric + summed Jacobian floats per residual are materialised (SoA in HBM), "
                             "the 40 per-pixel floats + projectedTo only when a residual is fixed (fixLinearizationF)",
                "e2e_api": "hslam_ba_batch_step_host + hslam_ba_batch_resubstitute_host (host buffers, 2 waits per step)"})
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms, "steps": k_e2e,
                "outputs": "energy, resInA/L, 3x(H,b), per residual newState, new frameEnergyTH of the newest frame "
                           "(setNewFrameEnergyTH on the device), per point step",
                "with_per_residual_energies": {"value": full_val, "unit": UNIT, "ms_per_step": full_ms,
                                               "d2h_bytes_per_step": d2h_full}},
        "gpu_launches": int(launches),
        "clocks": _clock_summary(samples),
        "roofline": roofline,
        "all_active": {"value": value_all_active, "unit": UNIT, "ms_per_step": ms_all / args.steps,
                       "workload": "same windows before fixLinearizationF: 0 % linearised, every residual is re-linearised"},
    }

    # ---- the other multi-GPU axis (BASELINE.json configs[4]): ONE 2M-residual window, points sharded over the ranks, one
    # hslam_ba_allreduce per iteration (strong scaling; at N=1 the whole window runs on the one GPU)
    if not args.no_sharded:
        batch.close()
        for ef in efs:
            ef.close()        # the batched windows' HBM is not needed any more
        for fr in frames_all:
            for f in fr:
                f.close()
        host_imgs.clear()
        from bench_sharded import sharded_record
        try:
            line["sharded"] = sharded_record(args, rank, local_rank, world, capi, synth, ctx=ctx)
        except Exception as e:   # must not take the bench line down
            line["sharded"] = {"error": repr(e)}
            if world > 1:
                raise
    if rank == 0:
        if not args.no_extras:
            line["extra"] = extras(ctx, capi, synth, torch, stream, None, wins[0])
        if not args.no_cpu_baseline and world == 1:      # the CPU baseline is a rank-0, N=1 figure (contract)
            line["cpu_baseline"] = cpu_baseline(synth, wins[0], lins[0])
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def extras(ctx, capi, synth, torch, stream, ef0, win0):
    """Latency figures of the other two headline quantities: BA ms/window and tracking ms/frame."""
    out = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # BA: one window, one GN iteration, device time (with the chunking of a window that runs alone, not the batch's)
    fr0 = [capi.Frame(ctx, im, 1) for im in win0.images]
    ef0 = capi.EnergyFunctional(ctx)
    ef0.set_window(win0, fr0); ef0.set_topology(win0); ef0.set_jacobian_mode(1)
    for _ in range(3):
        ef0.iteration_device()
    ctx.synchronize()
    reps = 20
    e0.record(stream)
    for _ in range(reps):
        ef0.iteration_device()
    e1.record(stream)
    torch.cuda.synchronize()
    out["ba_ms_per_window_device"] = e0.elapsed_time(e1) / reps
    out["ba_window_note"] = "one cfg2 window alone (84k residuals, all re-linearised), the launch chain as one CUDA graph"
    # cfg4 (BASELINE.json configs[3]): 1280x1024, 10-keyframe window, 8000 points per keyframe (R = 720k), one GN iteration
    try:
        kf4 = synth.render_keyframes(1280, 1024, 10, seed=4444)
        win4, _ = synth.make_window(points_per_frame=8000, seed=4445, keyframes=kf4)
        fr4 = [capi.Frame(ctx, im, 1) for im in win4.images]
        ef4 = capi.EnergyFunctional(ctx)
        ef4.set_window(win4, fr4); ef4.set_topology(win4); ef4.set_jacobian_mode(1)
        for _ in range(3):
            ef4.iteration_device()
        ctx.synchronize()
        e0.record(stream)
        for _ in range(reps):
            ef4.iteration_device()
        e1.record(stream)
        torch.cuda.synchronize()
        out["cfg4_ba_ms_per_window_device"] = e0.elapsed_time(e1) / reps
        out["cfg4_residuals"] = int(win4.R)
        out["cfg4_mres_per_s"] = win4.R * 8 / (out["cfg4_ba_ms_per_window_device"] * 1e-3) / 1e6
        ef4.close()
        for f_ in fr4:
            f_.close()
    except Exception as e:
        out["cfg4_error"] = repr(e)
    ef0.close()
    for f_ in fr0:
        f_.close()
    # tracking: config 1 (640x480, 5 levels, ~28k level-0 points)
    tp = synth.make_tracking_pair()
    trk = capi.CoarseTracker(ctx, tp.cam.w, tp.cam.h)
    trk.makeK(tp.cam, tp.levels)
    trk.setCoarseTrackingRef(tp.pc)
    frame = capi.Frame(ctx, tp.img_new, tp.levels)
    T0 = np.concatenate([np.eye(3).reshape(9), np.zeros(3)])
    out["tracking_mode_cluster_ctas"] = trk.get_mode()
    for mode, key in ((trk.get_mode(), "tracking_ms_per_frame_device"), (0, "tracking_ms_per_frame_device_cooperative")):
        trk.set_mode(mode)
        for _ in range(3):
            trk.trackNewestCoarse(frame, T0, (0.0, 0.0), tp.levels - 1)
        e0.record(stream)
        for _ in range(reps):
            trk.track_enqueue(frame, T0, (0.0, 0.0), tp.levels - 1)
        e1.record(stream)
        torch.cuda.synchronize()
        trk.track_finish()
        out[key] = e0.elapsed_time(e1) / reps
        out[key + "_phase_cycles"] = trk.last_profile()
    trk.set_mode(out["tracking_mode_cluster_ctas"])
    # throughput with several independent trackers in flight (one context = one stream each): the one-cluster kernel
    # occupies 16 SMs per frame, so eight frames run side by side; the cooperative kernel wants most of the GPU per frame
    nb = 8
    ctxs = [capi.Context(0) for _ in range(nb)]
    trks, frs = [], []
    for cx in ctxs:
        t_ = capi.CoarseTracker(cx, tp.cam.w, tp.cam.h); t_.makeK(tp.cam, tp.levels); t_.setCoarseTrackingRef(tp.pc)
        trks.append(t_); frs.append(capi.Frame(cx, tp.img_new, tp.levels))
    for mode, key in ((trk.get_mode(), "tracking_8_streams_ms_per_frame"), (0, "tracking_8_streams_ms_per_frame_cooperative")):
        for t_ in trks:
            t_.set_mode(mode)
        for _ in range(2):
            for t_, f_ in zip(trks, frs):
                t_.track_enqueue(f_, T0, (0.0, 0.0), tp.levels - 1)
            for t_ in trks:
                t_.track_finish()
        t0 = time.perf_counter()
        for _ in range(reps):
            for t_, f_ in zip(trks, frs):
                t_.track_enqueue(f_, T0, (0.0, 0.0), tp.levels - 1)
            for t_ in trks:
                t_.track_finish()
        out[key] = 1e3 * (time.perf_counter() - t0) / (reps * nb)
    for t_, f_ in zip(trks, frs):
        t_.close(); f_.close()
    for cx in ctxs:
        cx.close()
    # batched entry point (hslam_tracker_track_batch): n members — motion hypotheses of one frame (System.cpp:348-481) or
    # several frames — in ONE launch, one 16-CTA cluster each; per-member results are bit-identical to single calls
    trk.set_mode(out["tracking_mode_cluster_ctas"])
    its = None
    if trk.get_mode() > 0:
        for nb_ in (8, 32):
            args_ = ([frame] * nb_, [T0] * nb_, [(0.0, 0.0)] * nb_, tp.levels - 1)
            for _ in range(2):
                r_ = trk.trackNewestCoarse_batch(*args_)
            t0 = time.perf_counter()
            for _ in range(reps):
                r_ = trk.trackNewestCoarse_batch(*args_)
            out[f"tracking_batch{nb_}_ms_per_frame"] = 1e3 * (time.perf_counter() - t0) / (reps * nb_)
            its = r_["iterations"][0]
        # the same 32 members with 8-CTA clusters (twice as many clusters resident at once)
        try:
            trk.set_mode(8)
            args_ = ([frame] * 32, [T0] * 32, [(0.0, 0.0)] * 32, tp.levels - 1)
            for _ in range(2):
                trk.trackNewestCoarse_batch(*args_)
            t0 = time.perf_counter()
            for _ in range(reps):
                trk.trackNewestCoarse_batch(*args_)
            out["tracking_batch32_cluster8_ms_per_frame"] = 1e3 * (time.perf_counter() - t0) / (reps * 32)
        except Exception as e:
            out["tracking_batch32_cluster8_error"] = repr(e)
        trk.set_mode(out["tracking_mode_cluster_ctas"])
    # roofline of Stage T: 64 algorithmic bytes per point-evaluation, evaluations ~ sum_l (1 + iterations_l) * N_l
    if its is not None:
        peak_ = 6555.2
        try:
            peak_ = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        evals = sum((1 + int(i)) * len(p_[0]) for i, p_ in zip(its, tp.pc))
        out["tracking_point_evaluations_per_frame"] = int(evals)
        for key in ("tracking_ms_per_frame_device", "tracking_8_streams_ms_per_frame", "tracking_batch8_ms_per_frame", "tracking_batch32_ms_per_frame",
                    "tracking_batch32_cluster8_ms_per_frame"):
            if key in out:
                out[key.replace("_ms_per_frame", "") + "_roofline_frac"] = evals * 64 / (out[key] * 1e-3) / 1e9 / peak_
        out["tracking_roofline_note"] = ("64 B per point-evaluation / measured HBM peak; a single frame is latency-bound by construction "
                                         "(SURVEY 8d), the batched launch is where the fraction is meaningful")
    # cfg1 as the reference builds it at 640x480: 4 pyramid levels (CalibData.h:110), next to the 5-level variant above
    try:
        tp4 = synth.make_tracking_pair(levels=4)
        trk4 = capi.CoarseTracker(ctx, tp4.cam.w, tp4.cam.h); trk4.makeK(tp4.cam, 4); trk4.setCoarseTrackingRef(tp4.pc)
        fr4_ = capi.Frame(ctx, tp4.img_new, 4)
        for _ in range(3):
            trk4.trackNewestCoarse(fr4_, T0, (0.0, 0.0), 3)
        e0.record(stream)
        for _ in range(reps):
            trk4.track_enqueue(fr4_, T0, (0.0, 0.0), 3)
        e1.record(stream)
        torch.cuda.synchronize()
        trk4.track_finish()
        out["tracking_4levels_ms_per_frame_device"] = e0.elapsed_time(e1) / reps
        out["tracking_4levels_iterations"] = [int(i) for i in trk4.iterations[:4]]
        trk4.close(); fr4_.close()
    except Exception as e:
        out["tracking_4levels_error"] = repr(e)
    t0 = time.perf_counter()
    for _ in range(reps):
        frame.update(tp.img_new)                                   # H2D 1.2 MB + 5-level pyramid
        trk.trackNewestCoarse(frame, T0, (0.0, 0.0), tp.levels - 1)  # LM on device + D2H of the result
    out["tracking_ms_per_frame_e2e"] = 1e3 * (time.perf_counter() - t0) / reps
    # §8(f) rows that run per frame / per keyframe: immature-point tracing and activation, reference-cloud build
    try:
        import copy
        kf = synth.render_keyframes(640, 480, 7, seed=78)
        hosts = [0, 1, 2, 3, 4]
        ims = synth.make_immature_points(kf, hosts=hosts, points_per_host=2000, seed=5)
        synth.trace_tables(kf, hosts, 5, ims)
        fr5 = capi.Frame(ctx, kf.images[5], 1)
        base = copy.deepcopy(ims)
        capi.traceNewCoarse(ctx, fr5, copy.deepcopy(base))
        t0 = time.perf_counter()
        for _ in range(5):
            work = copy.deepcopy(base)
            capi.traceNewCoarse(ctx, fr5, work)
        out["trace_ms_per_frame_e2e"] = 1e3 * (time.perf_counter() - t0) / 5
        out["trace_points"] = int(len(base.u))
        synth.trace_tables(kf, hosts, 6, work)
        fr6 = capi.Frame(ctx, kf.images[6], 1)
        capi.traceNewCoarse(ctx, fr6, work)
        keep = np.nonzero(np.isfinite(work.idepth_max) & (work.status != 1))[0]
        for k in ("host_idx", "u", "v", "gradH", "color", "weights", "energyTH", "idepth_min", "idepth_max"):
            setattr(work, k, np.ascontiguousarray(getattr(work, k)[keep]))
        frames7 = [capi.Frame(ctx, im, 1) for im in kf.images]
        T = synth.activation_tables(kf)
        capi.optimizeImmaturePoints(ctx, frames7, kf.cam, T, work, minObs=2)
        t0 = time.perf_counter()
        for _ in range(5):
            ok_, _, _ = capi.optimizeImmaturePoints(ctx, frames7, kf.cam, T, work, minObs=2)
        out["activation_ms_e2e"] = 1e3 * (time.perf_counter() - t0) / 5
        out["activation_candidates"] = int(len(keep)); out["activation_accepted"] = int(ok_.sum())
        prof0 = ctx.profile_read() if False else None
        ctx.profile_enable(True)
        capi.traceNewCoarse(ctx, fr5, copy.deepcopy(base))
        capi.optimizeImmaturePoints(ctx, frames7, kf.cam, T, work, minObs=2)
        pr = ctx.profile_read(); ctx.profile_enable(False)
        if "trace" in pr:
            out["trace_plus_activation_kernels_ms_device"] = pr["trace"][0]
        for f_ in frames7 + [fr5, fr6]:
            f_.close()
    except Exception as e:   # measurement extras must not take the bench line down
        out["trace_error"] = repr(e)
    out["tracking_iterations"] = [int(i) for i in trk.iterations]
    out["tracking_points_per_level"] = [int(len(p[0])) for p in tp.pc]
    trk.close(); frame.close()
    return out


def cpu_baseline(synth, win, is_lin):
    from oracle import oraclelib as O
    cores = os.cpu_count() or 1
    dIs = [O.pyramid(im, 1)[0][0] for im in win.images]
    ba = O.BAOracle(win, dIs, prec="f", threads=cores, fast=True)
    ba.linearize(False); ba.apply_res()
    ba.fix_linearization(np.nonzero(is_lin)[0].astype(np.int32))

    def step():
        ba.linearize(False); ba.new_frame_energy_th(win.n - 1); ba.apply_res()
        ba.accumulate(0); ba.accumulate(1); ba.accumulate(2)

    step()
    reps, t0 = 0, time.perf_counter()
    while reps < 200 and (time.perf_counter() - t0 < 10.0 or reps < 5):
        step(); reps += 1
    dt = time.perf_counter() - t0
    # Stage T on the CPU: the oracle's trackNewestCoarse (single-threaded like the reference's, CoarseTracker.cpp:506-683), cfg1
    trk_cpu = None
    try:
        tp = synth.make_tracking_pair()
        orc = O.TrackerOracle(tp.cam, tp.levels, tp.pc, fast=True)
        _, _, dI = O.pyramid(tp.img_new, tp.levels)
        T0 = np.concatenate([np.eye(3).reshape(9), np.zeros(3)])
        orc.track(dI, T0, (0.0, 0.0), tp.levels - 1, prec="f")
        tt0, nrep = time.perf_counter(), 0
        while nrep < 50 and (time.perf_counter() - tt0 < 3.0 or nrep < 3):
            r_ = orc.track(dI, T0, (0.0, 0.0), tp.levels - 1, prec="f"); nrep += 1
        trk_cpu = {"tracking_ms_per_frame": 1e3 * (time.perf_counter() - tt0) / nrep, "cores": 1,
                   "sample": f"{nrep} x trackNewestCoarse on cfg1 (640x480, 5 levels), oracle port -O3 -march=native, single thread"}
    except Exception as e:
        trk_cpu = {"error": repr(e)}
    return {"value": win.R * 8 * reps / dt / 1e6, "unit": UNIT, "cores": cores, "kind": "port", "tracking": trk_cpu,
            "ms_per_window": 1e3 * dt / reps,
            "sample": f"{reps} GN iterations of one cfg2 window (84k residuals, 30 % linearised) on {cores} threads, oracle port -O3 -march=native"}


if __name__ == "__main__":
    main()
