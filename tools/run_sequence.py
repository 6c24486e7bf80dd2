#!/usr/bin/env python
"""BASELINE.json configs[2]: 640x480 rendered sequence (default 500 frames), CUDA tracker vs the CPU oracle with
identical decisions; prints one JSON line with both ATEs, their relative difference and the per-frame timings.
    python tools/run_sequence.py --frames 500 > profiles/r01_sequence.json      (on the GPU box)"""
import argparse, importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=500); ap.add_argument("--w", type=int, default=640); ap.add_argument("--h", type=int, default=480)
ap.add_argument("--levels", type=int, default=5); ap.add_argument("--kf-every", type=int, default=8); ap.add_argument("--points", type=int, default=6200)
ap.add_argument("--ba", action="store_true", help="both stages: tracking + windowed BA of every keyframe (sequence_ba.py)")
ap.add_argument("--pts-per-kf", type=int, default=2000); ap.add_argument("--cpu-threads", type=int, default=os.cpu_count() or 1)
a = ap.parse_args()
if a.ba:
    capi = importlib.import_module("h-slam_b200.capi"); sb = importlib.import_module("h-slam_b200.sequence_ba")
    from oracle import oraclelib as O
    from seq_backends import OracleBackend
    t0 = time.time(); cam, Ts, frames = sb.render_sequence(a.w, a.h, a.frames, seed=21); t_render = time.time() - t0
    ctx = capi.Context(0)
    kw = dict(kf_every=a.kf_every, pts_per_kf=a.pts_per_kf)
    sb.run_sequence_ba(cam, Ts[:20], frames[:20], a.levels, sb.CudaBackend(capi, ctx), **kw)     # warm-up (module load, first launches)
    t0 = time.time(); g = sb.run_sequence_ba(cam, Ts, frames, a.levels, sb.CudaBackend(capi, ctx), **kw); t_gpu = time.time() - t0
    t0 = time.time(); f = sb.run_sequence_ba(cam, Ts, frames, a.levels, OracleBackend(O, "f", threads=a.cpu_threads), **kw); t_cpu = time.time() - t0
    path = float(np.linalg.norm(np.diff(np.array([T[9:] for T in Ts]), axis=0), axis=1).sum())
    diff = tot = 0
    for sg, sf in zip(g.ba_states, f.ba_states):
        for x, y in zip(sg, sf):
            diff += int((x != y).sum()); tot += len(x)
    d = max(float(np.abs(np.asarray(x) - np.asarray(y)).max()) for x, y in zip(g.T_w2c_est, f.T_w2c_est))
    print(json.dumps({"config": f"{a.w}x{a.h}, {a.frames} frames, {a.levels} levels, keyframe every {a.kf_every}, window <= 7 keyframes, "
                                f"{a.pts_per_kf} points per keyframe, <= 6 GN iterations per keyframe; tracking + windowed BA in the loop",
                      "ate_rmse_cuda_m": g.ate_rmse, "ate_rmse_oracle_m": f.ate_rmse,
                      "rel_diff": abs(g.ate_rmse - f.ate_rmse) / max(f.ate_rmse, 1e-300), "path_length_m": path, "max_abs_pose_diff": d,
                      "all_ok": bool(all(g.ok) and all(f.ok)), "keyframes": g.n_keyframes, "gn_iterations_cuda": g.ba_iterations,
                      "gn_iterations_equal": g.ba_iterations == f.ba_iterations, "residual_states_compared": tot, "residual_states_different": diff,
                      "wall_s": {"render": t_render, "cuda_loop": t_gpu, "oracle_loop": t_cpu, "oracle_threads": a.cpu_threads,
                                 "cuda_breakdown": g.wall_s, "oracle_breakdown": f.wall_s}}))
    sys.exit(0)
capi = importlib.import_module("h-slam_b200.capi"); seq = importlib.import_module("h-slam_b200.sequence")
from oracle import oraclelib as O
from test_gpu_sequence import OracleTracker
t0 = time.time(); cam, Ts, frames = seq.render_sequence(a.w, a.h, a.frames, seed=21); t_render = time.time() - t0
ctx = capi.Context(0)
t0 = time.time(); g = seq.run_sequence(cam, Ts, frames, a.levels, lambda c, l, pc: seq.CudaTracker(capi, ctx, c, l, pc), a.kf_every, a.points); t_gpu = time.time() - t0
t0 = time.time(); f = seq.run_sequence(cam, Ts, frames, a.levels, lambda c, l, pc: OracleTracker(O, c, l, pc, "f"), a.kf_every, a.points); t_cpu = time.time() - t0
path = float(np.linalg.norm(np.diff(np.array([T[9:] for T in Ts]), axis=0), axis=1).sum())
d = max(float(np.abs(np.asarray(x) - np.asarray(y)).max()) for x, y in zip(g.T_w2c_est, f.T_w2c_est))
print(json.dumps({"config": f"{a.w}x{a.h}, {a.frames} frames, {a.levels} levels, keyframe every {a.kf_every}, {a.points} seed points",
                  "ate_rmse_cuda_m": g.ate_rmse, "ate_rmse_oracle_m": f.ate_rmse, "rel_diff": abs(g.ate_rmse - f.ate_rmse) / max(f.ate_rmse, 1e-300),
                  "path_length_m": path, "max_abs_pose_diff": d, "all_ok": bool(all(g.ok) and all(f.ok)), "keyframes": g.n_keyframes,
                  "wall_s": {"render": t_render, "cuda_driver_loop_incl_cloud_building": t_gpu, "oracle_driver_loop": t_cpu}}))
