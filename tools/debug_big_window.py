"""Debug/verification: the 2M-residual window on ONE GPU — full window vs the sum of its point shards."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
capi = importlib.import_module("h-slam_b200.capi")
synth = importlib.import_module("h-slam_b200.synth")
shard = importlib.import_module("h-slam_b200.shard")
ppf = int(os.environ.get("HSLAM_SHARDED_PPF", "47620"))
t0 = time.time()
kf = synth.render_keyframes(640, 480, 7, seed=4321)
win, _ = synth.make_window(points_per_frame=ppf, seed=900, keyframes=kf)
print("window built", win.P, win.R, time.time() - t0, flush=True)
ctx = capi.Context(0)
frames = [capi.Frame(ctx, im, 1) for im in win.images]
N = 8 * win.n + 4
per = N * N + N
def run(w, priors=True):
    ef = capi.EnergyFunctional(ctx)
    ef.set_window(w, frames); ef.set_topology(w); ef.set_add_priors(priors)
    ef.iteration_device()
    out = ef.download_Hb()
    flat = np.concatenate([np.concatenate([H.T.reshape(-1), b]) for H, b in out])
    ef.close()
    return flat
full = run(win)
print("full norm", np.linalg.norm(full), "nan", np.isnan(full).any(), flush=True)
for world in (2, 4):
    acc = np.zeros_like(full)
    for r in range(world):
        acc += run(shard.shard_window(win, r, world), priors=(r == 0))
    print("world", world, "rel err", np.linalg.norm(acc - full) / np.linalg.norm(full), flush=True)
