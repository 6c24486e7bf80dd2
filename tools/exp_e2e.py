"""Experiment: where the e2e step time goes (host-side timers around the two batch calls)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
capi = importlib.import_module("h-slam_b200.capi"); synth = importlib.import_module("h-slam_b200.synth")
ctx = capi.Context(0)
wins = bench.build_windows(synth, 8, seed0=100)
efs, frs = [], []
for w in wins:
    fr = [capi.Frame(ctx, im, 1) for im in w.images]
    ef = capi.EnergyFunctional(ctx); ef.set_window(w, fr); ef.set_topology(w); efs.append(ef); frs.append(fr)
batch = capi.EnergyFunctionalBatch(ctx, efs)
io = batch.make_io(wins, frs, pinned="bound")
xs = [np.ascontiguousarray(np.random.default_rng(i).normal(0, 1e-4, 60)) for i in range(8)]
for _ in range(5):
    batch.step_host(io); batch.resubstitute_host(io, xs)
t1 = t2 = t3 = 0.0
K = 50
for _ in range(K):
    a = time.perf_counter(); batch.step_host(io); b = time.perf_counter(); batch.resubstitute_host(io, xs); c = time.perf_counter()
    batch.iteration(); ctx.synchronize(); d = time.perf_counter()
    t1 += b - a; t2 += c - b; t3 += d - c
print(f"step_host {t1/K*1e3:.3f} ms, resubstitute_host {t2/K*1e3:.3f} ms, device-only iteration+sync {t3/K*1e3:.3f} ms")
io2 = batch.make_io(wins, frs, pinned="bound")
for o in io2["arr"]:
    o.window = None; o.idepth = None; o.deltaF = None
t1 = 0.0
for _ in range(K):
    a = time.perf_counter(); batch.step_host(io2); t1 += time.perf_counter() - a
print(f"step_host without inputs {t1/K*1e3:.3f} ms")
