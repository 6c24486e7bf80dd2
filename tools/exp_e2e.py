"""Experiment: where the e2e step time goes (host-side timers around the two batch calls).
usage: python tools/exp_e2e.py [windows]"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
capi = importlib.import_module("h-slam_b200.capi"); synth = importlib.import_module("h-slam_b200.synth")
NW = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ctx = capi.Context(0)
wins, lins = bench.build_windows(synth, NW, seed0=100)
efs, frs = [], []
for w in wins:
    fr = [capi.Frame(ctx, im, 1) for im in w.images]
    ef = capi.EnergyFunctional(ctx); ef.set_window(w, fr); ef.set_topology(w); ef.set_jacobian_mode(1); efs.append(ef); frs.append(fr)
batch = capi.EnergyFunctionalBatch(ctx, efs)
batch.iteration(); ctx.synchronize()
for ef, l in zip(efs, lins):
    ef.fixLinearization(np.nonzero(l)[0].astype(np.int32))
xs = [np.ascontiguousarray(np.random.default_rng(i).normal(0, 1e-4, 60)) for i in range(NW)]
K = 20


def timed(label, io, resub=True):
    for _ in range(3):
        batch.step_host(io)
        if resub:
            batch.resubstitute_host(io, xs)
    t1 = t2 = 0.0
    for _ in range(K):
        a = time.perf_counter(); batch.step_host(io); b = time.perf_counter()
        if resub:
            batch.resubstitute_host(io, xs)
        c = time.perf_counter()
        t1 += b - a; t2 += c - b
    print(f"{label:42s} step_host {t1/K*1e3:.3f} ms  resubstitute_host {t2/K*1e3:.3f} ms", flush=True)


io = batch.make_io(wins, frs, pinned="bound")
timed("all inputs, all outputs", io)
ios = batch.make_io(wins, frs, pinned="bound")
for o in ios["arr"]:
    o.newEnergy = None; o.newEnergyWithOutlier = None
timed("all inputs, slim outputs (device TH)", ios)
t3 = 0.0
for _ in range(K):
    c = time.perf_counter(); batch.iteration(); ctx.synchronize(); t3 += time.perf_counter() - c
print(f"device-only iteration + sync {t3/K*1e3:.3f} ms", flush=True)
io2 = batch.make_io(wins, frs, pinned="bound")
for o in io2["arr"]:
    o.window = None
timed("no tables (idepth/deltaF only)", io2, resub=False)
for o in io2["arr"]:
    o.idepth = None; o.deltaF = None
timed("no inputs at all", io2, resub=False)
for o in io2["arr"]:
    o.newEnergy = None; o.newEnergyWithOutlier = None
timed("no inputs, states+systems only", io2, resub=False)
for o in io2["arr"]:
    o.newState = None
timed("no inputs, systems only", io2, resub=False)
io3 = batch.make_io(wins, frs, pinned="bound")
for o in io3["arr"]:
    o.newEnergy = None; o.newEnergyWithOutlier = None; o.newState = None
timed("all inputs, systems only", io3, resub=False)
