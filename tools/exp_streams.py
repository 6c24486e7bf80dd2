"""Experiment: the bench batch split into G groups on G contexts/streams (kernels of different groups overlap)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
capi = importlib.import_module("h-slam_b200.capi")
synth = importlib.import_module("h-slam_b200.synth")
wins = bench.build_windows(synth, 8, seed0=100)
for G in (1, 2, 4, 8):
    ctxs = [capi.Context(0) for _ in range(G)]
    batches, keep = [], []
    for g in range(G):
        efs = []
        for w in wins[g::G]:
            fr = [capi.Frame(ctxs[g], im, 1) for im in w.images]
            ef = capi.EnergyFunctional(ctxs[g]); ef.set_window(w, fr); ef.set_topology(w)
            efs.append(ef); keep.append(fr)
        batches.append(capi.EnergyFunctionalBatch(ctxs[g], efs))
    for _ in range(5):
        for b in batches: b.iteration()
    for c in ctxs: c.synchronize()
    K = 50
    t0 = time.perf_counter()
    for _ in range(K):
        for b in batches: b.iteration()
    for c in ctxs: c.synchronize()
    dt = (time.perf_counter() - t0) / K
    print(f"G={G}: {dt*1e3:.4f} ms per 8-window step", flush=True)
