import importlib, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
capi = importlib.import_module("h-slam_b200.capi"); synth = importlib.import_module("h-slam_b200.synth")
from oracle import oraclelib as O
ctx = capi.Context(0)
win, is_lin = synth.make_window(points_per_frame=300, frac_linearized=0.3)
frames = [capi.Frame(ctx, im, 1) for im in win.images]
dIs = [O.pyramid(im, 1)[0][0] for im in win.images]
ef = capi.EnergyFunctional(ctx); ef.set_window(win, frames); ef.set_topology(win)
od = O.BAOracle(win, dIs, prec="d")
ef.linearizeAll(False); ef.applyRes(); od.linearize(False); od.apply_res()
sys3 = ef.accumulate_all()
n = win.n; N = 8 * n + 4
for k, nm in enumerate("ALS"):
    Ho, bo, _ = od.accumulate(k)
    H, b = sys3[k]
    print(nm, "rel", np.linalg.norm(H - Ho) / np.linalg.norm(Ho), np.linalg.norm(b - bo) / max(np.linalg.norm(bo), 1e-30))
    if nm == "S":
        blk = lambda M, i, j: M[(0 if i < 0 else 4 + 8 * i):(4 if i < 0 else 12 + 8 * i), (0 if j < 0 else 4 + 8 * j):(4 if j < 0 else 12 + 8 * j)]
        for i in range(-1, n):
            print(" ".join("%8.1e" % (np.linalg.norm(blk(H, i, j) - blk(Ho, i, j)) / max(np.linalg.norm(blk(Ho, i, j)), 1e-30)) for j in range(-1, n)))
pr = ef.point_results(); po = od.point_results()
for k in pr:
    a, b = np.asarray(pr[k], np.float64), np.asarray(po[k], np.float64).reshape(np.asarray(pr[k]).shape)
    print(k, np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
rr = ef.residual_results(); ro = od.residual_results()
print("JpJdF", np.abs(rr["JpJdF"] - ro["JpJdF"]).max() / np.abs(ro["JpJdF"]).max(), "active", (rr["is_active"] != ro["is_active"]).sum())
