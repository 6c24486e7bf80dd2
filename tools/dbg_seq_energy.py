import importlib, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
capi = importlib.import_module("h-slam_b200.capi"); sb = importlib.import_module("h-slam_b200.sequence_ba"); synth = importlib.import_module("h-slam_b200.synth")
from oracle import oraclelib as O
ctx = capi.Context(0)
cam, Ts, frames = sb.render_sequence(320, 240, 8, seed=21)
rng = np.random.default_rng(5)
kfs = []
for k in (0, 6):
    img, depth = frames[k]
    u, v, idp, col, wts = sb.select_points(img, depth, 300, rng)
    st = np.zeros(10)
    kfs.append(sb.KeyFrame(k, img, None, np.array(Ts[k], np.float64), st.copy(), st.copy(), 512.0, u, v, idp, col, wts))
win = sb.build_window(cam, kfs)
fr = [capi.Frame(ctx, q.img, 4) for q in kfs]
ef = capi.EnergyFunctional(ctx); ef.set_window(win, fr); ef.set_topology(win)
g = ef.linearizeAll(False, outputs="states")
g2 = ef.linearizeAll(False)
ob = O.BAOracle(win, [O.pyramid(q.img, 4)[0][0] for q in kfs], prec="f")
f = ob.linearize(False)
print("energy gpu", g["energy"], g2["energy"], "orc", f["energy"], "sum newEnergy gpu", g2["newEnergy"].astype(np.float64).sum(), "orc", f["newEnergy"].astype(np.float64).sum())
print("states equal", np.array_equal(g["newState"], f["newState"]), "newEnergy equal", np.array_equal(g2["newEnergy"], f["newEnergy"]), np.bincount(f["newState"], minlength=3))
d = np.nonzero(g2["newEnergy"] != f["newEnergy"])[0]
print("diff idx", d[:10], g2["newEnergy"][d[:10]], f["newEnergy"][d[:10]], f["newState"][d[:10]])
oob = f["newState"] == 1
print("oob energy sum orc", f["newEnergy"][oob].sum(), "gpu", g2["newEnergy"][oob].sum())
