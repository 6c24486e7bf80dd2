#!/usr/bin/env python
"""Generates tests/golden/*.npz.

The reference ships no golden vectors for this path and cannot be run here, so these
fixtures are produced by the CPU oracle (float-faithful 'f' and fp64 'd' restatements) on small synthetic
inputs.  They pin the oracle against regressions and give the GPU tests size-independent, seed-independent
inputs that do not depend on the synthetic generator staying unchanged.

    python tests/golden/make_golden.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    synth = importlib.import_module("h-slam_b200.synth")
    from oracle import oraclelib as O

    # ------------------------------------------------------------------ tracking
    tp = synth.make_tracking_pair(w=160, h=120, levels=3, n_seed_points=500, seed=21)
    orc = O.TrackerOracle(tp.cam, tp.levels, tp.pc)
    lv, _, dI = O.pyramid(tp.img_new, tp.levels)
    T_id = np.concatenate([np.eye(3).reshape(9), np.zeros(3)])
    R, t = synth.se3_exp([0.006, -0.003, 0.008, 0.002, -0.003, 0.001])
    T1 = synth.T12(R, t)
    out = dict(img_ref=tp.img_ref, img_new=tp.img_new, cam=np.array([tp.cam.w, tp.cam.h, tp.cam.fx, tp.cam.fy, tp.cam.cx, tp.cam.cy], np.float64),
               levels=np.int32(tp.levels), T1=T1)
    for l in range(tp.levels):
        for k, a in zip("uvic", tp.pc[l]):
            out[f"pc{l}_{k}"] = a
    for name, T, aff in (("id", T_id, (0.0, 0.0)), ("t1", T1, (0.01, 1.0))):
        for l in range(tp.levels):
            f6, wf, nwf, mask = orc.calc_res(l, lv[l], T, aff, 20.0, "f")
            d6, wd, nwd, maskd = orc.calc_res(l, lv[l], T, aff, 20.0, "d")
            Hd, bd = orc.calc_gs(l, wd, nwd, aff, "d")
            Hf, bf = orc.calc_gs(l, wf, nwf, aff, "f")
            out[f"res_{name}_{l}_f6"] = f6; out[f"res_{name}_{l}_d6"] = d6
            out[f"res_{name}_{l}_mask"] = mask; out[f"res_{name}_{l}_maskd"] = maskd
            out[f"res_{name}_{l}_nw"] = np.int32(nwf)
            out[f"gs_{name}_{l}_Hd"] = Hd; out[f"gs_{name}_{l}_bd"] = bd
            out[f"gs_{name}_{l}_Hf"] = Hf; out[f"gs_{name}_{l}_bf"] = bf
    for prec in "fd":
        r = orc.track(dI, T_id, (0.0, 0.0), tp.levels - 1, prec=prec)
        out[f"track_{prec}_T"] = r["T"]; out[f"track_{prec}_aff"] = r["aff"]
        out[f"track_{prec}_res"] = r["lastResiduals"]; out[f"track_{prec}_its"] = r["iterations"]
        out[f"track_{prec}_ok"] = np.int32(r["ok"])
    np.savez_compressed(os.path.join(HERE, "tracking_small.npz"), **out)

    # ------------------------------------------------------------------ window BA
    win, is_lin = synth.make_window(w=160, h=120, n_frames=3, points_per_frame=60, seed=33, frac_linearized=0.3)
    dIs = [O.pyramid(im, 1)[0][0] for im in win.images]
    fields = ["T_w2c_eval", "state", "state_zero", "ab_exposure", "frameEnergyTH", "precalc", "adHost", "adTarget",
              "adHTdeltaF", "cDeltaF", "cPrior", "framePrior", "frameDeltaPrior", "u", "v", "idepth", "idepth_zero",
              "color", "weights", "priorF", "deltaF", "host_idx", "res_begin", "res_index", "point_idx", "target_idx",
              "state_state", "is_linearized", "is_active"]
    out = {k: np.asarray(getattr(win, k)) for k in fields}
    out["images"] = np.stack(win.images)
    out["cam"] = np.array([win.cam.w, win.cam.h, win.cam.fx, win.cam.fy, win.cam.cx, win.cam.cy], np.float64)
    out["is_lin"] = is_lin
    idp2 = (win.idepth * np.float32(1.001)).astype(np.float32)
    out["idepth2"] = idp2
    for prec in "fd":
        ba = O.BAOracle(win, dIs, prec=prec)
        l1 = ba.linearize(False); ba.apply_res()
        ba.fix_linearization(np.nonzero(is_lin)[0].astype(np.int32))
        ba.update_points(idepth=idp2, deltaF=(idp2 - win.idepth_zero).astype(np.float32))
        l2 = ba.linearize(False); ba.apply_res()
        for i, l in enumerate((l1, l2)):
            out[f"{prec}_lin{i}_state"] = l["newState"]; out[f"{prec}_lin{i}_energy"] = np.float64(l["energy"])
            out[f"{prec}_lin{i}_newEnergy"] = l["newEnergy"]; out[f"{prec}_lin{i}_newEnergyWO"] = l["newEnergyWithOutlier"]
        for k, nm in enumerate("ALS"):
            H, b, nres = ba.accumulate(k)
            out[f"{prec}_H{nm}"] = H; out[f"{prec}_b{nm}"] = b; out[f"{prec}_nres{nm}"] = np.int32(nres)
        rr = ba.residual_results()
        out[f"{prec}_J"] = rr["J"].astype(np.float32 if prec == "f" else np.float64)
        out[f"{prec}_JpJdF"] = rr["JpJdF"].astype(np.float32 if prec == "f" else np.float64)
        out[f"{prec}_active"] = rr["is_active"]
        pr = ba.point_results()
        out[f"{prec}_HdiF"] = pr["HdiF"]; out[f"{prec}_bdSumF"] = pr["bdSumF"]
    np.savez_compressed(os.path.join(HERE, "ba_small.npz"), **out)
    for f in ("tracking_small.npz", "ba_small.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
