"""Committed golden vectors (tests/golden/*.npz, produced by tests/golden/make_golden.py from the oracle,
because the reference ships none): the oracle must keep reproducing them bit for bit (CPU), and the CUDA
path must match them (GPU) — independently of the synthetic generator staying unchanged."""
import os
import types

import numpy as np
import pytest

from conftest import rel_fro

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gt():
    return np.load(os.path.join(HERE, "golden", "tracking_small.npz"))


@pytest.fixture(scope="module")
def gb():
    return np.load(os.path.join(HERE, "golden", "ba_small.npz"))


def _cam(synth, c):
    return synth.Camera(int(c[0]), int(c[1]), np.float32(c[2]), np.float32(c[3]), np.float32(c[4]), np.float32(c[5]))


def _pc(g):
    return [tuple(g[f"pc{l}_{k}"] for k in "uvic") for l in range(int(g["levels"]))]


def _window(synth, g):
    w = types.SimpleNamespace()
    for k in ["T_w2c_eval", "state", "state_zero", "ab_exposure", "frameEnergyTH", "precalc", "adHost", "adTarget",
              "adHTdeltaF", "cDeltaF", "cPrior", "framePrior", "frameDeltaPrior", "u", "v", "idepth", "idepth_zero",
              "color", "weights", "priorF", "deltaF", "host_idx", "res_begin", "res_index", "point_idx", "target_idx",
              "state_state", "is_linearized", "is_active"]:
        setattr(w, k, g[k])
    w.cam = _cam(synth, g["cam"])
    w.images = list(g["images"])
    w.n = len(w.images)
    w.P, w.R = len(w.u), len(w.point_idx)
    return w


T_ID = np.concatenate([np.eye(3).reshape(9), np.zeros(3)])


# ------------------------------------------------------------------ CPU: the oracle reproduces its goldens
def test_oracle_reproduces_tracking_golden(O, synth, gt):
    cam = _cam(synth, gt["cam"])
    levels = int(gt["levels"])
    orc = O.TrackerOracle(cam, levels, _pc(gt))
    lv, _, dI = O.pyramid(gt["img_new"], levels)
    for name, T, aff in (("id", T_ID, (0.0, 0.0)), ("t1", gt["T1"], (0.01, 1.0))):
        for l in range(levels):
            f6, wf, nwf, mask = orc.calc_res(l, lv[l], T, aff, 20.0, "f")
            assert np.array_equal(f6, gt[f"res_{name}_{l}_f6"], equal_nan=True)
            assert np.array_equal(mask, gt[f"res_{name}_{l}_mask"]) and nwf == int(gt[f"res_{name}_{l}_nw"])
            Hf, bf = orc.calc_gs(l, wf, nwf, aff, "f")
            assert np.array_equal(Hf, gt[f"gs_{name}_{l}_Hf"]) and np.array_equal(bf, gt[f"gs_{name}_{l}_bf"])
    r = orc.track(dI, T_ID, (0.0, 0.0), levels - 1, prec="f")
    assert np.array_equal(r["T"], gt["track_f_T"]) and list(r["iterations"]) == list(gt["track_f_its"])


def test_oracle_reproduces_ba_golden(O, synth, gb):
    win = _window(synth, gb)
    dIs = [O.pyramid(im, 1)[0][0] for im in win.images]
    ba = O.BAOracle(win, dIs, prec="f")
    l1 = ba.linearize(False); ba.apply_res()
    ba.fix_linearization(np.nonzero(gb["is_lin"])[0].astype(np.int32))
    ba.update_points(idepth=gb["idepth2"], deltaF=(gb["idepth2"] - win.idepth_zero).astype(np.float32))
    l2 = ba.linearize(False); ba.apply_res()
    assert np.array_equal(l1["newState"], gb["f_lin0_state"]) and np.array_equal(l2["newState"], gb["f_lin1_state"])
    assert np.array_equal(l2["newEnergy"], gb["f_lin1_newEnergy"])
    for k, nm in enumerate("ALS"):
        H, b, nres = ba.accumulate(k)
        assert np.array_equal(H, gb[f"f_H{nm}"]) and np.array_equal(b, gb[f"f_b{nm}"]) and nres == int(gb[f"f_nres{nm}"])


# ------------------------------------------------------------------ GPU vs goldens
@pytest.mark.gpu
def test_gpu_tracking_vs_golden(ctx, capi, synth, O, gt):
    cam = _cam(synth, gt["cam"])
    levels = int(gt["levels"])
    trk = capi.CoarseTracker(ctx, cam.w, cam.h)
    trk.makeK(cam, levels)
    trk.setCoarseTrackingRef(_pc(gt))
    frame = capi.Frame(ctx, gt["img_new"], levels)
    for name, T, aff in (("id", T_ID, (0.0, 0.0)), ("t1", gt["T1"], (0.01, 1.0))):
        for l in range(levels):
            g6, nw = trk.calcRes(l, frame, T, aff, 20.0)
            f6 = gt[f"res_{name}_{l}_f6"]
            assert g6[1] == f6[1] and nw == int(gt[f"res_{name}_{l}_nw"]) and g6[5] == f6[5]
            if np.array_equal(gt[f"res_{name}_{l}_mask"], gt[f"res_{name}_{l}_maskd"]):
                Hg, bg = trk.calcGSSSE(l, frame, T, aff, 20.0)
                Hd, bd, Hf, bf = (gt[f"gs_{name}_{l}_{k}"] for k in ("Hd", "bd", "Hf", "bf"))
                eg, ef = rel_fro(np.c_[Hg, bg], np.c_[Hd, bd]), rel_fro(np.c_[Hf, bf], np.c_[Hd, bd])
                assert eg <= max(1e-5, 1.5 * ef), (name, l, eg, ef)
    ok, T, aff = trk.trackNewestCoarse(frame, T_ID, (0.0, 0.0), levels - 1)
    assert ok == bool(gt["track_d_ok"])
    assert list(trk.iterations[:levels]) == list(gt["track_f_its"][:levels])
    dg = np.abs(O.se3_log(T) - O.se3_log(gt["track_d_T"])).max()
    df = np.abs(O.se3_log(gt["track_f_T"]) - O.se3_log(gt["track_d_T"])).max()
    assert dg <= max(1e-6, 2 * df), (dg, df)
    trk.close(); frame.close()


@pytest.mark.gpu
def test_gpu_ba_vs_golden(ctx, capi, synth, gb):
    win = _window(synth, gb)
    frames = [capi.Frame(ctx, im, 1) for im in win.images]
    ef = capi.EnergyFunctional(ctx)
    ef.set_window(win, frames)
    ef.set_topology(win)
    l1 = ef.linearizeAll(False); ef.applyRes()
    ef.fixLinearization(np.nonzero(gb["is_lin"])[0].astype(np.int32))
    ef.update_points(idepth=gb["idepth2"], deltaF=(gb["idepth2"] - win.idepth_zero).astype(np.float32))
    l2 = ef.linearizeAll(False); ef.applyRes()
    assert np.array_equal(l1["newState"], gb["f_lin0_state"]) and np.array_equal(l2["newState"], gb["f_lin1_state"])
    assert np.array_equal(l2["newEnergy"], gb["f_lin1_newEnergy"])
    assert np.array_equal(l2["newEnergyWithOutlier"], gb["f_lin1_newEnergyWO"])
    assert l2["energy"] == pytest.approx(float(gb["f_lin1_energy"]), rel=1e-12)
    rr = ef.residual_results()
    ok = rr["state_state"] != 1
    assert np.array_equal(rr["is_active"], gb["f_active"])
    assert np.array_equal(rr["J"][ok], gb["f_J"][ok]) and np.array_equal(rr["JpJdF"][ok], gb["f_JpJdF"][ok])
    sys3 = ef.accumulate_all()
    for (H, b), nm in zip(sys3, "ALS"):
        Hd, bd, Hf, bf = gb[f"d_H{nm}"], gb[f"d_b{nm}"], gb[f"f_H{nm}"], gb[f"f_b{nm}"]
        assert rel_fro(H, Hd) <= max(1e-5, 1.5 * rel_fro(Hf, Hd)), nm
        assert rel_fro(b, bd) <= max(1e-5, 1.5 * rel_fro(bf, bd)), nm
    assert (ef.resInA, ef.resInL) == (int(gb["f_nresA"]), int(gb["f_nresL"]))
    pr = ef.point_results()
    assert np.array_equal(pr["HdiF"], gb["f_HdiF"].astype(np.float32))
    assert np.array_equal(pr["bdSumF"], gb["f_bdSumF"].astype(np.float32))
    ef.close()
    for f in frames:
        f.close()
