"""GPU parity: pyramid + CoarseTracker path through the C ABI vs the CPU oracle.

Tolerances (BASELINE.json north_star): residual counts / masks bit-exact; H,b <= 1e-5 relative
Frobenius vs the fp64 oracle and no worse than the reference's own float error; pose <= 1e-6.
"""
import numpy as np
import pytest

from conftest import rel_fro

pytestmark = pytest.mark.gpu

T_ID = np.concatenate([np.eye(3).reshape(9), np.zeros(3)])


@pytest.mark.parametrize("shape,levels", [((480, 640), 5), ((240, 320), 4), ((96, 112), 3)])
def test_pyramid_bit_exact(ctx, capi, O, shape, levels):
    rng = np.random.default_rng(3)
    img = (rng.random(shape) * 255).astype(np.float32)
    gamma = (0.5 + rng.random(256)).astype(np.float32)
    for g in (None, gamma):
        f = capi.Frame(ctx, img, levels, gamma=g)
        lv, ag, _ = O.pyramid(img, levels, gamma=g)
        for l in range(levels):
            dI, a = f.download(l)
            assert np.array_equal(dI, lv[l]), f"level {l} dI differs"
            assert np.array_equal(a, ag[l]), f"level {l} absSquaredGrad differs"
        f.close()


def _setup(ctx, capi, O, tp):
    trk = capi.CoarseTracker(ctx, tp.cam.w, tp.cam.h)
    trk.makeK(tp.cam, tp.levels)
    trk.setCoarseTrackingRef(tp.pc)
    frame = capi.Frame(ctx, tp.img_new, tp.levels)
    orc = O.TrackerOracle(tp.cam, tp.levels, tp.pc)
    lv, _, dI = O.pyramid(tp.img_new, tp.levels)
    return trk, frame, orc, lv, dI


def test_makeK_matches_oracle(ctx, capi, O, tracking_pair):
    tp = tracking_pair
    trk, frame, orc, lv, dI = _setup(ctx, capi, O, tp)
    for l in range(tp.levels):
        k, wh, ki = trk.get_K(l)
        ref = orc.level_K(l)
        assert np.array_equal(k, np.array([ref["fx"], ref["fy"], ref["cx"], ref["cy"]], np.float32))
        assert tuple(wh) == (ref["w"], ref["h"])
        assert np.array_equal(ki, ref["Ki"])
    trk.close(); frame.close()


def _poses(synth):
    out = [(T_ID, (0.0, 0.0))]
    R, t = synth.se3_exp([0.01, -0.004, 0.012, 0.003, -0.005, 0.002])
    out.append((synth.T12(R, t), (0.015, 2.0)))
    R, t = synth.se3_exp([0.05, 0.03, -0.04, 0.02, 0.015, -0.01])   # far off: many saturated / gated points
    out.append((synth.T12(R, t), (-0.05, -8.0)))
    return out


@pytest.mark.parametrize("which", ["full", "small"])
def test_calc_res_counts_exact_and_energy(ctx, capi, O, synth, tracking_pair, tracking_pair_small, which):
    tp = tracking_pair if which == "full" else tracking_pair_small
    trk, frame, orc, lv, dI = _setup(ctx, capi, O, tp)
    for T, aff in _poses(synth):
        for l in range(tp.levels):
            for cutoff in (20.0, 40.0):
                g6, gnw = trk.calcRes(l, frame, T, aff, cutoff)
                f6, _, fnw, fmask = orc.calc_res(l, lv[l], T, aff, cutoff, "f", want_warped=False)
                d6, _, dnw, dmask = orc.calc_res(l, lv[l], T, aff, cutoff, "d", want_warped=False)
                # integer outputs: bit-exact against the float-faithful restatement
                assert g6[1] == f6[1], (l, cutoff, g6, f6)
                assert gnw == fnw
                assert g6[5] == f6[5]
                # energy: tree-summed fp32 on the GPU must be at least as close to fp64 as the
                # reference's sequential float sum, and within 1e-5 when the masks agree
                if np.array_equal(fmask, dmask) and d6[0] > 0:
                    eg = abs(g6[0] - d6[0]) / d6[0]
                    ef = abs(f6[0] - d6[0]) / d6[0]
                    assert eg <= max(1e-5, 1.5 * ef), (l, eg, ef)
                if l == 0:
                    assert g6[2] == pytest.approx(d6[2], rel=1e-4, abs=1e-9)
                    assert g6[4] == pytest.approx(d6[4], rel=1e-4, abs=1e-9)
    trk.close(); frame.close()


@pytest.mark.parametrize("which", ["full", "small"])
def test_calc_gs_vs_double_oracle(ctx, capi, O, synth, tracking_pair, tracking_pair_small, which):
    tp = tracking_pair if which == "full" else tracking_pair_small
    trk, frame, orc, lv, dI = _setup(ctx, capi, O, tp)
    worst = 0.0
    for T, aff in _poses(synth)[:2]:
        for l in range(tp.levels):
            Hg, bg = trk.calcGSSSE(l, frame, T, aff, 20.0)
            _, wf, nwf, mf = orc.calc_res(l, lv[l], T, aff, 20.0, "f")
            _, wd, nwd, md = orc.calc_res(l, lv[l], T, aff, 20.0, "d")
            Hf, bf = orc.calc_gs(l, wf, nwf, aff, "f")
            Hd, bd = orc.calc_gs(l, wd, nwd, aff, "d")
            if not np.array_equal(mf, md):
                continue  # threshold-boundary case: float and double masks differ, documented in DESIGN.md
            eg = rel_fro(np.c_[Hg, bg], np.c_[Hd, bd])
            ef = rel_fro(np.c_[Hf, bf], np.c_[Hd, bd])
            worst = max(worst, eg)
            assert eg <= max(1e-5, 1.5 * ef), (l, eg, ef)
            assert rel_fro(Hg, Hd) <= 1e-5
            assert np.allclose(Hg, Hg.T)
    assert worst > 0  # at least one comparison happened
    trk.close(); frame.close()


def test_calc_gs_deterministic(ctx, capi, O, synth, tracking_pair):
    tp = tracking_pair
    trk, frame, orc, lv, dI = _setup(ctx, capi, O, tp)
    T, aff = _poses(synth)[1]
    H0, b0 = trk.calcGSSSE(0, frame, T, aff, 20.0)
    for _ in range(5):
        H1, b1 = trk.calcGSSSE(0, frame, T, aff, 20.0)
        assert np.array_equal(H0, H1) and np.array_equal(b0, b1)
    trk.close(); frame.close()


@pytest.mark.parametrize("which,coarsest", [("full", 4), ("full", 3), ("small", 3)])
def test_track_matches_oracle(ctx, capi, O, synth, tracking_pair, tracking_pair_small, which, coarsest):
    tp = tracking_pair if which == "full" else tracking_pair_small
    trk, frame, orc, lv, dI = _setup(ctx, capi, O, tp)
    ok, T, aff = trk.trackNewestCoarse(frame, T_ID, (0.0, 0.0), coarsest)
    rf = orc.track(dI, T_ID, (0.0, 0.0), coarsest, prec="f")
    rd = orc.track(dI, T_ID, (0.0, 0.0), coarsest, prec="d")
    assert ok == rf["ok"] == rd["ok"] is True
    # converged to the rendered motion
    xi_true = O.se3_log(tp.T_ref2new_true)
    assert np.allclose(O.se3_log(T), xi_true, atol=3e-4)
    # same LM path as the reference restatement (accept/reject decisions are E/N ratio compares)
    assert list(trk.iterations) == list(rf["iterations"]), (trk.iterations, rf["iterations"])
    # pose / affine increments: <= 1e-6 vs the fp64 oracle, no worse than the float restatement
    dg = np.abs(O.se3_log(T) - O.se3_log(rd["T"])).max()
    df = np.abs(O.se3_log(rf["T"]) - O.se3_log(rd["T"])).max()
    assert dg <= max(1e-6, 2 * df), (dg, df)
    assert abs(aff[0] - rd["aff"][0]) <= max(1e-6, 2 * abs(rf["aff"][0] - rd["aff"][0]))
    assert abs(aff[1] - rd["aff"][1]) <= max(1e-4, 2 * abs(rf["aff"][1] - rd["aff"][1]))
    assert np.allclose(trk.lastResiduals[: coarsest + 1], rd["lastResiduals"][: coarsest + 1], rtol=1e-4)
    assert np.allclose(trk.lastFlowIndicators, rd["flow"], rtol=1e-3)
    trk.close(); frame.close()


def test_track_abort_and_bad_args(ctx, capi, O, tracking_pair_small):
    tp = tracking_pair_small
    trk, frame, orc, lv, dI = _setup(ctx, capi, O, tp)
    # abort: minResForAbort tiny -> returns false at the coarsest level, pose untouched
    T0 = T_ID.copy()
    ok, T, aff = trk.trackNewestCoarse(frame, T0, (0.0, 0.0), 3, minResForAbort=np.full(5, 1e-3))
    r = orc.track(dI, T0, (0.0, 0.0), 3, min_res=np.full(5, 1e-3), prec="f")
    assert ok is False and r["ok"] is False
    assert np.array_equal(T, T0)
    assert np.isnan(trk.lastResiduals[0]) and not np.isnan(trk.lastResiduals[3])
    with pytest.raises(capi.HslamError):
        trk.trackNewestCoarse(frame, T0, (0.0, 0.0), 5)      # reference asserts coarsestLvl < 5
    with pytest.raises(capi.HslamError):
        trk.calcRes(9, frame, T0, (0.0, 0.0), 20.0)
    trk.close(); frame.close()


def test_empty_reference_cloud(ctx, capi, O, tracking_pair_small):
    tp = tracking_pair_small
    trk = capi.CoarseTracker(ctx, tp.cam.w, tp.cam.h)
    trk.makeK(tp.cam, tp.levels)
    e = np.zeros(0, np.float32)
    trk.setCoarseTrackingRef([(e, e, e, e)] * tp.levels)
    frame = capi.Frame(ctx, tp.img_new, tp.levels)
    out6, nw = trk.calcRes(0, frame, T_ID, (0.0, 0.0), 20.0)
    assert out6[0] == 0 and out6[1] == 0 and nw == 0 and np.isnan(out6[5])   # 0/0 as in the reference
    trk.close(); frame.close()


def test_make_coarse_depth_on_device_bit_exact(ctx, capi, O, synth):
    """CoarseTracker::makeCoarseDepthL0 on the device (splat with pixel collisions, sum pyramid, dilation, ordered
    compaction) == the oracle's restatement, per level and in the reference's raster order; tracking with the
    device-built cloud == tracking with the uploaded one."""
    tp = synth.make_tracking_pair(w=320, h=240, levels=4, n_seed_points=1500, seed=7)
    w, h, levels = tp.cam.w, tp.cam.h, tp.levels
    rng = np.random.default_rng(11)
    n = 3000
    cx = rng.uniform(3, w - 4, n).astype(np.float32); cy = rng.uniform(3, h - 4, n).astype(np.float32)
    cx[100:160] = cx[40:100]; cy[100:160] = cy[40:100]            # 60 pixels hit twice, 20 of them three times
    cx[160:180] = cx[40:60]; cy[160:180] = cy[40:60]
    idp = rng.uniform(0.1, 1.5, n).astype(np.float32)
    hdi = rng.uniform(1e-4, 1e-1, n).astype(np.float32)
    cpt = np.stack([cx, cy, idp], 1)
    su = (cx + np.float32(0.5)).astype(np.int32); sv = (cy + np.float32(0.5)).astype(np.int32)
    sw = np.sqrt((1e-3 / (hdi.astype(np.float64) + 1e-12)).astype(np.float32)).astype(np.float32)
    lv, _, dI_flat = O.pyramid(tp.img_ref, levels)
    ref = O.make_coarse_depth(levels, w, h, su, sv, idp, sw, dI_flat)
    trk = capi.CoarseTracker(ctx, w, h)
    trk.makeK(tp.cam, levels)
    fr_ref = capi.Frame(ctx, tp.img_ref, levels)
    trk.makeCoarseDepthL0(fr_ref, cpt, hdi)
    for l in range(levels):
        got = trk.get_ref(l)
        assert len(got[0]) == len(ref[l][0]) > 0, (l, len(got[0]), len(ref[l][0]))
        for a, b in zip(got, ref[l]):
            assert np.array_equal(a, b), l
    # the tracker runs on the cloud it built
    fr_new = capi.Frame(ctx, tp.img_new, levels)
    ok1, T1, aff1 = trk.trackNewestCoarse(fr_new, T_ID, (0.0, 0.0), levels - 1)
    trk2 = capi.CoarseTracker(ctx, w, h); trk2.makeK(tp.cam, levels); trk2.setCoarseTrackingRef(ref)
    ok2, T2, aff2 = trk2.trackNewestCoarse(fr_new, T_ID, (0.0, 0.0), levels - 1)
    assert ok1 == ok2 and np.array_equal(T1, T2) and np.array_equal(aff1, aff2)
    for o in (trk, trk2, fr_ref, fr_new):
        o.close()


def test_make_coarse_depth_from_ba_handle(ctx, capi, O, synth):
    """Device-resident variant: centerProjectedTo / HdiF come straight from a window handle (residuals into the newest
    keyframe, state IN); equals the host-array variant fed with the downloaded values."""
    kf = synth.render_keyframes(320, 240, 4, seed=12)
    win, _ = synth.make_window(points_per_frame=300, seed=40, keyframes=kf)
    levels = 4
    frames = [capi.Frame(ctx, im, levels) for im in win.images]
    ef = capi.EnergyFunctional(ctx); ef.set_window(win, frames); ef.set_topology(win)
    lin = ef.linearizeAll(False); ef.applyRes(); ef.accumulate_all()
    newest = win.n - 1
    ids = np.nonzero((win.target_idx == newest) & (lin["newState"] == 0))[0].astype(np.int32)
    assert len(ids) > 200
    hdi = ef.point_results()["HdiF"][win.point_idx[ids]]
    cpt = lin["centerProjectedTo"][ids]
    a = capi.CoarseTracker(ctx, 320, 240); a.makeK(win.cam, levels); a.makeCoarseDepthL0(frames[newest], cpt, hdi)
    b = capi.CoarseTracker(ctx, 320, 240); b.makeK(win.cam, levels); b.makeCoarseDepthL0_from_ba(frames[newest], ef, ids)
    for l in range(levels):
        ra, rb = a.get_ref(l), b.get_ref(l)
        assert len(ra[0]) > 0
        for x, y in zip(ra, rb):
            assert np.array_equal(x, y), l
    for o in [a, b, ef] + frames:
        o.close()


def test_cluster_and_cooperative_modes_agree(ctx, capi, O, tracking_pair):
    """The one-cluster LM kernel (cluster barrier + distributed shared memory) and the cooperative grid kernel take the
    same LM path and land on the same pose (their per-block partial sums are grouped differently: fp32 round-off)."""
    tp = tracking_pair
    trk, frame, orc, lv, dI = _setup(ctx, capi, O, tp)
    mode = trk.get_mode()
    assert mode in (8, 16), "cluster launch unavailable on this device?"
    ok1, T1, aff1 = trk.trackNewestCoarse(frame, T_ID, (0.0, 0.0), tp.levels - 1)
    it1 = trk.iterations.copy(); r1 = trk.lastResiduals.copy()
    trk.set_mode(0)
    ok2, T2, aff2 = trk.trackNewestCoarse(frame, T_ID, (0.0, 0.0), tp.levels - 1)
    assert ok1 and ok2 and np.array_equal(it1, trk.iterations)
    assert np.abs(O.se3_log(T1) - O.se3_log(T2)).max() < 1e-6 and np.abs(np.asarray(aff1) - np.asarray(aff2)).max() < 1e-5
    assert np.allclose(r1, trk.lastResiduals, rtol=1e-5)
    trk.set_mode(8)
    ok3, T3, _ = trk.trackNewestCoarse(frame, T_ID, (0.0, 0.0), tp.levels - 1)
    assert ok3 and np.abs(O.se3_log(T1) - O.se3_log(T3)).max() < 1e-6
    # deterministic within a mode
    ok4, T4, _ = trk.trackNewestCoarse(frame, T_ID, (0.0, 0.0), tp.levels - 1)
    assert np.array_equal(T3, T4)
    trk.close(); frame.close()


def test_track_batch_members_bit_identical_to_single_calls(ctx, capi, synth, tracking_pair_small):
    """hslam_tracker_track_batch: several initial poses (motion hypotheses, System.cpp:348-481) and two different frames
    in ONE launch, one cluster per member: every member's pose, affine parameters, residuals, flow indicators, iteration
    counts and return value are bit-identical to hslam_tracker_track of that member alone."""
    tp = tracking_pair_small
    trk = capi.CoarseTracker(ctx, tp.cam.w, tp.cam.h)
    trk.makeK(tp.cam, tp.levels)
    trk.setCoarseTrackingRef(tp.pc)
    if trk.get_mode() == 0:
        pytest.skip("no cluster launch on this device")
    f_new = capi.Frame(ctx, tp.img_new, tp.levels)
    f_ref = capi.Frame(ctx, tp.img_ref, tp.levels)
    rng = np.random.default_rng(4)
    Ts, affs, frames = [], [], []
    for m in range(9):
        xi = np.zeros(6) if m == 0 else rng.normal(0, [0.01, 0.01, 0.01, 0.004, 0.004, 0.004])
        R, t = synth.se3_exp(xi)
        Ts.append(synth.T12(R, t)); affs.append((0.0, 0.0) if m < 5 else (0.01 * m, 0.5 * m))
        frames.append(f_ref if m in (3, 7) else f_new)
    # a hopeless start (tracking is expected to fail or to abort): the failure must be reproduced too
    R, t = synth.se3_exp(np.array([0.5, -0.4, 0.3, 0.3, -0.2, 0.25]))
    Ts.append(synth.T12(R, t)); affs.append((0.0, 0.0)); frames.append(f_new)
    mra = np.array([np.nan, np.nan, np.nan, np.nan, np.nan])
    got = trk.trackNewestCoarse_batch(frames, Ts, affs, tp.levels - 1, mra)
    n_ok = 0
    for m in range(len(frames)):
        ok, T, aff = trk.trackNewestCoarse(frames[m], Ts[m], affs[m], tp.levels - 1, mra)
        assert ok == got["ok"][m]
        assert np.array_equal(T, got["T"][m]) and np.array_equal(aff, got["aff"][m]), m
        assert np.array_equal(trk.lastResiduals, got["lastResiduals"][m], equal_nan=True)
        assert np.array_equal(trk.lastFlowIndicators, got["lastFlow"][m])
        assert np.array_equal(trk.iterations, got["iterations"][m])
        n_ok += int(ok)
    assert n_ok >= 8
    trk.close(); f_new.close(); f_ref.close()
