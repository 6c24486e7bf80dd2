"""PINS THE ORACLE TO THE REFERENCE'S OWN CODE.  oracle/_ref/libhslam_ref.so is built from the UNMODIFIED reference sources
(/root/reference/Src/OptimizationClasses.cpp, AccumulatedTopHessian.cpp, AccumulatedSCHessian.cpp, EnergyFunctional.cpp, Settings.cpp and their
headers, incl. Include/MatrixAccumulators.h and DirectProjection.h) against the library stand-ins of oracle/ref_shim/
(recipe: oracle/Makefile, target _ref; the .so travels with the repository, the reference tree does not).  The same
window goes through that library and through the oracle's float restatement:
  * PointFrameResidual::linearize / applyRes: states, energies, all 74 Jacobian floats, JpJdF, centerProjectedTo,
    projectedTo — BIT-IDENTICAL; fixLinearizationF's res_toZeroF to fp32 round-off (Eigen dot-product order);
  * AccumulatedTopHessianSSE::addPoint<0/1> and AccumulatedSCHessianSSE::addPoint side effects on the points
    (Hdd/bd/Hcd_acc{A,L}F, HdiF, bdSumF, idepth_hessian), and stitchDouble of the three systems: the A system (no Eigen
    reduction upstream of it) to 1e-12, L and SC to fp32 round-off; resInA / resInL exact;
  * FrameFramePrecalc::set — the precalc tables the window generators produce.
Skipped only when the library has not been built (a checkout without /root/reference and without the shipped .so)."""
import importlib

import numpy as np
import pytest

from conftest import rel_fro


@pytest.fixture(scope="module")
def RL():
    from oracle import reflib
    if not reflib.available():
        pytest.skip("oracle/_ref/libhslam_ref.so not built (needs /root/reference: make -C oracle _ref)")
    return reflib


@pytest.fixture(scope="module")
def pair(O, RL, synth):
    win, is_lin = synth.make_window(points_per_frame=150, n_frames=6, seed=123, frac_linearized=0.3)
    dIs = [O.pyramid(im, 1)[0][0] for im in win.images]
    return win, is_lin, O.BAOracle(win, dIs, prec="f"), RL.RefBA(win, dIs), O.BAOracle(win, dIs, prec="d")


def _same_lin(a, b):
    assert np.array_equal(a["newState"], b["newState"])
    assert np.array_equal(a["newEnergy"], b["newEnergy"]) and np.array_equal(a["newEnergyWithOutlier"], b["newEnergyWithOutlier"])
    assert a["energy"] == b["energy"]
    ok = b["newState"] != 1
    assert np.array_equal(a["centerProjectedTo"][ok], b["centerProjectedTo"][ok])
    assert np.array_equal(a["projectedTo"][ok], b["projectedTo"][ok])


def _same_res(a, b):
    ok = (b["state_state"] != 1) & (b["is_active"] != 0)   # an OOB residual keeps a partially written Jacobian in the reference
    assert np.array_equal(a["state_state"], b["state_state"]) and np.array_equal(a["is_active"], b["is_active"])
    for k in ("JpJdF", "J"):
        assert np.array_equal(np.asarray(a[k])[ok], np.asarray(b[k])[ok]), k
    # fixLinearizationF forms Jp*delta with two 6-term and two 4-term Eigen dot products (OptimizationClasses.cpp:263-268):
    # the order Eigen reduces them in is an implementation detail of its version / SIMD flags (the stand-in splits
    # binary, the oracle sums left to right), so res_toZeroF agrees to fp32 round-off, not bit for bit
    ra, rb = np.asarray(a["res_toZeroF"])[ok], np.asarray(b["res_toZeroF"])[ok]
    assert np.abs(ra - rb).max() <= 2e-6 * max(np.abs(rb).max(), 1.0)


def test_linearize_apply_fix_bit_identical_to_reference_code(pair):
    win, is_lin, f, r, _ = pair
    lf, lr = f.linearize(False), r.linearize(False)
    assert (lr["newState"] == 0).sum() > 0.8 * len(lr["newState"]) and (lr["newState"] == 1).any() and (lr["newState"] == 2).any()
    _same_lin(lf, lr)
    f.apply_res(); r.apply_res()
    _same_res(f.residual_results(), r.residual_results())
    ids = np.nonzero(is_lin)[0].astype(np.int32)
    f.fix_linearization(ids); r.fix_linearization(ids)
    _same_res(f.residual_results(), r.residual_results())
    # second pass at a moved state: the linearised residuals stay frozen, the others are re-linearised
    idp = (win.idepth * np.float32(1.002)).astype(np.float32)
    for x in (f, r):
        x.update_points(idepth=idp, deltaF=(idp - win.idepth_zero).astype(np.float32))
    _same_lin(f.linearize(False), r.linearize(False))
    f.apply_res(); r.apply_res()
    _same_res(f.residual_results(), r.residual_results())


def test_accumulate_and_stitch_match_reference_code(pair):
    """Runs after the test above: the windows hold active A residuals, linearised L residuals and non-zero deltas."""
    win, is_lin, f, r, d = pair
    d.linearize(False); d.apply_res(); d.fix_linearization(np.nonzero(is_lin)[0].astype(np.int32))
    idp = (win.idepth * np.float32(1.002)).astype(np.float32)
    d.update_points(idepth=idp, deltaF=(idp - win.idepth_zero).astype(np.float32)); d.linearize(False); d.apply_res()
    for which in (0, 1, 2):
        Hf, bf, nf = f.accumulate(which)
        Hr, br, nr = r.accumulate(which)
        Hd, bd_, _ = d.accumulate(which)
        assert nf == nr and (which == 2 or nr > 0)
        # A: the fp32 accumulators hold the same numbers and the fp64 stitch is the same arithmetic.  L / SC sit downstream
        # of res_toZeroF (see _same_res): the linearised residuals' r carries fp32 round-off into b_L, bd_accLF, bdSumF
        tol_H, tol_b = (1e-12, 1e-12) if which == 0 else (1e-9, 1e-6)
        assert rel_fro(Hf, Hr) < tol_H and rel_fro(bf, br) < tol_b, (which, rel_fro(Hf, Hr), rel_fro(bf, br))
        assert np.abs(Hr).max() > 0
        assert rel_fro(Hr, Hd) < 1e-4          # and the reference's own float error vs the fp64 restatement is what the GPU is held to
    pf, pr = f.point_results(), r.point_results()
    for k in pf:
        a, b = np.asarray(pf[k], np.float32), np.asarray(pr[k], np.float32)
        if k in ("bd_accLF", "bdSumF"):
            assert np.allclose(a, b, rtol=1e-4, atol=1e-3 * np.abs(b).max()), k
        else:
            assert np.array_equal(a, b), k
    assert np.abs(pr["HdiF"]).max() > 0 and np.abs(pr["Hdd_accLF"]).max() > 0


def test_marginalisation_accumulate_matches_reference_code(pair):
    """Runs after the tests above.  EnergyFunctional::marginalizePointsF's part on the path — priorF scaling, addPoint<2> over
    the points to marginalise, SC addPoint(p, shiftPriorToZero=false), stitchDouble(usePrior=false) of both — compiled from the
    reference vs the oracle: M, Mb (mode 2 mixes active and linearised residuals, so they sit downstream of res_toZeroF like
    the L system) and Msc, Mbsc to fp32 round-off; resInM exact."""
    win, is_lin, f, r, d = pair
    rng = np.random.default_rng(5)
    ids = np.sort(rng.choice(win.P, size=win.P // 5, replace=False)).astype(np.int32)
    # flagPointsForRemoval fixes the linearisation of every residual of a point before it is marginalised (Mapping.cpp:286-293;
    # addPoint<2> asserts it, AccumulatedTopHessian.cpp:46)
    rids = np.concatenate([win.res_index[win.res_begin[p]:win.res_begin[p + 1]] for p in ids]).astype(np.int32)
    f.fix_linearization(rids); r.fix_linearization(rids)
    of = f.marginalize_points(ids, 0.5)
    rf = r.marginalize_points(ids, 0.5)
    assert of[4] == rf[4] and rf[4] > 0
    for a, b, tol in ((of[0], rf[0], 1e-9), (of[1], rf[1], 1e-6), (of[2], rf[2], 1e-9), (of[3], rf[3], 1e-6)):
        assert np.abs(b).max() > 0
        assert rel_fro(a, b) < tol, rel_fro(a, b)


def test_frame_frame_precalc_matches_reference_code(O, RL, synth):
    """FrameFramePrecalc::set compiled from the reference vs the table generators (numpy and the oracle's C restatement)."""
    kf = synth.render_keyframes(160, 120, 4, seed=9)
    win, _ = synth.make_window(points_per_frame=10, seed=10, keyframes=kf)
    ref = RL.precalc_set(win.cam, win.T_w2c_eval, win.state, win.state_zero, win.ab_exposure).reshape(-1, 28)
    orc = O.prepare_window(win.cam, win.T_w2c_eval, win.state, win.state_zero, win.ab_exposure)[0].reshape(-1, 28)
    npy = np.asarray(win.precalc, np.float32).reshape(-1, 28)
    n = win.n
    offdiag = np.array([h + t * n for h in range(n) for t in range(n) if h != t])
    # fp64 pose algebra (Sophus stand-in vs the restatements) rounded to fp32: equal to a few ulp
    assert np.allclose(ref[offdiag, :27], orc[offdiag, :27], rtol=3e-6, atol=3e-6)
    assert np.allclose(ref[offdiag, :27], npy[offdiag, :27], rtol=3e-5, atol=3e-5)   # numpy forms K R K^-1 in another order


# ---------------------------------------------------------------------------------------- Stage T, tracing, reference clouds
@pytest.fixture(scope="module")
def RT():
    from oracle import reflib
    if not reflib.available_trk():
        pytest.skip("oracle/_ref/libhslam_ref_trk.so not built (needs /root/reference: make -C oracle _ref)")
    return reflib


T_ID = np.concatenate([np.eye(3).reshape(9), np.zeros(3)])


def test_tracker_bit_identical_to_reference_code(O, RT, synth):
    """Src/CoarseTracker.cpp compiled unmodified: makeK, calcRes (Vec6 + the eight warped buffers), calcGSSSE with the SSE
    Accumulator9 (H, b) — BIT-IDENTICAL to the oracle's float restatement at several poses and cut-offs; trackNewestCoarse
    lands on the same pose to fp64 round-off with the same residuals."""
    tp = synth.make_tracking_pair(w=320, h=240, levels=4, n_seed_points=1500, seed=7)
    lv, _, dI = O.pyramid(tp.img_new, tp.levels)
    orc = O.TrackerOracle(tp.cam, tp.levels, tp.pc)
    ref = RT.RefTracker(tp.cam, tp.levels, tp.pc, dI)
    for l in range(tp.levels):
        ko, kr = orc.level_K(l), ref.level_K(l)
        assert all(np.float32(ko[k]) == np.float32(kr[k]) for k in ("w", "h", "fx", "fy", "cx", "cy")) and np.array_equal(ko["Ki"], kr["Ki"])
    R, t = synth.se3_exp(np.array([0.01, -0.004, 0.006, 0.003, -0.002, 0.004]))
    for T, aff, cutoff in ((T_ID, (0.0, 0.0), 20.0), (synth.T12(R, t), (0.02, 3.0), 20.0), (synth.T12(R, t), (0.0, 0.0), 5.0)):
        for l in range(tp.levels):
            o6, wo, nwo, _ = orc.calc_res(l, lv[l], T, aff, cutoff)
            r6, wr, nwr = ref.calc_res(l, T, aff, cutoff)
            assert np.array_equal(o6, r6) and nwr == (nwo + 3) // 4 * 4 and nwo > 50
            assert all(np.array_equal(a[:nwo], b[:nwo]) for a, b in zip(wo, wr))
            Ho, bo = orc.calc_gs(l, wo, nwo, aff)
            Hr, br = ref.calc_gs(l, T, aff)
            assert np.array_equal(Ho, Hr) and np.array_equal(bo, br)
    ro = orc.track(dI, T_ID, (0.0, 0.0), tp.levels - 1, prec="f")
    rr = ref.track(T_ID, (0.0, 0.0), tp.levels - 1)
    assert ro["ok"] and rr["ok"]
    assert np.abs(ro["T"] - rr["T"]).max() < 1e-12 and np.abs(ro["aff"] - rr["aff"]).max() < 1e-10   # 8x8 LDLT / SE3 exp round-off
    assert np.allclose(ro["lastResiduals"], rr["lastResiduals"], rtol=1e-12, equal_nan=True)
    assert np.allclose(ro["flow"], rr["flow"], rtol=1e-9)


def test_make_coarse_depth_bit_identical_to_reference_code(O, RT, synth):
    """CoarseTracker::makeCoarseDepthL0 compiled unmodified (splat incl. pixel collisions, pyramid, dilation, compaction)."""
    tp = synth.make_tracking_pair(w=320, h=240, levels=4, n_seed_points=1500, seed=7)
    w, h, levels = tp.cam.w, tp.cam.h, tp.levels
    rng = np.random.default_rng(11)
    n = 3000
    cx = rng.uniform(3, w - 4, n).astype(np.float32); cy = rng.uniform(3, h - 4, n).astype(np.float32)
    cx[100:160] = cx[40:100]; cy[100:160] = cy[40:100]
    idp = rng.uniform(0.1, 1.5, n).astype(np.float32); hdi = rng.uniform(1e-4, 1e-1, n).astype(np.float32)
    su = (cx + np.float32(0.5)).astype(np.int32); sv = (cy + np.float32(0.5)).astype(np.int32)
    sw = np.sqrt((1e-3 / (hdi.astype(np.float64) + 1e-12)).astype(np.float32)).astype(np.float32)
    _, _, dI_ref = O.pyramid(tp.img_ref, levels)
    orc = O.make_coarse_depth(levels, w, h, su, sv, idp, sw, dI_ref)
    ref = RT.RefTracker(tp.cam, levels, tp.pc, dI_ref).make_coarse_depth(np.stack([cx, cy, idp], 1), hdi, dI_ref)
    for l in range(levels):
        assert len(ref[l][0]) == len(orc[l][0]) > 0
        for a, b in zip(orc[l], ref[l]):
            assert np.array_equal(a, b), l


def test_trace_on_bit_identical_to_reference_code(O, RT, synth):
    """ImmaturePoint::traceOn (Src/ImmaturePoint.cpp, unmodified) over 1200 points, two consecutive frames."""
    import copy
    kf = synth.render_keyframes(320, 240, 5, seed=78)
    ims = synth.make_immature_points(kf, hosts=[0, 1, 2], points_per_host=400, seed=5)
    a, b = copy.deepcopy(ims), copy.deepcopy(ims)
    for new in (3, 4):
        dI0 = O.pyramid(kf.images[new], 1)[0][0]
        for x in (a, b):
            synth.trace_tables(kf, [0, 1, 2], new, x)
        O.trace_immature(dI0, a); RT.trace_immature(dI0, kf.cam, b)
        for k in ("status", "idepth_min", "idepth_max", "quality", "lastTraceUV", "lastTracePixelInterval"):
            assert np.array_equal(getattr(a, k), getattr(b, k), equal_nan=True), (new, k)
    assert (b.status == 0).sum() > 800 and (b.status != 0).any()


def test_point_activation_bit_identical_to_reference_code(O, RT, synth):
    """System::optimizeImmaturePoint (Src/FullSystemOptPoint.cpp) with ImmaturePoint::linearizeResidual (+ projectPoint /
    derive_idepth of DirectProjection.h), both compiled unmodified, vs the oracle's restatement: activation flag, optimised inverse depth and the final
    state of every residual, bit for bit — for candidates with tight intervals and for hopeless ones."""
    import copy
    kf = synth.render_keyframes(320, 240, 5, seed=79)
    ims = synth.make_immature_points(kf, hosts=[0, 1, 2], points_per_host=500, seed=6)
    dIs = [O.pyramid(im, 1)[0][0] for im in kf.images]
    for new in (3, 4):
        synth.trace_tables(kf, [0, 1, 2], new, ims)
        O.trace_immature(dIs[new], ims)
    cand = np.nonzero(np.isfinite(ims.idepth_max))[0]          # good and not-so-good traces alike
    assert len(cand) > 800
    sub = copy.deepcopy(ims)
    for k in ("host_idx", "u", "v", "gradH", "color", "weights", "energyTH", "idepth_min", "idepth_max"):
        setattr(sub, k, np.ascontiguousarray(getattr(ims, k)[cand]))
    T = synth.activation_tables(kf)
    for min_obs in (1, 3):
        o = O.activate_immature(dIs, kf.cam, T, sub, minObs=min_obs)
        r = RT.activate_immature(dIs, kf.cam, T, sub, minObs=min_obs)
        assert np.array_equal(o[0], r[0]) and np.array_equal(o[1], r[1]) and np.array_equal(o[2], r[2])
        assert 0.3 < r[0].mean() < 1.0 or min_obs == 3
    assert r[0].sum() > 100


def test_frame_pyramid_bit_identical_to_reference_code(O, synth):
    """Frame::CreateDirPyrs (Src/Frame.cpp:104-181, compiled unmodified and run by the reference's own Frame constructor) vs
    the oracle's pyramid, which the GPU pyramid kernel equals bit for bit: intensity of every level everywhere; dx, dy and
    absSquaredGrad on rows 1 .. h-2 (the reference never writes the first and the last row)."""
    from oracle import reflib
    if not reflib.available_frame():
        pytest.skip("oracle/_ref/libhslam_ref_frame.so not built (needs /root/reference: make -C oracle _ref)")
    kf = synth.render_keyframes(320, 240, 1, seed=77)
    rng = np.random.default_rng(3)
    for img, (w, h) in ((kf.images[0], (320, 240)), ((rng.random((96, 160)) * 255).astype(np.float32), (160, 96))):
        levels = 5 if w == 320 else 4
        lo, ao, _ = O.pyramid(img, levels)
        lr, ar = reflib.create_dir_pyrs(img, levels)
        for l in range(levels):
            hl = h >> l
            assert np.array_equal(lo[l][..., 0], lr[l][..., 0]), l
            assert np.array_equal(lo[l][1:hl - 1], lr[l][1:hl - 1]), l
            assert np.array_equal(ao[l][1:hl - 1], ar[l][1:hl - 1]), l
            assert np.abs(lr[l][1:hl - 1, :, 1]).max() > 0


def test_adjoints_and_deltas_match_reference_code(O, RL, synth):
    """EnergyFunctional::setAdjointsF + setDeltaF (Src/EnergyFunctional.cpp:22-82, 128-152, compiled unmodified) vs the table
    generator the oracle / the host mirror use: adTarget bit-identical, adHost to one fp64 ulp (the order inside SE3::Adj), the
    float adHTdeltaF to float round-off."""
    kf = synth.render_keyframes(160, 120, 5, seed=9)
    win, _ = synth.make_window(points_per_frame=10, seed=10, keyframes=kf)
    _, adH, adT, adHTd = O.prepare_window(win.cam, win.T_w2c_eval, win.state, win.state_zero, win.ab_exposure)
    rH, rT, rD = RL.adjoints_deltas(win.T_w2c_eval, win.state, win.state_zero, win.ab_exposure)
    assert np.array_equal(adT, rT)
    assert np.abs(adH - rH).max() <= 4e-16 * max(1.0, np.abs(rH).max()) and np.abs(rH).max() > 0.5
    assert np.abs(adHTd - rD).max() <= 1e-6 * np.abs(rD).max() and np.abs(rD).max() > 1e-3


def test_resubstitute_and_lenergy_match_reference_code(O, RL, synth):
    """EnergyFunctional::resubstituteF_MT / resubstituteFPt and calcLEnergyF_MT / calcLEnergyPt — the reference's own member
    functions (with its own IndexThreadReduce, one worker) after linearize / applyRes / fixLinearizationF / accumulate A, L, SC —
    vs the oracle: point steps to float round-off (the 8-term dot products xAd * JpJdF associate as the build's Eigen does),
    the linearised energy to 1e-8."""
    kf = synth.render_keyframes(320, 240, 5, seed=12)
    win, is_lin = synth.make_window(points_per_frame=150, seed=40, keyframes=kf, frac_linearized=0.3)
    dIs = [O.pyramid(im, 1)[0][0] for im in win.images]
    f, r = O.BAOracle(win, dIs, prec="f"), RL.RefBA(win, dIs)
    ids = np.nonzero(is_lin)[0].astype(np.int32)
    idp = (win.idepth * np.float32(1.002)).astype(np.float32)
    for x in (f, r):
        x.linearize(False); x.apply_res(); x.fix_linearization(ids)
        x.update_points(idepth=idp, deltaF=(idp - win.idepth_zero).astype(np.float32))
        x.linearize(False); x.apply_res()
        for k in (0, 1, 2):
            x.accumulate(k)
    xs = np.random.default_rng(1).normal(0, 1e-3, 8 * win.n + 4)
    sf, sr = f.resubstitute(xs), r.resubstitute(xs)
    assert np.abs(sr).max() > 1e-3 and (sr == 0).sum() < 0.2 * len(sr)
    assert np.abs(sf - sr).max() <= 1e-6 * np.abs(sr).max()
    assert (sf == sr).mean() > 0.25                      # (half of the points even bit for bit)
    ef, er = f.calc_l_energy(), r.calc_l_energy()
    assert er > 0 and abs(ef - er) <= 1e-8 * abs(er)


def test_system_linearize_all_and_new_frame_energy_th_bit_identical(O, RL, synth):
    """System::linearizeAll(false) — linearizeAll_Reductor over activeResiduals and System::setNewFrameEnergyTH
    (Src/FullSystemOptimize.cpp:19-165, compiled unmodified; the bare System object is set up by the driver) — over three
    GN-style iterations with the threshold fed back through Frame::frameEnergyTH, some residuals fixed after the first: the
    summed energy and the new threshold (std::nth_element at 70 %, the float arithmetic of :86-91) equal the oracle's BIT FOR
    BIT, and so do the residual states that the thresholds decide."""
    kf = synth.render_keyframes(320, 240, 5, seed=12)
    win, is_lin = synth.make_window(points_per_frame=150, seed=40, keyframes=kf, frac_linearized=0.3)
    dIs = [O.pyramid(im, 1)[0][0] for im in win.images]
    f, r = O.BAOracle(win, dIs, prec="f"), RL.RefBA(win, dIs)
    ids = np.nonzero(is_lin)[0].astype(np.int32)
    ths = []
    for it in range(3):
        lf = f.linearize(False)
        thf = f.new_frame_energy_th(win.n - 1)
        er, thr = r.system_linearize_all()
        assert lf["energy"] == er and er > 0
        assert np.float32(thf) == thr
        ths.append(float(thr))
        f.apply_res(); r.apply_res()
        if it == 0:
            f.fix_linearization(ids); r.fix_linearization(ids)
        idp = (win.idepth * np.float32(1.0 + 0.002 * (it + 1))).astype(np.float32)
        for x in (f, r):
            x.update_points(idepth=idp, deltaF=(idp - win.idepth_zero).astype(np.float32))
    assert len(set(ths)) == 3 and all(t != win.frameEnergyTH[win.n - 1] for t in ths)
    assert np.array_equal(f.residual_results()["state_state"], r.residual_results()["state_state"])

