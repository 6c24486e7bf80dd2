"""Pins the BA half of the CPU oracle: host tables vs an independent numpy restatement, stitched
systems vs an explicit dense normal-equation build + dense Schur elimination, float vs double."""
import numpy as np
import pytest

from conftest import rel_fro


@pytest.fixture(scope="module")
def setup(O, synth):
    win, is_lin = synth.make_window(points_per_frame=120, n_frames=5, seed=99, frac_linearized=0.25)
    dIs = [O.pyramid(im, 1)[0][0] for im in win.images]
    return win, is_lin, dIs


def test_host_tables_numpy_vs_oracle(O, setup):
    win, _, _ = setup
    pc, adH, adT, adHTd = O.prepare_window(win.cam, win.T_w2c_eval, win.state, win.state_zero, win.ab_exposure)
    assert np.allclose(pc, win.precalc, rtol=2e-6, atol=1e-4)     # fp32 products in two orders
    assert np.allclose(adH, win.adHost, rtol=1e-12, atol=1e-13)
    assert np.allclose(adT, win.adTarget, rtol=1e-12, atol=1e-13)
    assert np.allclose(adHTd, win.adHTdeltaF, rtol=1e-4, atol=1e-8)
    n = win.n
    # structure of the adjoints (Src/EnergyFunctional.cpp:38-58): AT is diagonal, AH(6,6)=AH(7,7)/100*10...
    for k in range(n * n):
        AT = adT[64 * k:64 * k + 64].reshape(8, 8).T
        assert np.count_nonzero(AT - np.diag(np.diag(AT))) == 0
        assert np.allclose(np.diag(AT)[:6], [0.5, 0.5, 0.5, 1, 1, 1])
        assert AT[7, 7] == -1000.0


def _dense_systems(win, res, pts_prior, pts_delta, mode):
    """Explicit J^T J over every pattern pixel in ABSOLUTE coordinates, then dense Schur elimination.
    Independent of addPoint's factored accumulators and of the relative->absolute stitch."""
    n = win.n
    N = 8 * n + 4
    P = win.P
    J, act = res["J"], res["is_active"].astype(bool)
    adH = win.adHost.reshape(n * n, 8, 8).transpose(0, 2, 1)   # [k][r][c]
    adT = win.adTarget.reshape(n * n, 8, 8).transpose(0, 2, 1)
    Hpp = np.zeros((N, N)); bp = np.zeros(N)
    Hpd = np.zeros((N, P)); Hdd = np.zeros(P); bd = np.zeros(P)
    for r in range(win.R):
        if not act[r] or (res["lin"][r] != (mode == 1)):
            continue
        p = win.point_idx[r]; h = win.host_idx[p]; t = win.target_idx[r]; k = h + t * n
        Jr = J[r]
        for i in range(8):
            jx, jy = Jr[30 + i], Jr[38 + i]
            rel = np.concatenate([jx * Jr[8:14] + jy * Jr[14:20], [Jr[46 + i], Jr[54 + i]]])   # d r_i / d (xi_rel, a, b)
            row = np.zeros(N)
            row[:4] = jx * Jr[20:24] + jy * Jr[24:28]
            row[4 + 8 * h:12 + 8 * h] += adH[k] @ rel
            row[4 + 8 * t:12 + 8 * t] += adT[k] @ rel
            jd = jx * Jr[28] + jy * Jr[29]
            ri = res["resid"][r, i]
            Hpp += np.outer(row, row); bp += row * ri
            Hpd[:, p] += row * jd; Hdd[p] += jd * jd; bd[p] += jd * ri
    return Hpp, bp, Hpd, Hdd, bd


def test_stitched_systems_equal_dense_elimination(O, setup):
    win, is_lin, dIs = setup
    ba = O.BAOracle(win, dIs, prec="d")
    ba.linearize(False); ba.apply_res()
    HA, bA, nA = ba.accumulate(0)
    HL, bL, nL = ba.accumulate(1)
    HS, bS, _ = ba.accumulate(2)
    rr = ba.residual_results()
    res = dict(J=rr["J"], is_active=rr["is_active"], lin=np.zeros(win.R, bool), resid=rr["J"][:, :8])
    Hpp, bp, Hpd, Hdd, bd = _dense_systems(win, res, win.priorF, win.deltaF, 0)
    assert nA == int(rr["is_active"].sum()) and nL == 0
    assert rel_fro(HA, Hpp) < 1e-11 and rel_fro(bA, bp) < 1e-10
    # H_L with no linearised residual = priors only (AccumulatedTopHessian.cpp:269-279)
    D = np.zeros_like(HL)
    np.fill_diagonal(D, np.concatenate([win.cPrior, win.framePrior]))
    assert np.allclose(HL, D)
    assert np.allclose(bL[:4], win.cPrior * win.cDeltaF) and np.allclose(bL[4:], win.framePrior * win.frameDeltaPrior)
    # Schur complement of the inverse depths (AccumulatedSCHessian.cpp:23-33 adds priorF to Hdd and
    # priorF*deltaF to bd); points without an active residual drop out
    has = np.zeros(win.P, bool)
    np.logical_or.at(has, win.point_idx[rr["is_active"].astype(bool)], True)
    Hd = np.maximum(Hdd + win.priorF, 1e-10)
    bdS = bd + win.priorF * win.deltaF
    Hsc = (Hpd[:, has] / Hd[has]) @ Hpd[:, has].T
    bsc = Hpd[:, has] @ (bdS[has] / Hd[has])
    assert rel_fro(HS, Hsc) < 1e-10 and rel_fro(bS, bsc) < 1e-9
    # the reduced system is PSD up to the gauge directions
    ev = np.linalg.eigvalsh(HA - HS)
    assert ev.min() > -1e-8 * ev.max()
    pr = ba.point_results()
    assert np.allclose(pr["idepth_hessian"][has], Hd[has], rtol=1e-12)
    assert np.all(pr["HdiF"][~has] == 0)


def test_linearized_mode_uses_frozen_jacobians(O, setup):
    """mode 1: resApprox = res_toZeroF + J*delta must reproduce resF when delta is unchanged since fixLinearizationF."""
    win, is_lin, dIs = setup
    ba = O.BAOracle(win, dIs, prec="d")
    ba.linearize(False); ba.apply_res()
    HA0, bA0, nA0 = ba.accumulate(0)
    ids = np.nonzero(is_lin)[0].astype(np.int32)
    ba.fix_linearization(ids)
    HA, bA, nA = ba.accumulate(0)
    HL, bL, nL = ba.accumulate(1)
    rr = ba.residual_results()
    assert nL == int((rr["is_active"].astype(bool) & is_lin.astype(bool)).sum()) and nA + nL == nA0
    # A + L (minus the priors that only L carries) = the un-split A
    D = np.zeros_like(HL)
    np.fill_diagonal(D, np.concatenate([win.cPrior, win.framePrior]))
    bprior = np.concatenate([win.cPrior * win.cDeltaF, win.framePrior * win.frameDeltaPrior])
    assert rel_fro(HA + HL - D, HA0) < 1e-9
    assert rel_fro(bA + bL - bprior, bA0) < 1e-6


def test_float_vs_double(O, setup):
    win, is_lin, dIs = setup
    out = {}
    for prec in "fd":
        ba = O.BAOracle(win, dIs, prec=prec)
        lin = ba.linearize(False); ba.apply_res()
        ba.fix_linearization(np.nonzero(is_lin)[0].astype(np.int32))
        out[prec] = (lin, ba.accumulate(0), ba.accumulate(1), ba.accumulate(2), ba.residual_results())
    lf, ld = out["f"][0], out["d"][0]
    assert (lf["newState"] != ld["newState"]).mean() < 2e-3
    if np.array_equal(lf["newState"], ld["newState"]):
        for k in (1, 2, 3):
            assert rel_fro(out["f"][k][0], out["d"][k][0]) < 2e-5
            assert rel_fro(out["f"][k][1], out["d"][k][1]) < 1e-4
    ok = ld["newState"] != 1
    assert np.allclose(out["f"][4]["J"][ok], out["d"][4]["J"][ok], rtol=2e-2, atol=2e-2)


def test_threads_match_single_thread(O, setup):
    """IndexThreadReduce semantics: per-thread accumulators + stitch give the single-thread result up to fp32 re-association."""
    win, is_lin, dIs = setup
    ref = None
    for threads in (1, 4):
        ba = O.BAOracle(win, dIs, prec="f", threads=threads)
        lin = ba.linearize(False); ba.apply_res()
        r = (lin["newState"].copy(), lin["energy"], ba.accumulate(0), ba.accumulate(2), ba.point_results())
        if ref is None:
            ref = r
        else:
            assert np.array_equal(ref[0], r[0])
            assert ref[1] == pytest.approx(r[1], rel=1e-12)
            assert rel_fro(r[2][0], ref[2][0]) < 1e-5 and r[2][2] == ref[2][2]
            assert rel_fro(r[3][0], ref[3][0]) < 1e-5
            for k in ref[4]:
                assert np.array_equal(ref[4][k], r[4][k])     # per-point sums do not depend on the thread split


def test_oob_and_outlier_states(O, synth, setup):
    win, is_lin, dIs = setup
    import copy
    w2 = copy.copy(win)
    w2.u = win.u.copy(); w2.u[:10] = -50.0                            # host pixel far outside the image => OOB
    w2.frameEnergyTH = np.full(win.n, 1e-3, np.float32)               # everything else becomes an outlier
    ba = O.BAOracle(w2, dIs, prec="f")
    lin = ba.linearize(False)
    ba.apply_res()
    first = np.isin(win.point_idx, np.arange(10))
    assert np.all(lin["newState"][first] == 1)
    assert np.all(lin["newState"][~first] != 0)
    assert np.all(lin["newEnergy"][lin["newState"] == 2] == np.float32(1e-3))     # clamped to the threshold
    assert np.all(lin["newEnergyWithOutlier"][lin["newState"] == 1] == -1)
    H, b, nres = ba.accumulate(0)
    assert nres == 0 and not H.any()
    # once OOB, always OOB (OptimizationClasses.cpp:47-51,239-242)
    lin2 = ba.linearize(False)
    assert np.all(lin2["newState"][first] == 1)


def test_marginalisation_pass_reduces_to_active_accumulate_at_zero_delta(O, synth):
    """Known answer for the marginalisation restatement (EnergyFunctional.cpp:545-609): with every delta zero
    (state == state_zero, deltaF = 0, cDeltaF = 0) res_toZeroF == resF, so addPoint<2> over ALL points after fixing all
    active residuals equals addPoint<0> before (M == H_A without priors, Mb == b_A), and the Schur part equals
    accumulateSCF's (the prior shift is zero).  Also: calcLEnergy is zero at zero delta."""
    win, _ = synth.make_window(w=320, h=240, points_per_frame=80, n_frames=4, seed=12, state_noise=0.0)
    win.deltaF[:] = 0
    win.cDeltaF[:] = 0
    win.frameDeltaPrior[:] = 0
    assert np.abs(win.adHTdeltaF).max() == 0
    dIs = [O.pyramid(im, 1)[0][0] for im in win.images]
    o = O.BAOracle(win, dIs, prec="d")
    o.linearize(False); o.apply_res()
    HA, bA, nA = o.accumulate(0)
    o.accumulate(1)
    HS, bS, _ = o.accumulate(2)
    act = o.residual_results()["is_active"].astype(bool)
    o.fix_linearization(np.nonzero(act)[0].astype(np.int32))
    assert abs(o.calc_l_energy()) < 1e-9
    M, Mb, Ms, Mbs, nres = o.marginalize_points(np.arange(win.P, dtype=np.int32), 1.0)
    assert nres == nA == act.sum()
    assert rel_fro(M, HA) < 1e-12 and rel_fro(Mb, bA) < 1e-12
    assert rel_fro(Ms, HS) < 1e-12 and rel_fro(Mbs, bS) < 1e-12


def test_new_frame_energy_th_is_the_order_statistic(O, setup):
    """System::setNewFrameEnergyTH (FullSystemOptimize.cpp:60-101): the oracle's nth_element value equals the sorted
    array's entry at (int)(0.7f * count); the threshold arithmetic; the empty case 12*12*patternNum; and the value is
    stored as the window's frameEnergyTH (the next linearisation classifies with it)."""
    win, is_lin, dIs = setup
    ba = O.BAOracle(win, dIs, prec="f")
    lin = ba.linearize(False)
    for newest in range(win.n):
        sel = (win.target_idx == newest) & (win.is_linearized == 0) & (lin["newEnergyWithOutlier"] >= 0)
        vals = np.sort(lin["newEnergyWithOutlier"][sel])
        th = ba.new_frame_energy_th(newest)
        if len(vals) == 0:
            assert th == np.float32(12 * 12 * 8)
            continue
        k = int(np.float32(0.7) * np.float32(len(vals)))
        t = np.sqrt(vals[k], dtype=np.float32) * np.float32(1.5)
        t = np.float32(26.0) * np.float32(0.5) + t * np.float32(0.5)
        t = np.float32(t * t)
        assert th == t and 50.0 < th < 1e5
    # thresholds are live: a tiny one turns inliers into outliers in the next pass
    n_in = int((lin["newState"] == 0).sum())
    lin2 = ba.linearize(False)
    assert int((lin2["newState"] == 2).sum()) >= int((lin["newState"] == 2).sum())
    assert n_in > 0
    # empty case: a window whose residuals into frame 0 are all linearised has no candidate
    import copy
    w2 = copy.copy(win)
    w2.is_linearized = np.where(win.target_idx == 0, 1, win.is_linearized).astype(np.uint8)
    ba2 = O.BAOracle(w2, dIs, prec="f")
    ba2.linearize(False)
    assert ba2.new_frame_energy_th(0) == np.float32(1152.0)
