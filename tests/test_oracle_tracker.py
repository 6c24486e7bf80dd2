"""Pins the tracker half of the CPU oracle with known answers (the reference ships no tests:
PARITY UNPINNED by the reference) and checks its two precisions against each other."""
import numpy as np
import pytest

from conftest import rel_fro

T_ID = np.concatenate([np.eye(3).reshape(9), np.zeros(3)])


# ---------------------------------------------------------------- Sophus identities (Thirdparty/Sophus/sophus/test_se3.cpp)
def test_se3_exp_log_roundtrip(O):
    rng = np.random.default_rng(0)
    for _ in range(50):
        xi = rng.normal(0, 0.5, 6)
        T = O.se3_exp(xi)
        R = T[:9].reshape(3, 3)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-13) and np.linalg.det(R) > 0
        assert np.allclose(O.se3_log(T), xi, atol=1e-10)
    assert np.allclose(O.se3_exp(np.zeros(6)), T_ID)
    assert np.allclose(O.se3_log(O.se3_exp([1e-13, 0, 0, 1e-12, 0, 0])), [1e-13, 0, 0, 1e-12, 0, 0], atol=1e-15)


def test_se3_group_laws_and_adjoint(O, synth):
    rng = np.random.default_rng(1)
    for _ in range(20):
        A, B = O.se3_exp(rng.normal(0, 0.3, 6)), O.se3_exp(rng.normal(0, 0.3, 6))
        assert np.allclose(O.se3_mul(A, O.se3_inv(A)), T_ID, atol=1e-13)
        assert np.allclose(O.se3_mul(O.se3_mul(A, B), O.se3_inv(B)), A, atol=1e-12)
        # Ad_A x = vee(A hat(x) A^-1)  <=>  A exp(x) A^-1 = exp(Ad_A x)
        x = rng.normal(0, 0.1, 6)
        lhs = O.se3_mul(O.se3_mul(A, O.se3_exp(x)), O.se3_inv(A))
        rhs = O.se3_exp(O.se3_adj(A) @ x)
        assert np.allclose(lhs, rhs, atol=1e-12)
        # numpy generator and oracle agree
        R, t = synth.se3_exp(x)
        assert np.allclose(synth.T12(R, t), O.se3_exp(x), atol=1e-14)
        assert np.allclose(synth.se3_adj(A), O.se3_adj(A), atol=1e-14)


def test_ldlt_solve(O):
    rng = np.random.default_rng(2)
    for n in (6, 7, 8):
        M = rng.normal(size=(n, n))
        A = M @ M.T + 1e-3 * np.eye(n)
        b = rng.normal(size=n)
        assert np.allclose(O.ldlt_solve(A, b), np.linalg.solve(A, b), rtol=1e-9, atol=1e-12)


# ---------------------------------------------------------------- pyramid
def test_pyramid_matches_numpy_restatement(O, synth):
    rng = np.random.default_rng(3)
    img = (rng.random((96, 128)) * 255).astype(np.float32)
    lv, ag, _ = O.pyramid(img, 4)
    pn, an = synth.build_pyramid_np(img, 4)
    for l in range(4):
        assert np.array_equal(lv[l], pn[l]) and np.array_equal(ag[l], an[l])
    # known answers: a horizontal ramp has dx = slope, dy = 0 away from the wrap columns
    ramp = np.tile(np.arange(64, dtype=np.float32) * 2.0, (48, 1))
    lv, ag, _ = O.pyramid(ramp, 2)
    assert np.all(lv[0][1:-1, 1:-1, 1] == 2.0) and np.all(lv[0][1:-1, 1:-1, 2] == 0.0)
    assert np.all(lv[1][1:-1, 1:-1, 1] == 4.0)          # half resolution doubles the per-pixel slope
    assert np.all(lv[0][0, :, 1:] == 0) and np.all(lv[0][-1, :, 1:] == 0)   # rows the reference never writes
    # row-end wrap of the flat-index central difference (Src/Frame.cpp:145-146)
    assert lv[0][5, 0, 1] == 0.5 * (ramp[5, 1] - ramp[4, 63])
    gamma = np.linspace(0.5, 1.5, 256).astype(np.float32)
    _, agg, _ = O.pyramid(ramp, 1, gamma=gamma)
    c = np.clip((ramp + 0.5).astype(np.int32), 5, 250)
    assert np.allclose(agg[0][1:-1, 1:-1], (4.0 * gamma[c] ** 2)[1:-1, 1:-1])


def test_coarse_depth_numpy_vs_oracle(O, synth, tracking_pair_small):
    tp = tracking_pair_small
    rng = np.random.default_rng(7)
    lv, ag, dI = O.pyramid(tp.img_ref, tp.levels)
    h, w = tp.img_ref.shape
    n = 800
    su = rng.integers(3, w - 3, n).astype(np.int32); sv = rng.integers(3, h - 3, n).astype(np.int32)
    sid = rng.uniform(0.2, 1.0, n).astype(np.float32); sw = rng.uniform(0.5, 2.0, n).astype(np.float32)
    a = O.make_coarse_depth(tp.levels, w, h, su, sv, sid, sw, dI)
    # duplicates make np.add.at order-dependent only in the last ulp; remove them for an exact check
    _, first = np.unique(sv.astype(np.int64) * w + su, return_index=True)
    first.sort()
    a = O.make_coarse_depth(tp.levels, w, h, su[first], sv[first], sid[first], sw[first], dI)
    b = synth.make_coarse_depth_np(tp.levels, w, h, su[first], sv[first], sid[first], sw[first], lv)
    for l in range(tp.levels):
        for x, y in zip(a[l], b[l]):
            assert np.array_equal(x, y), l


# ---------------------------------------------------------------- calcRes / calcGS
@pytest.fixture(scope="module")
def trk(O, tracking_pair_small):
    tp = tracking_pair_small
    orc = O.TrackerOracle(tp.cam, tp.levels, tp.pc)
    lv, _, dI = O.pyramid(tp.img_new, tp.levels)
    lvr, _, dIr = O.pyramid(tp.img_ref, tp.levels)
    return tp, orc, lv, dI, lvr, dIr


def test_zero_motion_known_answer(trk):
    """Tracking the reference image against itself at identity: every residual is 0, so E = 0 and b = 0."""
    tp, orc, lv, dI, lvr, dIr = trk
    for l in range(tp.levels):
        out6, warped, nw, mask = orc.calc_res(l, lvr[l], T_ID, (0.0, 0.0), 20.0, "d")
        # (K * Ki in fp32 is the identity only to ~1e-5 px, hence "tiny", not exactly 0)
        assert out6[0] < 1e-6 * max(nw, 1) and out6[5] == 0 and nw == out6[1] > 0
        H, b = orc.calc_gs(l, warped, nw, (0.0, 0.0), "d")
        assert np.abs(b).max() < 1e-4 * np.sqrt(np.abs(np.diag(H)).max())
        ev = np.linalg.eigvalsh(H)
        assert np.allclose(H, H.T) and ev.min() > -1e-9 * ev.max()


def test_float_and_double_restatements_agree(trk):
    tp, orc, lv, dI, lvr, dIr = trk
    for l in range(tp.levels):
        f6, wf, nf, mf = orc.calc_res(l, lv[l], T_ID, (0.0, 0.0), 20.0, "f")
        d6, wd, nd, md = orc.calc_res(l, lv[l], T_ID, (0.0, 0.0), 20.0, "d")
        assert (mf != md).mean() < 2e-3           # only threshold-boundary points may flip
        if np.array_equal(mf, md):
            assert f6[0] == pytest.approx(d6[0], rel=2e-4)   # sequential fp32 sum of ~1e3-1e4 terms
            Hf, bf = orc.calc_gs(l, wf, nf, (0.0, 0.0), "f")
            Hd, bd = orc.calc_gs(l, wd, nd, (0.0, 0.0), "d")
            assert rel_fro(Hf, Hd) < 1e-5 and rel_fro(bf, bd) < 1e-4


def test_b_is_half_energy_gradient(O, trk, synth):
    """Finite differences: dE/dxi_i = 2 * n_pad * b_i / scale_i (Huber energy, left-multiplied increments).
    The Jacobian uses interpolated gradient images, so the match is to a few percent, enough to catch any
    sign / ordering / scaling error (incl. the SCALE_XI_ROT/TRANS naming quirk, CoarseTracker.cpp:312-323)."""
    tp, orc0, lv, dI, lvr, dIr = trk
    l = 1
    # keep only points that stay well inside the image, so that no point crosses the bounds gate
    pc = []
    for k, (u, v, idp, col) in enumerate(tp.pc):
        wl, hl = tp.cam.w >> k, tp.cam.h >> k
        m = (u > 12) & (u < wl - 13) & (v > 12) & (v < hl - 13)
        pc.append((u[m], v[m], idp[m], col[m]))
    orc = O.TrackerOracle(tp.cam, tp.levels, pc)
    R, t = synth.se3_exp([0.004, -0.002, 0.005, 0.0015, -0.002, 0.001])
    T = synth.T12(R, t)
    aff = (0.01, 1.5)
    out6, warped, nw, mask = orc.calc_res(l, lv[l], T, aff, 1e9, "d")   # huge cutoff: no saturation, smooth energy
    H, b = orc.calc_gs(l, warped, nw, aff, "d")
    npad = (nw + 3) & ~3
    scales = np.array([1, 1, 1, 0.5, 0.5, 0.5, 10, 1000.0])
    g = np.zeros(8)
    for i in range(8):
        eps = 1e-5 if i < 6 else (1e-5 if i == 6 else 1e-3)
        Es = []
        for s in (+1, -1):
            if i < 6:
                xi = np.zeros(6); xi[i] = s * eps
                Tn = O.se3_mul(O.se3_exp(xi), T); an = aff
            else:
                Tn = T; an = (aff[0] + (s * eps if i == 6 else 0), aff[1] + (s * eps if i == 7 else 0))
            e6, _, n2, m2 = orc.calc_res(l, lv[l], Tn, an, 1e9, "d", want_warped=False)
            assert np.array_equal(m2, mask)       # same inlier set, otherwise the difference quotient is meaningless
            Es.append(e6[0])
        g[i] = (Es[0] - Es[1]) / (2 * eps)
    pred = 2 * npad * b / scales
    # affine columns are exact derivatives
    assert np.allclose(g[6:], pred[6:], rtol=1e-4), (g, pred)
    # pose columns: DSO differentiates with central-difference gradient images, which under-estimate
    # the slope of the bilinearly interpolated texture by a sinc factor (~0.6-0.85 for this texture), so
    # compare direction (sign/order of all six columns) and a bounded magnitude ratio
    cos = g[:6] @ pred[:6] / (np.linalg.norm(g[:6]) * np.linalg.norm(pred[:6]))
    assert cos > 0.99, (cos, g, pred)
    ratio = g[:6] / pred[:6]
    assert np.all(ratio > 0.9) and np.all(ratio < 1.9), ratio


def test_track_converges_to_rendered_motion(O, trk):
    tp, orc, lv, dI, lvr, dIr = trk
    for prec in "fd":
        r = orc.track(dI, T_ID, (0.0, 0.0), tp.levels - 1, prec=prec)
        assert r["ok"]
        assert np.allclose(O.se3_log(r["T"]), O.se3_log(tp.T_ref2new_true), atol=5e-4)
        assert abs(r["aff"][0] - tp.aff_new_true[0]) < 0.05   # a and b are only weakly separable
        assert np.all(r["lastResiduals"][: tp.levels] < 12)
        assert r["iterations"][: tp.levels].min() >= 1
    # abort path: a tiny minResForAbort stops after the coarsest level and leaves the pose untouched
    r = orc.track(dI, T_ID, (0.0, 0.0), tp.levels - 1, min_res=np.full(5, 1e-3), prec="f")
    assert not r["ok"] and np.array_equal(r["T"], T_ID) and np.isnan(r["lastResiduals"][0])


def test_accumulator9_tiers_do_not_change_the_sum(O, tracking_pair):
    """More than 1000 4-point updates exercise shiftUp (MatrixAccumulators.h:1325-1344)."""
    tp = tracking_pair
    orc = O.TrackerOracle(tp.cam, tp.levels, tp.pc)
    lv, _, _ = O.pyramid(tp.img_new, tp.levels)
    f6, wf, nf, mf = orc.calc_res(0, lv[0], T_ID, (0.0, 0.0), 20.0, "f")
    assert nf > 4 * 1000
    Hf, bf = orc.calc_gs(0, wf, nf, (0.0, 0.0), "f")
    Hd, bd = orc.calc_gs(0, [w.astype(np.float64) for w in wf], nf, (0.0, 0.0), "d")
    assert rel_fro(Hf, Hd) < 2e-6 and rel_fro(bf, bd) < 2e-5
