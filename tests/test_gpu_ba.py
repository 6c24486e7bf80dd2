"""GPU parity: windowed BA linearisation + accumulate A/L/SC + stitch through the C ABI vs the oracle.

Bit-exact: residual states, energies, the 74 Jacobian floats, JpJdF, res_toZeroF and the per-point
Hdd/bd/Hcd/HdiF/bdSumF (vs the float-faithful restatement).  H,b of the three systems: <= 1e-5
relative Frobenius vs the fp64 oracle (BASELINE.json north_star) and no worse than the float
restatement's own error.
"""
import numpy as np
import pytest

from conftest import rel_fro

pytestmark = pytest.mark.gpu


def run_sequence(ba, win, is_lin, kind):
    """The same call sequence on the GPU mirror (kind='gpu') and on the oracle (kind='orc')."""
    lin = ba.linearizeAll(False) if kind == "gpu" else ba.linearize(False)
    (ba.applyRes if kind == "gpu" else ba.apply_res)()
    ids = np.nonzero(is_lin)[0].astype(np.int32)
    (ba.fixLinearization if kind == "gpu" else ba.fix_linearization)(ids)
    idp = (win.idepth * np.float32(1.001)).astype(np.float32)
    ba.update_points(idepth=idp, deltaF=(idp - win.idepth_zero).astype(np.float32))
    lin2 = ba.linearizeAll(False) if kind == "gpu" else ba.linearize(False)
    (ba.applyRes if kind == "gpu" else ba.apply_res)()
    return lin, lin2


@pytest.fixture(scope="module")
def ba_setup(ctx, capi, O, synth, window_small):
    win, is_lin = window_small
    frames = [capi.Frame(ctx, im, 1) for im in win.images]
    dIs = [O.pyramid(im, 1)[0][0] for im in win.images]
    ef = capi.EnergyFunctional(ctx)
    ef.set_window(win, frames)
    ef.set_topology(win)
    of = O.BAOracle(win, dIs, prec="f")
    od = O.BAOracle(win, dIs, prec="d")
    g = run_sequence(ef, win, is_lin, "gpu")
    f = run_sequence(of, win, is_lin, "orc")
    d = run_sequence(od, win, is_lin, "orc")
    yield dict(win=win, is_lin=is_lin, ef=ef, of=of, od=od, g=g, f=f, d=d, frames=frames)
    ef.close()
    for fr in frames:
        fr.close()


def test_linearize_states_and_energies_bit_exact(ba_setup):
    for i in (0, 1):
        g, f = ba_setup["g"][i], ba_setup["f"][i]
        assert np.array_equal(g["newState"], f["newState"])
        assert np.bincount(g["newState"], minlength=3)[0] > 0.8 * len(g["newState"]) * (1 if i == 0 else 0.6)
        assert np.array_equal(g["newEnergy"], f["newEnergy"])
        assert np.array_equal(g["newEnergyWithOutlier"], f["newEnergyWithOutlier"])
        ok = g["newState"] != 1
        assert np.array_equal(g["centerProjectedTo"][ok], f["centerProjectedTo"][ok])
        assert np.array_equal(g["projectedTo"][ok], f["projectedTo"][ok])
        assert g["energy"] == pytest.approx(f["energy"], rel=1e-12)


def test_jacobians_bit_exact(ba_setup):
    rg = ba_setup["ef"].residual_results()
    rf = ba_setup["of"].residual_results()
    assert np.array_equal(rg["state_state"], rf["state_state"])
    assert np.array_equal(rg["is_active"], rf["is_active"])
    ok = rg["state_state"] != 1
    assert np.array_equal(rg["J"][ok], rf["J"][ok].astype(np.float32))
    assert np.array_equal(rg["JpJdF"][ok], rf["JpJdF"][ok].astype(np.float32))
    # fixLinearizationF is only meaningful for residuals whose Jacobian was completed (the reference calls
    # it on active residuals only); an OOB residual keeps a partially written J in the reference
    lin = ba_setup["is_lin"].astype(bool) & (ba_setup["g"][0]["newState"] != 1)
    assert lin.sum() > 100
    assert np.array_equal(rg["res_toZeroF"][lin], rf["res_toZeroF"][lin].astype(np.float32))


def test_accumulate_against_double_oracle(ba_setup):
    ef, of, od = ba_setup["ef"], ba_setup["of"], ba_setup["od"]
    HA, bA = ef.accumulateAF_MT()
    HL, bL = ef.accumulateLF_MT()
    HS, bS = ef.accumulateSCF_MT()
    ref = {}
    for name, o in (("f", of), ("d", od)):
        a = o.accumulate(0); l = o.accumulate(1); s = o.accumulate(2)
        ref[name] = (a, l, s)
    assert ef.resInA == ref["f"][0][2] and ef.resInL == ref["f"][1][2]
    assert ef.resInA > 0 and ef.resInL > 0
    for (Hg, bg), k, nm in (((HA, bA), 0, "A"), ((HL, bL), 1, "L"), ((HS, bS), 2, "SC")):
        Hd, bd = ref["d"][k][0], ref["d"][k][1]
        Hf, bf = ref["f"][k][0], ref["f"][k][1]
        eH, eb = rel_fro(Hg, Hd), rel_fro(bg, bd)
        fH, fb = rel_fro(Hf, Hd), rel_fro(bf, bd)
        assert eH <= max(1e-5, 1.5 * fH), (nm, eH, fH)
        assert eb <= max(1e-5, 1.5 * fb), (nm, eb, fb)
    # structure: A and L are symmetrised by the reference's stitch, SC only in the calibration columns
    assert np.allclose(HA, HA.T, rtol=0, atol=1e-9 * np.abs(HA).max())
    assert np.allclose(HS[:4, 4:], HS[4:, :4].T)
    # fused pass gives the same three systems bit for bit
    allr = ef.accumulate_all()
    for (H1, b1), (H2, b2) in zip(allr, ((HA, bA), (HL, bL), (HS, bS))):
        assert np.array_equal(H1, H2) and np.array_equal(b1, b2)


@pytest.mark.parametrize("n_frames", [2, 3, 9, 11])
def test_schur_kernel_variants_by_window_size(ctx, capi, O, synth, n_frames):
    """The Schur fold picks its kernel by the number of keyframes: 8x8 register tiles with eight point slices (n <= 7), with
    four slices (n <= 10), 4x4 tiles beyond.  Every variant: per-point results bit-exact vs the float restatement, the three
    systems <= 1e-5 vs the fp64 oracle (and no worse than 1.5 x the float restatement)."""
    kf = synth.render_keyframes(320, 240, n_frames, seed=90 + n_frames)
    win, is_lin = synth.make_window(points_per_frame=90, seed=300 + n_frames, keyframes=kf, frac_linearized=0.25)
    frames = [capi.Frame(ctx, im, 1) for im in win.images]
    dIs = [O.pyramid(im, 1)[0][0] for im in win.images]
    ef = capi.EnergyFunctional(ctx); ef.set_window(win, frames); ef.set_topology(win)
    of, od = O.BAOracle(win, dIs, prec="f"), O.BAOracle(win, dIs, prec="d")
    run_sequence(ef, win, is_lin, "gpu"); run_sequence(of, win, is_lin, "orc"); run_sequence(od, win, is_lin, "orc")
    sys3 = ef.accumulate_all()
    for k, (Hg, bg) in enumerate(sys3):
        Hd, bd, _ = od.accumulate(k)
        Hf, bf, _ = of.accumulate(k)
        assert rel_fro(Hg, Hd) <= max(1e-5, 1.5 * rel_fro(Hf, Hd)), (n_frames, k)
        assert rel_fro(bg, bd) <= max(1e-5, 1.5 * rel_fro(bf, bd)), (n_frames, k)
    pg, pf = ef.point_results(), of.point_results()
    for key in ("HdiF", "bdSumF"):
        assert np.array_equal(pg[key], pf[key].astype(np.float32)), (n_frames, key)
    ef.close()
    for f in frames:
        f.close()


def test_point_side_effects_bit_exact(ba_setup):
    ef, of = ba_setup["ef"], ba_setup["of"]
    ef.accumulateAF_MT(); ef.accumulateLF_MT(); ef.accumulateSCF_MT()
    of.accumulate(0); of.accumulate(1); of.accumulate(2)
    pg, pf = ef.point_results(), of.point_results()
    for k in pg:
        assert np.array_equal(pg[k], pf[k].astype(np.float32)), k
    assert (pg["HdiF"] > 0).sum() > 0.9 * len(pg["HdiF"])


def test_raw_schur_blocks_vs_double(ba_setup):
    """Schur identity on the stitched system: H_sc equals the dense elimination of the inverse depths."""
    od = ba_setup["od"]
    od.accumulate(0); od.accumulate(1); od.accumulate(2)
    raw = od.raw_sc()
    n = ba_setup["win"].n
    # accD[(h,t1,t2)] is symmetric under t1<->t2 transpose
    for h in range(n):
        for t1 in range(n):
            for t2 in range(n):
                a = raw["D"][h + n * t1 + n * n * t2]
                b = raw["D"][h + n * t2 + n * n * t1]
                assert np.allclose(a, b.T, rtol=1e-12, atol=1e-9)


def test_deterministic(ba_setup):
    ef = ba_setup["ef"]
    a = ef.accumulate_all()
    for _ in range(3):
        b = ef.accumulate_all()
        for (H1, b1), (H2, b2) in zip(a, b):
            assert np.array_equal(H1, H2) and np.array_equal(b1, b2)


def test_resubstitute(ba_setup):
    ef, of, od = ba_setup["ef"], ba_setup["of"], ba_setup["od"]
    ef.accumulate_all()
    for o in (of, od):
        o.accumulate(0); o.accumulate(1); o.accumulate(2)
    rng = np.random.default_rng(5)
    x = rng.normal(0, 1e-3, ef.N)
    sg, sf, sd = ef.resubstituteF_MT(x), of.resubstitute(x), od.resubstitute(x)
    assert np.array_equal(sg, sf)
    assert rel_fro(sg, sd) < 1e-4


def test_iteration_device_matches_stepwise(ctx, capi, O, synth, window_small):
    win, is_lin = window_small
    frames = [capi.Frame(ctx, im, 1) for im in win.images]
    a = capi.EnergyFunctional(ctx); a.set_window(win, frames); a.set_topology(win)
    b = capi.EnergyFunctional(ctx); b.set_window(win, frames); b.set_topology(win)
    a.linearizeAll(False, outputs=False); a.applyRes(); ra = a.accumulate_all()
    b.iteration_device(); rb = b.download_Hb()
    for (H1, b1), (H2, b2) in zip(ra, rb):
        assert np.array_equal(H1, H2) and np.array_equal(b1, b2)
    a.close(); b.close()
    for f in frames:
        f.close()


def test_bad_arguments(ctx, capi, synth, window_small):
    win, _ = window_small
    ef = capi.EnergyFunctional(ctx)
    with pytest.raises(capi.HslamError):
        ef.set_topology(win)            # window first
    frames = [capi.Frame(ctx, im, 1) for im in win.images]
    ef.set_window(win, frames)
    import copy
    bad = copy.copy(win)
    bad.target_idx = win.target_idx.copy(); bad.target_idx[0] = 99
    with pytest.raises(capi.HslamError):
        ef.set_topology(bad)
    ef.close()
    for f in frames:
        f.close()


def test_empty_window(ctx, capi, synth, window_small):
    """No points / no residuals: H_A = 0, H_L = priors only, H_SC = 0 (reference: loops do nothing)."""
    import copy
    win, _ = window_small
    e = copy.copy(win)
    z = lambda dt: np.zeros(0, dt)
    e.u = e.v = e.idepth = e.idepth_zero = e.priorF = e.deltaF = z(np.float32)
    e.color = e.weights = np.zeros((0, 8), np.float32)
    e.host_idx = z(np.int32); e.res_begin = np.zeros(1, np.int32); e.res_index = z(np.int32)
    e.point_idx = e.target_idx = z(np.int32); e.state_state = z(np.int8); e.is_linearized = e.is_active = z(np.uint8)
    frames = [capi.Frame(ctx, im, 1) for im in win.images]
    ef = capi.EnergyFunctional(ctx); ef.set_window(e, frames); ef.set_topology(e)
    ef.linearizeAll(False, outputs=False); ef.applyRes()
    (HA, bA), (HL, bL), (HS, bS) = ef.accumulate_all()
    assert not HA.any() and not bA.any() and not HS.any() and not bS.any()
    assert np.allclose(np.diag(HL)[:4], win.cPrior)
    assert np.allclose(np.diag(HL)[4:], win.framePrior)
    assert np.allclose(bL[4:], win.framePrior * win.frameDeltaPrior)
    ef.close()
    for f in frames:
        f.close()


def test_batch_matches_single_windows(ctx, capi, O, synth):
    """hslam_ba_batch_iteration (4 launches for all windows) == per-window iteration, bit for bit."""
    kf = synth.render_keyframes(320, 240, 5, seed=77)
    wins = [synth.make_window(points_per_frame=100 + 30 * i, seed=300 + i, keyframes=kf)[0] for i in range(3)]
    frames = [capi.Frame(ctx, im, 1) for im in kf.images]
    singles, batched = [], []
    for w in wins:
        a = capi.EnergyFunctional(ctx); a.set_window(w, frames); a.set_topology(w); singles.append(a)
        b = capi.EnergyFunctional(ctx); b.set_window(w, frames); b.set_topology(w); batched.append(b)
    batch = capi.EnergyFunctionalBatch(ctx, batched)
    for a, b in zip(singles, batched):   # same chunking => same grouping of the fp32 partial sums
        assert b.get_parallelism() > 0
        a.set_parallelism(b.get_parallelism())
    for _ in range(2):   # two iterations: tickets / partial buffers are reusable
        batch.iteration()
        for a in singles:
            a.iteration_device()
    for a, b in zip(singles, batched):
        ra, rb = a.download_Hb(), b.download_Hb()
        for (H1, b1), (H2, b2) in zip(ra, rb):
            assert np.array_equal(H1, H2) and np.array_equal(b1, b2)
            assert np.abs(H1).max() > 0 or True
        pa, pb = a.point_results(), b.point_results()
        for k in pa:
            assert np.array_equal(pa[k], pb[k]), k
    batch.close()
    for e in singles + batched:
        e.close()
    for f in frames:
        f.close()


def test_async_calls_match_sync(ctx, capi, synth, window_small):
    win, _ = window_small
    frames = [capi.Frame(ctx, im, 1) for im in win.images]
    a = capi.EnergyFunctional(ctx); a.set_window(win, frames); a.set_topology(win)
    b = capi.EnergyFunctional(ctx); b.set_window(win, frames); b.set_topology(win)
    la = a.linearizeAll(False, outputs="states"); a.applyRes(); ra = a.accumulate_all()
    b.begin_async()
    lb = b.linearizeAll(False, outputs="states"); b.applyRes(); rb = b.accumulate_all()
    with pytest.raises(capi.HslamError):
        lb.get()                      # not before wait()
    b.wait()
    lb, rb = lb.get(), rb.get()
    assert np.array_equal(la["newState"], lb["newState"]) and la["energy"] == lb["energy"]
    for (H1, b1), (H2, b2) in zip(ra, rb):
        assert np.array_equal(H1, H2) and np.array_equal(b1, b2)
    a.close(); b.close()
    for f in frames:
        f.close()


@pytest.mark.parametrize("pinned", ["bound", True, False])
def test_batch_step_host_matches_single_window_calls(ctx, capi, synth, pinned):
    """hslam_ba_batch_step_host / _resubstitute_host (one staging block up, one down) == the per-window
    host calls (set_window, update_points, linearize, applyRes, accumulate_all, resubstitute), bit for bit,
    with page-locked (direct DMA) and with pageable (staged) caller buffers."""
    kf = synth.render_keyframes(320, 240, 5, seed=78)
    wins = [synth.make_window(points_per_frame=90 + 40 * i, seed=400 + i, keyframes=kf, frac_linearized=0.2 * i)[0] for i in range(3)]
    frames = [capi.Frame(ctx, im, 1) for im in kf.images]
    singles, batched = [], []
    for w in wins:
        a = capi.EnergyFunctional(ctx); a.set_window(w, frames); a.set_topology(w); singles.append(a)
        b = capi.EnergyFunctional(ctx); b.set_window(w, frames); b.set_topology(w); batched.append(b)
    batch = capi.EnergyFunctionalBatch(ctx, batched)
    for a, b in zip(singles, batched):
        a.set_parallelism(b.get_parallelism())
    io = batch.make_io(wins, [frames] * len(wins), pinned=pinned)
    rng = np.random.default_rng(9)
    for it in range(2):
        # new point state every iteration, through the io block's input arrays
        for w, (ws, ka, inp) in zip(wins, io["keep"]):
            inp["idepth"][:] = (w.idepth * np.float32(1.0 + 0.001 * it)).astype(np.float32)
            inp["deltaF"][:] = inp["idepth"] - w.idepth_zero
        outs = batch.step_host(io)
        xs = [rng.normal(0, 1e-3, 8 * w.n + 4) for w in wins]
        steps = batch.resubstitute_host(io, xs)
        for a, w, o, (ws, ka, inp), x, st in zip(singles, wins, outs, io["keep"], xs, steps):
            a.set_window(w, frames)
            a.update_points(idepth=inp["idepth"], deltaF=inp["deltaF"])
            lin = a.linearizeAll(False, outputs="states"); a.applyRes()
            sys3 = a.accumulate_all()
            assert lin["energy"] == o["energy"][0]
            assert np.array_equal(lin["newState"], o["newState"])
            assert np.array_equal(lin["newEnergy"], o["newEnergy"])
            assert np.array_equal(lin["newEnergyWithOutlier"], o["newEnergyWithOutlier"])
            assert (a.resInA, a.resInL) == (int(o["nres2"][0]), int(o["nres2"][1]))
            N = a.N; per = N * N + N
            for s3, (H, b) in enumerate(sys3):
                blk = o["Hb"][s3 * per:(s3 + 1) * per]
                assert np.array_equal(H, blk[:N * N].reshape(N, N).T) and np.array_equal(b, blk[N * N:])
            assert np.array_equal(a.resubstituteF_MT(x), st)
            assert np.abs(st).max() > 0
    batch.close()
    for e in singles + batched:
        e.close()
    for f in frames:
        f.close()


def test_batch_step_host_window_groups(ctx, capi, synth, monkeypatch):
    """hslam_ba_batch_step_host pipelines a batch of >= 8 windows in window groups (linearise group g+1 while the
    results of group g travel): every output is bit-identical to the un-grouped call (HSLAM_STEP_GROUPS=1) and to an
    odd group count that does not divide the batch."""
    kf = synth.render_keyframes(320, 240, 4, seed=79)
    wins = [synth.make_window(points_per_frame=60 + 11 * i, seed=500 + i, keyframes=kf)[0] for i in range(9)]
    frames = [capi.Frame(ctx, im, 1) for im in kf.images]
    results = []
    for groups in (None, "1", "5"):
        if groups is None:
            monkeypatch.delenv("HSLAM_STEP_GROUPS", raising=False)
        else:
            monkeypatch.setenv("HSLAM_STEP_GROUPS", groups)
        efs = []
        for w in wins:
            b = capi.EnergyFunctional(ctx); b.set_window(w, frames); b.set_topology(w); efs.append(b)
        batch = capi.EnergyFunctionalBatch(ctx, efs)
        io = batch.make_io(wins, [frames] * len(wins), pinned="bound")
        outs = batch.step_host(io)
        rng = np.random.default_rng(12)
        steps = batch.resubstitute_host(io, [rng.normal(0, 1e-3, 8 * w.n + 4) for w in wins])
        assert all(np.abs(st).max() > 0 for st in steps)
        results.append([{k: np.array(v, copy=True) for k, v in o.items()} for o in outs])
        batch.close()
        for e in efs:
            e.close()
    for other in results[1:]:
        for o0, o1 in zip(results[0], other):
            for k in o0:
                assert np.array_equal(o0[k], o1[k]), k
    assert all(np.abs(o["Hb"]).max() > 0 and (o["newState"] == 0).any() for o in results[0])
    for f in frames:
        f.close()


def test_marginalize_points_and_lenergy(ctx, capi, O, synth):
    """EnergyFunctional::marginalizePointsF (addPoint<2> + SC addPoint(p,false) + stitchDouble without priors) and
    calcLEnergyF_MT against the oracle: M, Mb, Msc, Mbsc <= 1e-5 vs fp64 and no worse than the float restatement."""
    win, _ = synth.make_window(w=320, h=240, n_frames=5, points_per_frame=300, seed=31)
    frames = [capi.Frame(ctx, im, 1) for im in win.images]
    dIs = [O.pyramid(im, 1)[0][0] for im in win.images]
    ef = capi.EnergyFunctional(ctx); ef.set_window(win, frames); ef.set_topology(win)
    of = O.BAOracle(win, dIs, prec="f"); od = O.BAOracle(win, dIs, prec="d")
    for b in (ef, of, od):
        (b.linearizeAll if b is ef else b.linearize)(False)
        (b.applyRes if b is ef else b.apply_res)()
    rng = np.random.default_rng(4)
    pts = np.sort(rng.choice(win.P, size=win.P // 5, replace=False)).astype(np.int32)
    act = ef.residual_results(want_J=False)["is_active"]
    ids = np.array([r for p in pts for r in win.res_index[win.res_begin[p]:win.res_begin[p + 1]] if act[r]], np.int32)
    assert len(ids) > 100
    # move the linearisation point away so that res_toZeroF differs from resF, then fix those residuals
    idp = (win.idepth * np.float32(1.002)).astype(np.float32)
    for b in (ef, of, od):
        b.update_points(idepth=idp, deltaF=(idp - win.idepth_zero).astype(np.float32))
        (b.fixLinearization if b is ef else b.fix_linearization)(ids)
    eg, ef_, ed = ef.calcLEnergyF_MT(), of.calc_l_energy(), od.calc_l_energy()
    assert ed != 0 and abs(eg - ed) <= max(1e-5, 1.5 * abs(ef_ - ed) / abs(ed)) * abs(ed), (eg, ef_, ed)
    scale = 36.0
    g = ef.marginalizePointsF(pts, scale)
    f = of.marginalize_points(pts, scale)
    d = od.marginalize_points(pts, scale)
    assert g[4] == f[4] == d[4] and g[4] > 0
    for k, nm in enumerate(("M", "Mb", "Msc", "Mbsc")):
        e, fe = rel_fro(g[k], d[k]), rel_fro(f[k], d[k])
        assert np.abs(d[k]).max() > 0
        assert e <= max(1e-5, 1.5 * fe), (nm, e, fe)
    # the pass leaves the handle usable: a normal accumulate afterwards equals the oracle's
    sys3 = ef.accumulate_all()
    for k, (H, b) in enumerate(sys3):
        Hd, bd, _ = od.accumulate(k)
        assert rel_fro(H, Hd) <= 1e-5 and rel_fro(b, bd) <= 1e-4, k
    ef.close()
    for fr in frames:
        fr.close()


def test_new_frame_energy_th_bit_exact(ctx, capi, O, synth, window_small):
    """hslam_ba_new_frame_energy_th == System::setNewFrameEnergyTH of the oracle (std::nth_element), bit for bit, for
    every frame of the window as `newest`, incl. the empty case (12*12*patternNum); the stored threshold drives the
    next linearisation exactly like the oracle's."""
    import copy
    win, _ = window_small
    frames = [capi.Frame(ctx, im, 1) for im in win.images]
    dIs = [O.pyramid(im, 1)[0][0] for im in win.images]
    ef = capi.EnergyFunctional(ctx); ef.set_window(win, frames); ef.set_topology(win)
    of = O.BAOracle(win, dIs, prec="f")
    ef.linearizeAll(False, outputs=False); of.linearize(False)
    ths = []
    for newest in range(win.n):
        tg, to = ef.setNewFrameEnergyTH(newest), of.new_frame_energy_th(newest)
        assert np.float32(tg) == to, (newest, tg, to)
        ths.append(tg)
    assert len(set(ths)) > 1 and all(50 < t < 1e5 for t in ths)
    # both sides now classify with the new thresholds
    g, f = ef.linearizeAll(False, outputs="states"), of.linearize(False)
    assert np.array_equal(g["newState"], f["newState"]) and np.array_equal(g["newEnergy"], f["newEnergy"])
    assert g["energy"] == pytest.approx(f["energy"], rel=1e-12)
    ef.close()
    # empty case: every residual into frame 0 is linearised -> no candidate -> 1152
    w2 = copy.copy(win)
    w2.is_linearized = np.where(win.target_idx == 0, 1, win.is_linearized).astype(np.uint8)
    ef2 = capi.EnergyFunctional(ctx); ef2.set_window(w2, frames); ef2.set_topology(w2)
    ef2.linearizeAll(False, outputs=False)
    assert ef2.setNewFrameEnergyTH(0) == 1152.0
    ef2.close()
    # and a window without residuals at all
    e = copy.copy(win)
    z = lambda dt: np.zeros(0, dt)
    e.u = e.v = e.idepth = e.idepth_zero = e.priorF = e.deltaF = z(np.float32)
    e.color = e.weights = np.zeros((0, 8), np.float32)
    e.host_idx = z(np.int32); e.res_begin = np.zeros(1, np.int32); e.res_index = z(np.int32)
    e.point_idx = e.target_idx = z(np.int32); e.state_state = z(np.int8); e.is_linearized = e.is_active = z(np.uint8)
    ef3 = capi.EnergyFunctional(ctx); ef3.set_window(e, frames); ef3.set_topology(e)
    ef3.linearizeAll(False, outputs=False)
    assert ef3.setNewFrameEnergyTH(win.n - 1) == 1152.0
    ef3.close()
    for fr in frames:
        fr.close()


def test_batch_step_host_slim_outputs_with_device_threshold(ctx, capi, O, synth):
    """The slim result set of hslam_ba_batch_step_host (no per-residual energies on the host; the new frameEnergyTH of
    the newest frame from the device select): over three GN iterations the thresholds, states, energies and systems
    equal the oracle's linearizeAll -> setNewFrameEnergyTH -> accumulate sequence; the caller hands the threshold back
    in its hslam_window like the reference's Frame::frameEnergyTH."""
    kf = synth.render_keyframes(320, 240, 5, seed=81)
    wins = [synth.make_window(points_per_frame=100 + 30 * i, seed=600 + i, keyframes=kf, frac_linearized=0.15 * i)[0] for i in range(9)]
    frames = [capi.Frame(ctx, im, 1) for im in kf.images]
    dIs = [O.pyramid(im, 1)[0][0] for im in kf.images]
    efs = []
    for w in wins:
        b = capi.EnergyFunctional(ctx); b.set_window(w, frames); b.set_topology(w); efs.append(b)
    batch = capi.EnergyFunctionalBatch(ctx, efs)
    io = batch.make_io(wins, [frames] * len(wins), pinned="bound")
    for o in io["arr"]:
        o.newEnergy = None; o.newEnergyWithOutlier = None
    orcs = [O.BAOracle(w, dIs, prec="f") for w in wins]
    launches0 = None
    for it in range(3):
        for w, (ws, ka, inp) in zip(wins, io["keep"]):
            inp["idepth"][:] = (w.idepth * np.float32(1.0 + 0.002 * it)).astype(np.float32)
            inp["deltaF"][:] = inp["idepth"] - w.idepth_zero
        outs = batch.step_host(io)
        for w, o, orc, (ws, ka, inp) in zip(wins, outs, orcs, io["keep"]):
            orc.update_points(idepth=inp["idepth"], deltaF=inp["deltaF"])
            lo = orc.linearize(False)
            th = orc.new_frame_energy_th(w.n - 1)
            orc.apply_res()
            assert np.float32(o["frameEnergyTH"][0]) == th
            assert np.array_equal(o["newState"], lo["newState"])
            assert o["energy"][0] == pytest.approx(lo["energy"], rel=1e-12)
            # the reference keeps the threshold in the Frame; the caller's window table carries it into the next call
            ka[1]["frameEnergyTH"][w.n - 1] = o["frameEnergyTH"][0]
    # systems of the last iteration against the oracle (fp32 accumulation order differs: tolerance, not bits)
    for w, o, orc in zip(wins, outs, orcs):
        N = 8 * w.n + 4; per = N * N + N
        for s3 in range(3):
            Ho, bo, _ = orc.accumulate(s3)
            blk = o["Hb"][s3 * per:(s3 + 1) * per]
            assert rel_fro(blk[:N * N].reshape(N, N).T, Ho) < 2e-5
    batch.close()
    for e in efs:
        e.close()
    for f in frames:
        f.close()


def test_batch_step_host_split_accumulate_is_bit_identical(ctx, capi, synth):
    """hslam_ba_batch_step_host on >= 8 same-shape windows with bound outputs takes the SPLIT accumulate path (A / L systems,
    energy and counts stitched and copied under the Schur fold, the Schur-complement system behind it): every output is
    bit-identical to the plain path (HSLAM_NO_SPLIT); so is a run that lets the device form deltaF (derive_deltaF)."""
    import os
    kf = synth.render_keyframes(320, 240, 5, seed=85)
    wins = [synth.make_window(points_per_frame=150, seed=800 + i, keyframes=kf, frac_linearized=0.2)[0] for i in range(8)]
    assert len({w.R for w in wins}) == 1
    frames = [capi.Frame(ctx, im, 1) for im in kf.images]
    res = {}
    for mode in ("split", "plain", "derive"):
        if mode == "plain":
            os.environ["HSLAM_NO_SPLIT"] = "1"
        try:
            efs = []
            for w in wins:
                b = capi.EnergyFunctional(ctx); b.set_window(w, frames); b.set_topology(w); efs.append(b)
            batch = capi.EnergyFunctionalBatch(ctx, efs)
            io = batch.make_io(wins, [frames] * len(wins), pinned="bound")
            for o in io["arr"]:
                o.newEnergy = None; o.newEnergyWithOutlier = None      # the slim result set: what the split form is for
                if mode == "derive":                                   # deltaF = idepth - idepth_zero formed on the device
                    o.deltaF = None; o.derive_deltaF = 1
            outs = None
            for it in range(2):
                for w, (ws, ka, inp) in zip(wins, io["keep"]):
                    inp["idepth"][:] = (w.idepth * np.float32(1.0 + 0.003 * (it + 1))).astype(np.float32)
                    inp["deltaF"][:] = inp["idepth"] - w.idepth_zero
                outs = batch.step_host(io)
            res[mode] = [{k: np.array(o[k]).copy() for k in ("Hb", "energy", "nres2", "newState", "frameEnergyTH")} for o in outs]
            batch.close()
            for e in efs:
                e.close()
        finally:
            os.environ.pop("HSLAM_NO_SPLIT", None)
    for a, b, c in zip(res["split"], res["plain"], res["derive"]):
        for k in a:
            assert np.array_equal(a[k], b[k]), k
            assert np.array_equal(a[k], c[k]), ("derive_deltaF", k)
        assert np.isfinite(a["Hb"]).all() and np.abs(a["Hb"]).max() > 0
    for f in frames:
        f.close()


def test_batch_iteration_with_device_threshold(ctx, capi, O, synth):
    """Device-resident GN loop: hslam_ba_batch_set_newest_frame makes every hslam_ba_batch_iteration run
    setNewFrameEnergyTH on the device; the NEXT linearisation classifies with the new threshold, like the oracle's
    linearizeAll -> setNewFrameEnergyTH -> applyRes -> linearizeAll sequence."""
    kf = synth.render_keyframes(320, 240, 5, seed=84)
    wins = [synth.make_window(points_per_frame=120 + 40 * i, seed=650 + i, keyframes=kf, frac_linearized=0.1 * i)[0] for i in range(5)]
    frames = [capi.Frame(ctx, im, 1) for im in kf.images]
    dIs = [O.pyramid(im, 1)[0][0] for im in kf.images]
    efs = []
    for w in wins:
        b = capi.EnergyFunctional(ctx); b.set_window(w, frames); b.set_topology(w); efs.append(b)
    batch = capi.EnergyFunctionalBatch(ctx, efs)
    batch.set_newest_frame(wins[0].n - 1)
    orcs = [O.BAOracle(w, dIs, prec="f") for w in wins]
    changed = 0
    for it in range(3):
        batch.iteration()
        for w, ef, orc in zip(wins, efs, orcs):
            lo = orc.linearize(False)
            th = orc.new_frame_energy_th(w.n - 1)
            orc.apply_res()
            rg = ef.residual_results(want_J=False)
            assert np.array_equal(rg["state_state"], orc.residual_results()["state_state"]), it
            e, _, _ = ef.download_stats()
            assert e == pytest.approx(lo["energy"], rel=1e-12)
            changed += int(th != np.float32(w.frameEnergyTH[w.n - 1]))   # (the oracle keeps the new threshold itself)
    assert changed > 0
    batch.close()
    for e in efs:
        e.close()
    for f in frames:
        f.close()


def test_calc_l_energy_after_step_host_with_new_window(ctx, capi, O, synth):
    """calcLEnergy's frame-prior term must follow the window tables passed through hslam_ba_batch_step_host."""
    import copy
    kf = synth.render_keyframes(320, 240, 4, seed=83)
    win, is_lin = synth.make_window(points_per_frame=80, seed=700, keyframes=kf, frac_linearized=0.3)
    frames = [capi.Frame(ctx, im, 1) for im in kf.images]
    dIs = [O.pyramid(im, 1)[0][0] for im in kf.images]
    ef = capi.EnergyFunctional(ctx); ef.set_window(win, frames); ef.set_topology(win)
    batch = capi.EnergyFunctionalBatch(ctx, [ef])
    w2 = copy.copy(win)
    w2.frameDeltaPrior = (np.asarray(win.frameDeltaPrior) + 0.37).copy()
    w2.framePrior = (np.asarray(win.framePrior) * 1.5 + 1.0).copy()
    io = batch.make_io([w2], [frames], pinned=False)
    batch.step_host(io)
    ids = np.nonzero(is_lin)[0].astype(np.int32)
    ef.fixLinearization(ids)
    orc = O.BAOracle(w2, dIs, prec="d")
    orc.linearize(False); orc.apply_res(); orc.fix_linearization(ids)
    assert ef.calcLEnergyF_MT() == pytest.approx(orc.calc_l_energy(), rel=1e-5)
    orc0 = O.BAOracle(win, dIs, prec="d")
    orc0.linearize(False); orc0.apply_res(); orc0.fix_linearization(ids)
    assert abs(orc0.calc_l_energy() - orc.calc_l_energy()) > 1e-3 * abs(orc.calc_l_energy())   # the prior term matters here
    batch.close(); ef.close()
    for f in frames:
        f.close()


def test_gn_loop_jacobian_mode_matches_full_mode(ctx, capi, O, synth, window_small):
    """hslam_ba_set_jacobian_mode(1): the lineariser does not store resF / JIdx / JabF / projectedTo for the residuals it
    re-linearises every iteration; fix_linearization re-forms them.  States, energies, JpJdF, res_toZeroF, every
    Jacobian float of the linearised residuals, the three systems, L energy and the back-substitution are bit-identical
    to full mode."""
    win, is_lin = window_small
    frames = [capi.Frame(ctx, im, 1) for im in win.images]
    res = []
    for mode in (0, 1):
        ef = capi.EnergyFunctional(ctx); ef.set_window(win, frames); ef.set_topology(win)
        ef.set_jacobian_mode(mode)
        lin, lin2 = run_sequence(ef, win, is_lin, "gpu")
        sys3 = ef.accumulate_all()
        rr = ef.residual_results()
        x = np.random.default_rng(5).normal(0, 1e-3, ef.N)
        res.append(dict(lin=lin, lin2=lin2, sys3=sys3, rr=rr, E=ef.calcLEnergyF_MT(), step=ef.resubstituteF_MT(x), pr=ef.point_results()))
        ef.close()
    a, b = res
    for k in ("newState", "newEnergy", "newEnergyWithOutlier", "centerProjectedTo"):
        assert np.array_equal(a["lin"][k], b["lin"][k]) and np.array_equal(a["lin2"][k], b["lin2"][k]), k
    assert a["lin2"]["energy"] == b["lin2"]["energy"] and a["E"] == b["E"]
    for (Ha, ba_), (Hb, bb) in zip(a["sys3"], b["sys3"]):
        assert np.array_equal(Ha, Hb) and np.array_equal(ba_, bb)
    assert np.array_equal(a["step"], b["step"])
    for k in a["pr"]:
        assert np.array_equal(a["pr"][k], b["pr"][k]), k
    ra, rb = a["rr"], b["rr"]
    for k in ra:
        va, vb = np.asarray(ra[k]), np.asarray(rb[k])
        if va.ndim == 2 and va.shape[1] == 74:
            # geometric + summed floats always; per-pixel rows for the linearised residuals
            geo = np.r_[8:30, 62:74]
            assert np.array_equal(va[:, geo], vb[:, geo])
            m = is_lin.astype(bool)
            assert m.any() and np.array_equal(va[m], vb[m])
        else:
            assert np.array_equal(va, vb), k
    for f in frames:
        f.close()


def test_download_stats_matches_stepwise(ctx, capi, synth, window_small):
    """hslam_ba_download_stats (the three doubles behind the systems that ride through hslam_ba_allreduce) == energy of
    linearizeAll and resInA / resInL of the step-wise calls."""
    win, is_lin = window_small
    frames = [capi.Frame(ctx, im, 1) for im in win.images]
    a = capi.EnergyFunctional(ctx); a.set_window(win, frames); a.set_topology(win)
    b = capi.EnergyFunctional(ctx); b.set_window(win, frames); b.set_topology(win)
    for ef in (a, b):
        ef.linearizeAll(False, outputs=False); ef.applyRes()
        ef.fixLinearization(np.nonzero(is_lin)[0].astype(np.int32))
    lin = a.linearizeAll(False, outputs=False); a.applyRes(); a.accumulate_all()
    b.iteration_device()
    e, nA, nL = b.download_stats()
    assert e == lin["energy"] and (nA, nL) == (a.resInA, a.resInL) and nA > 0 and nL > 0
    a.close(); b.close()
    for f in frames:
        f.close()
