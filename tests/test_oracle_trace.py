"""Pins the traceOn restatement (oracle/oracle_trace.cpp) with known answers on a rendered keyframe pair: the traced
inverse-depth interval brackets the true inverse depth, the interval shrinks from trace to trace, points that leave
the image go OOB and stay OOB."""
import copy

import numpy as np


def _setup(synth, O, w=320, h=240):
    kf = synth.render_keyframes(w, h, 4, seed=77)
    ims = synth.make_immature_points(kf, hosts=[0, 1], points_per_host=300, seed=3)
    dIs = [O.pyramid(im, 1)[0][0] for im in kf.images]
    return kf, ims, dIs


def test_trace_brackets_true_depth_and_shrinks(O, synth):
    kf, ims, dIs = _setup(synth, O)
    synth.trace_tables(kf, [0, 1], 2, ims)
    O.trace_immature(dIs[2], ims)
    st = ims.status
    good = st == 0
    assert good.mean() > 0.5, np.bincount(st, minlength=6)
    assert set(np.unique(st)) <= {0, 1, 2, 3, 4}
    inside = (ims.idepth_min[good] <= ims.idepth_true[good] * 1.02) & (ims.idepth_max[good] >= ims.idepth_true[good] * 0.98)
    assert inside.mean() > 0.9, inside.mean()
    assert np.all(ims.lastTraceUV[st == 1] == -1) and np.all(ims.lastTracePixelInterval[st == 1] == 0)
    w1 = (ims.idepth_max - ims.idepth_min)[good].copy()
    # second trace into the next frame (larger baseline): finite idepth_max branch; intervals do not grow
    first = copy.deepcopy(ims)
    synth.trace_tables(kf, [0, 1], 3, ims)
    O.trace_immature(dIs[3], ims)
    both = good & (ims.status == 0)
    assert both.sum() > 50
    w2 = (ims.idepth_max - ims.idepth_min)[both]
    assert np.median(w2) < np.median((first.idepth_max - first.idepth_min)[both])
    assert np.all(ims.status[first.status == 1] == 1)          # OOB is final (ImmaturePoint.cpp:43)
    assert len(w1) > 0


def test_trace_identity_motion_is_oob_or_skipped(O, synth):
    """No baseline: Kt = 0, the projected interval has zero length -> 'too certain' (SKIPPED) once idepth_max is finite."""
    kf, ims, dIs = _setup(synth, O)
    synth.trace_tables(kf, [0, 1], 2, ims)
    ims.KRKi[:] = np.eye(3, dtype=np.float32).reshape(-1); ims.Kt[:] = 0
    ims.idepth_min[:] = 0.2; ims.idepth_max[:] = 0.4
    O.trace_immature(dIs[0], ims)
    assert np.all(ims.status == 3)
    assert np.allclose(ims.lastTraceUV[:, 0], ims.u) and np.all(ims.lastTracePixelInterval == 0)


def test_activation_converges_to_true_depth(O, synth):
    """optimizeImmaturePoint restatement: after two traces the candidates whose interval is tight are activated and the
    1-D Gauss-Newton lands on the true inverse depth; candidates with a hopeless interval are rejected."""
    kf, ims, dIs = _setup(synth, O)
    for new in (2, 3):
        synth.trace_tables(kf, [0, 1], new, ims)
        O.trace_immature(dIs[new], ims)
    cand = np.nonzero((ims.status == 0) & np.isfinite(ims.idepth_max))[0]
    assert len(cand) > 100
    import copy
    sub = copy.deepcopy(ims)
    for k in ("host_idx", "u", "v", "gradH", "color", "weights", "energyTH", "idepth_min", "idepth_max", "idepth_true"):
        setattr(sub, k, np.ascontiguousarray(getattr(ims, k)[cand]))
    ok, idp, st = O.activate_immature(dIs, kf.cam, synth.activation_tables(kf), sub, minObs=2)
    assert ok.mean() > 0.7, ok.mean()
    rel = np.abs(idp[ok == 1] - sub.idepth_true[ok == 1]) / sub.idepth_true[ok == 1]
    assert np.median(rel) < 0.01 and (rel < 0.05).mean() > 0.9, (np.median(rel), (rel < 0.05).mean())
    # the host's own slot is OOB, activated points have >= minObs residuals IN
    assert np.all(st[np.arange(len(cand)), sub.host_idx] == 1)
    assert np.all((st[ok == 1] == 0).sum(1) >= 2)
    assert np.all(idp[ok == 0] == 0)
