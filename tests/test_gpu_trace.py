"""GPU parity of the immature-point epipolar search (hslam_trace_immature, ImmaturePoint::traceOn) with the oracle:
every output — status, idepth interval, quality, lastTraceUV, lastTracePixelInterval — bit for bit, over two consecutive
traces (infinite and finite idepth_max branches) and at 640x480 with the points of a whole window."""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FIELDS = ("status", "idepth_min", "idepth_max", "quality", "lastTraceUV", "lastTracePixelInterval")


def _same(a, b):
    for k in FIELDS:
        x, y = getattr(a, k), getattr(b, k)
        assert np.array_equal(x, y, equal_nan=True), (k, int((x != y).sum()))


@pytest.mark.parametrize("w,h,hosts,pph", [(320, 240, [0, 1], 400), (640, 480, [0, 1, 2, 3, 4], 2000)])
def test_trace_bit_exact(ctx, capi, O, synth, w, h, hosts, pph):
    kf = synth.render_keyframes(w, h, len(hosts) + 2, seed=78)
    ims_g = synth.make_immature_points(kf, hosts=hosts, points_per_host=pph, seed=5)
    ims_o = copy.deepcopy(ims_g)
    for new in (len(hosts), len(hosts) + 1):
        synth.trace_tables(kf, hosts, new, ims_g); synth.trace_tables(kf, hosts, new, ims_o)
        frame = capi.Frame(ctx, kf.images[new], 1)
        capi.traceNewCoarse(ctx, frame, ims_g)
        O.trace_immature(O.pyramid(kf.images[new], 1)[0][0], ims_o)
        _same(ims_g, ims_o)
        frame.close()
    st = np.bincount(ims_g.status, minlength=6)
    assert st[0] > 0.3 * len(ims_g.u) and st[1] > 0           # good traces and some OOB both occur


def test_trace_bad_arguments(ctx, capi, synth):
    kf = synth.render_keyframes(160, 120, 3, seed=1)
    ims = synth.make_immature_points(kf, hosts=[0], points_per_host=20, seed=5)
    synth.trace_tables(kf, [0], 1, ims)
    frame = capi.Frame(ctx, kf.images[1], 1)
    ims.host_idx[3] = 7
    with pytest.raises(capi.HslamError):
        capi.traceNewCoarse(ctx, frame, ims)
    frame.close()


def test_activation_bit_exact(ctx, capi, O, synth):
    """hslam_activate_immature (System::optimizeImmaturePoint) == the oracle: activation flag, optimised idepth and the
    final residual states, bit for bit, on the candidates left after two traces plus deliberately bad intervals."""
    w, h, hosts = 640, 480, [0, 1, 2, 3]
    kf = synth.render_keyframes(w, h, 6, seed=78)
    ims = synth.make_immature_points(kf, hosts=hosts, points_per_host=1500, seed=5)
    dIs = [O.pyramid(im, 1)[0][0] for im in kf.images]
    for new in (4, 5):
        synth.trace_tables(kf, hosts, new, ims)
        O.trace_immature(dIs[new], ims)
    keep = np.nonzero(np.isfinite(ims.idepth_max) & (ims.status != 1))[0]
    sub = copy.deepcopy(ims)
    for k in ("host_idx", "u", "v", "gradH", "color", "weights", "energyTH", "idepth_min", "idepth_max", "idepth_true"):
        setattr(sub, k, np.ascontiguousarray(getattr(ims, k)[keep]))
    sub.idepth_min[::17] *= np.float32(3.0)          # broken intervals: rejected / OOB paths
    sub.idepth_max[::17] *= np.float32(5.0)
    T = synth.activation_tables(kf)
    frames = [capi.Frame(ctx, im, 1) for im in kf.images]
    g = capi.optimizeImmaturePoints(ctx, frames, kf.cam, T, sub, minObs=2)
    o = O.activate_immature(dIs, kf.cam, T, sub, minObs=2)
    assert np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1]) and np.array_equal(g[2], o[2])
    assert 0.3 < g[0].mean() < 1.0 and len(keep) > 2000
    for f in frames:
        f.close()


def test_trace_and_activation_empty_and_all_oob(ctx, capi, synth):
    """Edge cases: no points at all; every point already OOB (traceOn returns at once and leaves every member untouched)."""
    kf = synth.render_keyframes(160, 120, 3, seed=2)
    ims = synth.make_immature_points(kf, hosts=[0], points_per_host=50, seed=5)
    synth.trace_tables(kf, [0], 1, ims)
    frame = capi.Frame(ctx, kf.images[1], 1)
    empty = copy.deepcopy(ims)
    for k in ("host_idx", "u", "v", "gradH", "color", "weights", "energyTH", "idepth_min", "idepth_max", "quality", "status",
              "lastTraceUV", "lastTracePixelInterval"):
        setattr(empty, k, getattr(ims, k)[:0].copy())
    capi.traceNewCoarse(ctx, frame, empty)                       # n = 0: no launch, no error
    oob = copy.deepcopy(ims)
    oob.status[:] = 1
    before = copy.deepcopy(oob)
    capi.traceNewCoarse(ctx, frame, oob)
    for k in ("status", "idepth_min", "quality", "lastTraceUV", "lastTracePixelInterval"):
        assert np.array_equal(getattr(oob, k), getattr(before, k), equal_nan=True), k
    frames = [capi.Frame(ctx, im, 1) for im in kf.images]
    ok, idp, st = capi.optimizeImmaturePoints(ctx, frames, kf.cam, synth.activation_tables(kf), empty)
    assert len(ok) == 0 and st.shape == (0, 3)
    for f in frames + [frame]:
        f.close()
