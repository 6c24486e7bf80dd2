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
