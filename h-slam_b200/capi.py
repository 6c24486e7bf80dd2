"""ctypes binding of include/hslam_b200.h plus thin Python mirrors of the reference classes.

Everything here goes through the C ABI of libhslam_b200.so (hand-written sm_100a kernels).  There
is no CPU path: if the library is missing or no CUDA device is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HSLAM_B200_LIB", os.path.join(_HERE, "libhslam_b200.so"))

ACC_A, ACC_L, ACC_SC = 0, 1, 2
RES_IN, RES_OOB, RES_OUT = 0, 1, 2


class HslamError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [
        ("huberTH", C.c_float), ("coarseCutoffTH", C.c_float), ("outlierTHSumComponent", C.c_float),
        ("affineOptModeA", C.c_float), ("affineOptModeB", C.c_float), ("scale_xi_rot", C.c_float),
        ("scale_xi_trans", C.c_float), ("scale_a", C.c_float), ("scale_b", C.c_float), ("scale_f", C.c_float),
        ("scale_c", C.c_float), ("scale_idepth", C.c_float), ("max_tracker_iterations", C.c_int * 5),
        ("frameEnergyTHN", C.c_float), ("frameEnergyTHFacMedian", C.c_float), ("frameEnergyTHConstWeight", C.c_float),
        ("overallEnergyTHWeight", C.c_float),
    ]


class WindowStruct(C.Structure):
    _fields_ = [
        ("n_frames", C.c_int), ("w", C.c_int), ("h", C.c_int),
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("fxi", C.c_float), ("fyi", C.c_float),
        ("frames", C.c_void_p), ("frameEnergyTH", C.c_void_p), ("precalc", C.c_void_p), ("adHost", C.c_void_p),
        ("adTarget", C.c_void_p), ("adHTdeltaF", C.c_void_p), ("cDeltaF", C.c_float * 4), ("cPrior", C.c_double * 4),
        ("framePrior", C.c_void_p), ("frameDeltaPrior", C.c_void_p),
    ]


class PointSoA(C.Structure):
    _fields_ = [("n_points", C.c_int)] + [(k, C.c_void_p) for k in (
        "u", "v", "idepth", "idepth_zero", "color", "weights", "priorF", "deltaF", "host_idx", "res_begin", "res_index")]


class ResidualSoA(C.Structure):
    _fields_ = [("n_residuals", C.c_int)] + [(k, C.c_void_p) for k in (
        "point_idx", "target_idx", "state_state", "is_linearized", "is_active", "state_energy")]


class StepIO(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("window", "idepth", "deltaF", "idepth_zero", "energy", "newState", "newEnergy",
                                          "newEnergyWithOutlier", "Hb", "nres2", "frameEnergyTH_out")] + [("newest_frame", C.c_int32), ("derive_deltaF", C.c_int32)]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


_lib = None


def load():
    """Load the CUDA library; raises loudly if it was not built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HslamError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                             "or `make -C h-slam_b200/csrc` (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.hslam_last_error.restype = C.c_char_p
        L.hslam_launch_count.restype = C.c_int64
        L.hslam_launch_count.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _chk(rc):
    if rc != 0:
        raise HslamError(f"hslam error {rc}: {load().hslam_last_error().decode()}")


def default_params():
    p = Params()
    load().hslam_params_default(C.byref(p))
    return p


class Context:
    def __init__(self, device=0, params=None):
        self.L = load()
        self.h = C.c_void_p()
        self.params = params or default_params()
        _chk(self.L.hslam_create(C.byref(self.params), C.c_int(device), C.byref(self.h)))

    def close(self):
        if self.h:
            self.L.hslam_destroy(self.h)
            self.h = C.c_void_p()

    def synchronize(self):
        _chk(self.L.hslam_synchronize(self.h))

    @property
    def stream(self):
        s = C.c_void_p()
        _chk(self.L.hslam_get_stream(self.h, C.byref(s)))
        return s.value or 0

    @property
    def launches(self):
        return int(self.L.hslam_launch_count(self.h))

    def host_empty(self, shape, dtype):
        """numpy array in page-locked memory owned by the context (hslam_host_alloc): copies are direct DMA."""
        dt = np.dtype(dtype)
        n = int(np.prod(shape)) if np.ndim(shape) else int(shape)
        p = C.c_void_p()
        _chk(self.L.hslam_host_alloc(self.h, C.c_size_t(max(n * dt.itemsize, 1)), C.byref(p)))
        buf = (C.c_char * (n * dt.itemsize)).from_address(p.value)
        return np.frombuffer(buf, dtype=dt, count=n).reshape(shape)

    def profile_enable(self, on=True):
        _chk(self.L.hslam_profile_enable(self.h, C.c_int(int(on))))

    def profile_read(self):
        """{kernel class: (summed ms, launches)} measured with CUDA events on the context stream."""
        K = 19
        ms = np.zeros(K, np.float32); cnt = np.zeros(K, np.int32)
        _chk(self.L.hslam_profile_read(self.h, _ptr(ms), _ptr(cnt)))
        self.L.hslam_kernel_name.restype = C.c_char_p
        return {self.L.hslam_kernel_name(k).decode(): (float(ms[k]), int(cnt[k])) for k in range(K) if cnt[k] > 0}


class Frame:
    """Device mirror of Frame::DirPyr / absSquaredGrad (reference Src/Frame.cpp:104-181)."""

    def __init__(self, ctx: Context, img, levels, gamma=None, device_ptr=None, w=None, h=None):
        self.ctx, self.L = ctx, ctx.L
        self.h = C.c_void_p()
        g = np.ascontiguousarray(gamma, np.float32) if gamma is not None else None
        if device_ptr is not None:
            self.w, self.hh = w, h
            _chk(self.L.hslam_frame_create_from_device(ctx.h, C.c_void_p(device_ptr), w, h, levels, _ptr(g), C.byref(self.h)))
        else:
            img = np.ascontiguousarray(img, np.float32)
            self.hh, self.w = img.shape
            _chk(self.L.hslam_frame_create(ctx.h, _ptr(img), self.w, self.hh, levels, _ptr(g), C.byref(self.h)))
        self.levels = levels

    def update(self, img=None, device_ptr=None):
        if device_ptr is not None:
            _chk(self.L.hslam_frame_update(self.h, C.c_void_p(device_ptr), 1))
        else:
            img = np.ascontiguousarray(img, np.float32)
            _chk(self.L.hslam_frame_update(self.h, _ptr(img), 0))

    def download(self, lvl):
        w, h = self.w >> lvl, self.hh >> lvl
        dI = np.zeros((h, w, 3), np.float32)
        ag = np.zeros((h, w), np.float32)
        _chk(self.L.hslam_frame_download(self.h, lvl, _ptr(dI), _ptr(ag)))
        return dI, ag

    def close(self):
        if self.h:
            self.L.hslam_frame_destroy(self.h)
            self.h = C.c_void_p()


class CoarseTracker:
    """Mirror of FSLAM::CoarseTracker (reference Include/CoarseTracker.h:14-95)."""

    def __init__(self, ctx: Context, w, h):
        self.ctx, self.L = ctx, ctx.L
        self.h = C.c_void_p()
        _chk(self.L.hslam_tracker_create(ctx.h, w, h, C.byref(self.h)))
        self.lastResiduals = np.full(5, np.nan)
        self.lastFlowIndicators = np.full(3, 1000.0)
        self.iterations = np.zeros(5, np.int32)

    def set_mode(self, cluster_ctas):
        _chk(self.L.hslam_tracker_set_mode(self.h, C.c_int(cluster_ctas)))

    def get_mode(self):
        v = C.c_int(0)
        _chk(self.L.hslam_tracker_get_mode(self.h, C.byref(v)))
        return v.value

    def last_profile(self):
        cyc = (C.c_longlong * 4)(); ev = C.c_int(0)
        _chk(self.L.hslam_tracker_last_profile(self.h, cyc, C.byref(ev)))
        return [int(c) for c in cyc], ev.value

    def makeK(self, cam, levels):
        self.levels = levels
        _chk(self.L.hslam_tracker_make_K(self.h, levels, cam.w, cam.h, C.c_float(cam.fx), C.c_float(cam.fy),
                                         C.c_float(cam.cx), C.c_float(cam.cy)))

    def get_K(self, lvl):
        k = np.zeros(4, np.float32); wh = np.zeros(2, np.int32); ki = np.zeros(9, np.float32)
        _chk(self.L.hslam_tracker_get_K(self.h, lvl, _ptr(k), _ptr(wh), _ptr(ki)))
        return k, wh, ki

    def setCoarseTrackingRef(self, pc, ref_exposure=1.0, ref_aff=(0.0, 0.0)):
        """pc: per level (u, v, idepth, color) as produced by makeCoarseDepthL0."""
        for l, (u, v, idp, col) in enumerate(pc):
            a = [np.ascontiguousarray(x, np.float32) for x in (u, v, idp, col)]
            _chk(self.L.hslam_tracker_set_ref(self.h, l, _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]), len(a[0])))
        _chk(self.L.hslam_tracker_set_ref_aff(self.h, C.c_float(ref_exposure), C.c_double(ref_aff[0]), C.c_double(ref_aff[1])))

    def makeCoarseDepthL0(self, lastRef: Frame, centerProjectedTo, HdiF, ref_exposure=1.0, ref_aff=(0.0, 0.0)):
        """CoarseTracker::makeCoarseDepthL0 on the device from per-point (centerProjectedTo[3], HdiF)."""
        c = np.ascontiguousarray(centerProjectedTo, np.float32).reshape(-1, 3)
        hd = np.ascontiguousarray(HdiF, np.float32)
        _chk(self.L.hslam_tracker_make_coarse_depth(self.h, lastRef.h, C.c_int(len(hd)), _ptr(c), _ptr(hd)))
        _chk(self.L.hslam_tracker_set_ref_aff(self.h, C.c_float(ref_exposure), C.c_double(ref_aff[0]), C.c_double(ref_aff[1])))

    def makeCoarseDepthL0_from_ba(self, lastRef: Frame, ef, residual_ids, ref_exposure=1.0, ref_aff=(0.0, 0.0)):
        ids = np.ascontiguousarray(residual_ids, np.int32)
        _chk(self.L.hslam_tracker_make_coarse_depth_from_ba(self.h, lastRef.h, ef.h, _ptr(ids), C.c_int(len(ids))))
        _chk(self.L.hslam_tracker_set_ref_aff(self.h, C.c_float(ref_exposure), C.c_double(ref_aff[0]), C.c_double(ref_aff[1])))

    def get_ref(self, lvl):
        n = C.c_int(0)
        _chk(self.L.hslam_tracker_get_ref(self.h, lvl, 0, None, None, None, None, C.byref(n)))
        a = [np.zeros(n.value, np.float32) for _ in range(4)]
        _chk(self.L.hslam_tracker_get_ref(self.h, lvl, n.value, _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]), C.byref(n)))
        return tuple(a)

    def calcRes(self, lvl, frame: Frame, T, aff, cutoffTH, new_exposure=1.0):
        out6 = np.zeros(6)
        nw = C.c_int(0)
        T = np.ascontiguousarray(T, np.float64)
        _chk(self.L.hslam_tracker_calc_res(self.h, lvl, frame.h, _ptr(T), C.c_double(aff[0]), C.c_double(aff[1]),
                                           C.c_float(new_exposure), C.c_float(cutoffTH), _ptr(out6), C.byref(nw)))
        return out6, nw.value

    def calcGSSSE(self, lvl, frame: Frame, T, aff, cutoffTH, new_exposure=1.0):
        H = np.zeros(64); b = np.zeros(8)
        T = np.ascontiguousarray(T, np.float64)
        _chk(self.L.hslam_tracker_calc_gs(self.h, lvl, frame.h, _ptr(T), C.c_double(aff[0]), C.c_double(aff[1]),
                                          C.c_float(new_exposure), C.c_float(cutoffTH), _ptr(H), _ptr(b)))
        return H.reshape(8, 8).T.copy(), b

    def trackNewestCoarse(self, frame: Frame, T, aff, coarsestLvl, minResForAbort=None, new_exposure=1.0):
        T = np.array(T, np.float64).copy()
        aff = np.array(aff, np.float64).copy()
        mra = np.full(5, np.nan) if minResForAbort is None else np.ascontiguousarray(minResForAbort, np.float64)
        ok = C.c_int(0)
        _chk(self.L.hslam_tracker_track(self.h, frame.h, C.c_float(new_exposure), _ptr(T), _ptr(aff), C.c_int(coarsestLvl),
                                        _ptr(mra), _ptr(self.lastResiduals), _ptr(self.lastFlowIndicators), C.byref(ok),
                                        _ptr(self.iterations)))
        return bool(ok.value), T, aff

    def trackNewestCoarse_batch(self, frames, Ts, affs, coarsestLvl, minResForAbort=None, new_exposures=None):
        """hslam_tracker_track_batch: n members (frames and/or initial poses) in ONE launch, one cluster each.
        Returns dict(ok[n], T[n,12], aff[n,2], lastResiduals[n,5], lastFlow[n,3], iterations[n,5])."""
        n = len(frames)
        T = np.ascontiguousarray(np.array(Ts, np.float64).reshape(n, 12)).copy()
        aff = np.ascontiguousarray(np.array(affs, np.float64).reshape(n, 2)).copy()
        mra = np.full(5, np.nan) if minResForAbort is None else np.ascontiguousarray(minResForAbort, np.float64)
        ne = None if new_exposures is None else np.ascontiguousarray(new_exposures, np.float32)
        fr = (C.c_void_p * n)(*[f.h for f in frames])
        res = np.zeros((n, 5)); flow = np.zeros((n, 3)); ok = np.zeros(n, np.int32); its = np.zeros((n, 5), np.int32)
        _chk(self.L.hslam_tracker_track_batch(self.h, C.c_int(n), fr, _ptr(ne), _ptr(T), _ptr(aff), C.c_int(coarsestLvl), _ptr(mra),
                                              _ptr(res), _ptr(flow), _ptr(ok), _ptr(its)))
        return dict(ok=ok.astype(bool), T=T, aff=aff, lastResiduals=res, lastFlow=flow, iterations=its)

    def track_enqueue(self, frame: Frame, T, aff, coarsestLvl, minResForAbort=None, new_exposure=1.0):
        T = np.ascontiguousarray(T, np.float64)
        aff = np.ascontiguousarray(aff, np.float64)
        mra = np.full(5, np.nan) if minResForAbort is None else np.ascontiguousarray(minResForAbort, np.float64)
        _chk(self.L.hslam_tracker_track_enqueue(self.h, frame.h, C.c_float(new_exposure), _ptr(T), _ptr(aff),
                                                C.c_int(coarsestLvl), _ptr(mra)))

    def track_finish(self):
        T = np.zeros(12); aff = np.zeros(2); ok = C.c_int(0)
        _chk(self.L.hslam_tracker_track_finish(self.h, _ptr(T), _ptr(aff), _ptr(self.lastResiduals),
                                               _ptr(self.lastFlowIndicators), C.byref(ok), _ptr(self.iterations)))
        return bool(ok.value), T, aff

    def close(self):
        if self.h:
            self.L.hslam_tracker_destroy(self.h)
            self.h = C.c_void_p()


class TraceSettings(C.Structure):
    _fields_ = [(k, C.c_float) for k in ("maxPixSearch", "slackInterval", "stepsize", "minImprovementFactor", "GNThreshold",
                                         "extraSlackOnTH", "huberTH")] + [("minTraceTestRadius", C.c_int), ("GNIterations", C.c_int)]


class ImmatureSoA(C.Structure):
    _fields_ = [("n_points", C.c_int)] + [(k, C.c_void_p) for k in (
        "host_idx", "u", "v", "gradH", "color", "weights", "energyTH", "idepth_min", "idepth_max", "quality", "lastTraceStatus",
        "lastTraceUV", "lastTracePixelInterval")]


def traceNewCoarse(ctx: "Context", frame: "Frame", ims, settings=None):
    """System::traceNewCoarse (Src/Mapping.cpp:494-538): ImmaturePoint::traceOn for every point of `ims` (an object
    with the ImmatureSet fields of synth.py) against `frame`; the in/out members are updated in place."""
    keep = dict(host_idx=np.ascontiguousarray(ims.host_idx, np.int32), u=np.ascontiguousarray(ims.u, np.float32),
                v=np.ascontiguousarray(ims.v, np.float32), gradH=np.ascontiguousarray(ims.gradH, np.float32),
                color=np.ascontiguousarray(ims.color, np.float32), weights=np.ascontiguousarray(ims.weights, np.float32),
                energyTH=np.ascontiguousarray(ims.energyTH, np.float32))
    for k in ("idepth_min", "idepth_max", "quality", "lastTraceUV", "lastTracePixelInterval"):
        assert getattr(ims, k).dtype == np.float32 and getattr(ims, k).flags.c_contiguous
    assert ims.status.dtype == np.int8
    s = ImmatureSoA(len(keep["u"]))
    for k, a in keep.items():
        setattr(s, k, _ptr(a))
    s.idepth_min, s.idepth_max, s.quality = _ptr(ims.idepth_min), _ptr(ims.idepth_max), _ptr(ims.quality)
    s.lastTraceStatus, s.lastTraceUV, s.lastTracePixelInterval = _ptr(ims.status), _ptr(ims.lastTraceUV), _ptr(ims.lastTracePixelInterval)
    K = np.ascontiguousarray(ims.KRKi, np.float32); T = np.ascontiguousarray(ims.Kt, np.float32); A = np.ascontiguousarray(ims.aff, np.float32)
    _chk(ctx.L.hslam_trace_immature(ctx.h, frame.h, C.byref(settings) if settings is not None else None, C.c_int(len(K)),
                                    _ptr(K), _ptr(T), _ptr(A), C.byref(s)))
    return ims


class ActivateSoA(C.Structure):
    _fields_ = [("n_points", C.c_int)] + [(k, C.c_void_p) for k in (
        "host_idx", "u", "v", "color", "weights", "energyTH", "idepth_min", "idepth_max", "ok_out", "idepth_out", "res_state_out")]


def optimizeImmaturePoints(ctx: "Context", frames, cam, RtAff, ims, minObs=1, minIdepthH_act=100.0, GNIts=3):
    """System::optimizeImmaturePoint for every point of `ims`: returns (ok int8[n], idepth float32[n], res_state int8[n, n_frames])."""
    n, nf = len(ims.u), len(frames)
    keep = dict(host_idx=np.ascontiguousarray(ims.host_idx, np.int32), u=np.ascontiguousarray(ims.u, np.float32),
                v=np.ascontiguousarray(ims.v, np.float32), color=np.ascontiguousarray(ims.color, np.float32),
                weights=np.ascontiguousarray(ims.weights, np.float32), energyTH=np.ascontiguousarray(ims.energyTH, np.float32),
                idepth_min=np.ascontiguousarray(ims.idepth_min, np.float32), idepth_max=np.ascontiguousarray(ims.idepth_max, np.float32))
    ok = np.zeros(n, np.int8); idp = np.zeros(n, np.float32); st = np.zeros((n, nf), np.int8)
    s = ActivateSoA(n)
    for k, a in keep.items():
        setattr(s, k, _ptr(a))
    s.ok_out, s.idepth_out, s.res_state_out = _ptr(ok), _ptr(idp), _ptr(st)
    fr = (C.c_void_p * nf)(*[f.h for f in frames])
    T = np.ascontiguousarray(RtAff, np.float32)
    fxi, fyi = np.float32(1.0) / np.float32(cam.fx), np.float32(1.0) / np.float32(cam.fy)
    _chk(ctx.L.hslam_activate_immature(ctx.h, C.c_int(nf), fr, C.c_float(cam.fx), C.c_float(cam.fy), C.c_float(cam.cx), C.c_float(cam.cy),
                                       C.c_float(fxi), C.c_float(fyi), _ptr(T), C.c_float(minIdepthH_act), C.c_int(GNIts), C.c_int(minObs),
                                       C.byref(s)))
    return ok, idp, st


class EnergyFunctional:
    """Mirror of FSLAM::EnergyFunctional's accumulate path + the residual linearisation it consumes
    (reference Include/EnergyFunctional.h:30-121, Src/FullSystemOptimize.cpp:19-165)."""

    def __init__(self, ctx: Context):
        self.ctx, self.L = ctx, ctx.L
        self.h = C.c_void_p()
        _chk(self.L.hslam_ba_create(ctx.h, C.byref(self.h)))
        self._keep = []
        self.resInA = self.resInL = 0
        self._async = False
        self._alive = []      # output buffers the C side fills at wait()
        self._pendings = []

    class Pending:
        """Result of a call made between begin_async() and wait(); .get() after wait()."""

        def __init__(self, fn):
            self.fn, self.value, self.done = fn, None, False

        def get(self):
            if not self.done:
                raise HslamError("result requested before wait()")
            return self.value

    def _result(self, keep, fn):
        if not self._async:
            return fn()
        self._alive.append(keep)
        p = EnergyFunctional.Pending(fn)
        self._pendings.append(p)
        return p

    @staticmethod
    def window_struct(win, frames):
        """(hslam_window, keep-alive list) for `win`; the arrays are referenced, not copied."""
        cam = win.cam
        ws = WindowStruct()
        ws.n_frames, ws.w, ws.h = win.n, cam.w, cam.h
        ws.fx, ws.fy, ws.cx, ws.cy = cam.fx, cam.fy, cam.cx, cam.cy
        ws.fxi, ws.fyi = np.float32(1.0) / np.float32(cam.fx), np.float32(1.0) / np.float32(cam.fy)
        fr = (C.c_void_p * win.n)(*[f.h for f in frames])
        ws.frames = C.cast(fr, C.c_void_p)
        ka = dict(frameEnergyTH=np.ascontiguousarray(win.frameEnergyTH, np.float32), precalc=np.ascontiguousarray(win.precalc, np.float32),
                  adHost=np.ascontiguousarray(win.adHost, np.float64), adTarget=np.ascontiguousarray(win.adTarget, np.float64),
                  adHTdeltaF=np.ascontiguousarray(win.adHTdeltaF, np.float32), framePrior=np.ascontiguousarray(win.framePrior, np.float64),
                  frameDeltaPrior=np.ascontiguousarray(win.frameDeltaPrior, np.float64))
        for k, a in ka.items():
            setattr(ws, k, _ptr(a))
        for i in range(4):
            ws.cDeltaF[i] = float(win.cDeltaF[i])
            ws.cPrior[i] = float(win.cPrior[i])
        return ws, [fr, ka]

    def set_window(self, win, frames):
        ws, _keep = EnergyFunctional.window_struct(win, frames)
        _chk(self.L.hslam_ba_set_window(self.h, C.byref(ws)))
        self.n = win.n
        self.N = 8 * win.n + 4

    def set_topology(self, win):
        keep = dict(
            u=np.ascontiguousarray(win.u, np.float32), v=np.ascontiguousarray(win.v, np.float32),
            idepth=np.ascontiguousarray(win.idepth, np.float32), idepth_zero=np.ascontiguousarray(win.idepth_zero, np.float32),
            color=np.ascontiguousarray(win.color, np.float32), weights=np.ascontiguousarray(win.weights, np.float32),
            priorF=np.ascontiguousarray(win.priorF, np.float32), deltaF=np.ascontiguousarray(win.deltaF, np.float32),
            host_idx=np.ascontiguousarray(win.host_idx, np.int32), res_begin=np.ascontiguousarray(win.res_begin, np.int32),
            res_index=np.ascontiguousarray(win.res_index, np.int32))
        ps = PointSoA(len(keep["u"]))
        for k, a in keep.items():
            setattr(ps, k, _ptr(a))
        keep2 = dict(
            point_idx=np.ascontiguousarray(win.point_idx, np.int32), target_idx=np.ascontiguousarray(win.target_idx, np.int32),
            state_state=np.ascontiguousarray(win.state_state, np.int8), is_linearized=np.ascontiguousarray(win.is_linearized, np.uint8),
            is_active=np.ascontiguousarray(win.is_active, np.uint8))
        rs = ResidualSoA(len(keep2["point_idx"]))
        for k, a in keep2.items():
            setattr(rs, k, _ptr(a))
        rs.state_energy = None
        _chk(self.L.hslam_ba_set_topology(self.h, C.byref(ps), C.byref(rs)))
        self.P, self.R = ps.n_points, rs.n_residuals

    def update_points(self, idepth=None, deltaF=None, idepth_zero=None):
        a = [None if x is None else np.ascontiguousarray(x, np.float32) for x in (idepth, deltaF, idepth_zero)]
        _chk(self.L.hslam_ba_update_points(self.h, _ptr(a[0]), _ptr(a[1]), _ptr(a[2])))

    def linearizeAll(self, fixLinearization=False, outputs=True):
        R = self.R
        e = C.c_double(0)
        ns = ne = nwo = cpt = pto = None
        if outputs:
            ns = np.zeros(R, np.int8); ne = np.zeros(R, np.float32); nwo = np.zeros(R, np.float32)
        if outputs is True:
            cpt = np.zeros(R * 3, np.float32); pto = np.zeros(R * 16, np.float32)
        _chk(self.L.hslam_ba_linearize(self.h, C.c_int(int(fixLinearization)), C.byref(e), _ptr(ns), _ptr(ne), _ptr(nwo),
                                       _ptr(cpt), _ptr(pto)))

        def fin():
            if not outputs:
                return dict(energy=e.value)
            if outputs is not True:   # "states": what setNewFrameEnergyTH / the accept test need
                return dict(energy=e.value, newState=ns, newEnergy=ne, newEnergyWithOutlier=nwo)
            return dict(energy=e.value, newState=ns, newEnergy=ne, newEnergyWithOutlier=nwo,
                        centerProjectedTo=cpt.reshape(R, 3), projectedTo=pto.reshape(R, 8, 2))
        return self._result((e, ns, ne, nwo, cpt, pto), fin)

    def applyRes(self):
        _chk(self.L.hslam_ba_apply_res(self.h))

    def fixLinearization(self, ids):
        ids = np.ascontiguousarray(ids, np.int32)
        _chk(self.L.hslam_ba_fix_linearization(self.h, _ptr(ids), C.c_int(len(ids))))

    def _acc(self, which):
        N = self.N
        H = np.zeros(N * N); b = np.zeros(N); nres = C.c_int(0)
        _chk(self.L.hslam_ba_accumulate(self.h, C.c_int(which), _ptr(H), _ptr(b), C.byref(nres)))

        def fin():
            if which == ACC_A:
                self.resInA = nres.value
            elif which == ACC_L:
                self.resInL = nres.value
            return H.reshape(N, N).T.copy(), b
        return self._result((H, b, nres), fin)

    def accumulateAF_MT(self):
        return self._acc(ACC_A)

    def accumulateLF_MT(self):
        return self._acc(ACC_L)

    def accumulateSCF_MT(self):
        return self._acc(ACC_SC)

    def accumulate_all(self):
        N = self.N
        per = N * N + N
        out = np.zeros(3 * per)
        nres = np.zeros(2, np.int32)
        _chk(self.L.hslam_ba_accumulate_all(self.h, _ptr(out), _ptr(nres)))

        def fin():
            self.resInA, self.resInL = int(nres[0]), int(nres[1])
            res = []
            for s in range(3):
                blk = out[s * per:(s + 1) * per]
                res.append((blk[:N * N].reshape(N, N).T.copy(), blk[N * N:].copy()))
            return res
        return self._result((out, nres), fin)

    def iteration_device(self):
        """linearize + applyRes + accumulate A/L/SC enqueued with no host round trip; returns (device ptr, count)."""
        p = C.c_void_p(); cnt = C.c_int(0)
        _chk(self.L.hslam_ba_iteration_device(self.h, C.byref(p), C.byref(cnt)))
        return p.value, cnt.value

    def accumulate_all_device(self):
        p = C.c_void_p(); cnt = C.c_int(0)
        _chk(self.L.hslam_ba_accumulate_all_device(self.h, C.byref(p), C.byref(cnt)))
        return p.value, cnt.value

    def download_Hb(self):
        N = self.N
        per = N * N + N
        out = np.zeros(3 * per)
        _chk(self.L.hslam_ba_download_Hb(self.h, _ptr(out)))
        return self._result((out,), lambda: [(out[s * per:s * per + N * N].reshape(N, N).T.copy(),
                                              out[s * per + N * N:(s + 1) * per].copy()) for s in range(3)])

    def download_stats(self):
        """(energy, resInA, resInL) as the device holds them (window totals after allreduce())."""
        e = C.c_double(0); nres = np.zeros(2, np.int32)
        _chk(self.L.hslam_ba_download_stats(self.h, C.byref(e), _ptr(nres)))
        return self._result((e, nres), lambda: (e.value, int(nres[0]), int(nres[1])))

    def begin_async(self):
        _chk(self.L.hslam_ba_begin_async(self.h))
        self._async = True

    def wait(self):
        _chk(self.L.hslam_ba_wait(self.h))
        self._async = False
        for p in self._pendings:
            p.value, p.done = p.fn(), True
        self._pendings, self._alive = [], []

    def set_parallelism(self, target_blocks):
        _chk(self.L.hslam_ba_set_parallelism(self.h, C.c_int(int(target_blocks))))

    def get_parallelism(self):
        v = C.c_int(0)
        _chk(self.L.hslam_ba_get_parallelism(self.h, C.byref(v)))
        return v.value

    def set_add_priors(self, flag):
        _chk(self.L.hslam_ba_set_add_priors(self.h, C.c_int(int(flag))))

    def point_results(self):
        P = self.P
        names = ["HdiF", "bdSumF", "idepth_hessian", "Hdd_accAF", "bd_accAF", "Hcd_accAF", "Hdd_accLF", "bd_accLF", "Hcd_accLF"]
        arrs = [np.zeros(P * (4 if "Hcd" in k else 1), np.float32) for k in names]
        _chk(self.L.hslam_ba_point_results(self.h, *[_ptr(a) for a in arrs]))

        def fin():
            d = dict(zip(names, arrs))
            d["Hcd_accAF"] = d["Hcd_accAF"].reshape(P, 4); d["Hcd_accLF"] = d["Hcd_accLF"].reshape(P, 4)
            return d
        return self._result(arrs, fin)

    def residual_results(self, want_J=True):
        R = self.R
        st = np.zeros(R, np.int8); act = np.zeros(R, np.uint8); jp = np.zeros(R * 8, np.float32)
        j74 = np.zeros(R * 74, np.float32) if want_J else None
        rtz = np.zeros(R * 8, np.float32)
        _chk(self.L.hslam_ba_residual_results(self.h, _ptr(st), _ptr(act), _ptr(jp), _ptr(j74), _ptr(rtz)))
        return self._result((st, act, jp, j74, rtz), lambda: dict(
            state_state=st, is_active=act, JpJdF=jp.reshape(R, 8), J=None if j74 is None else j74.reshape(R, 74),
            res_toZeroF=rtz.reshape(R, 8)))

    def marginalizePointsF(self, point_ids, prior_scale):
        """(M, Mb, Msc, Mbsc, nres) of the points to marginalise (EnergyFunctional::marginalizePointsF)."""
        N = self.N
        M = np.zeros(N * N); Mb = np.zeros(N); Ms = np.zeros(N * N); Mbs = np.zeros(N); nres = C.c_int(0)
        ids = np.ascontiguousarray(point_ids, np.int32)
        _chk(self.L.hslam_ba_marginalize_points(self.h, _ptr(ids), C.c_int(len(ids)), C.c_float(prior_scale), _ptr(M), _ptr(Mb),
                                                _ptr(Ms), _ptr(Mbs), C.byref(nres)))
        return M.reshape(N, N).T.copy(), Mb, Ms.reshape(N, N).T.copy(), Mbs, nres.value

    def allreduce(self, comm: "NcclComm"):
        """Sharded window: ONE fp64 NCCL all-reduce of the stitched systems (+ energy, resInA/L) on the context stream."""
        _chk(self.L.hslam_ba_allreduce(self.h, comm.h))

    def iteration_allreduce(self, comm: "NcclComm"):
        """One GN iteration of this rank's shard (one CUDA graph launch) + the all-reduce on the same stream."""
        _chk(self.L.hslam_ba_iteration_allreduce(self.h, comm.h))

    def set_jacobian_mode(self, mode):
        """0: all 74 Jacobian floats stored; 1: GN-loop mode (per-pixel rows + projectedTo only on fixLinearization)."""
        _chk(self.L.hslam_ba_set_jacobian_mode(self.h, C.c_int(int(mode))))

    def setNewFrameEnergyTH(self, newest_frame):
        """System::setNewFrameEnergyTH (FullSystemOptimize.cpp:60-101) on the device; returns the new threshold."""
        th = np.zeros(1, np.float32)
        _chk(self.L.hslam_ba_new_frame_energy_th(self.h, C.c_int(int(newest_frame)), _ptr(th)))
        return self._result((th,), lambda: float(th[0]))

    def calcLEnergyF_MT(self):
        e = C.c_double(0)
        _chk(self.L.hslam_ba_calc_l_energy(self.h, C.byref(e)))
        return e.value

    def resubstituteF_MT(self, x):
        step = np.zeros(self.P, np.float32)
        _chk(self.L.hslam_ba_resubstitute(self.h, _ptr(np.ascontiguousarray(x, np.float64)), _ptr(step)))
        return self._result((step,), lambda: step)

    def close(self):
        if self.h:
            self.L.hslam_ba_destroy(self.h)
            self.h = C.c_void_p()


def use_torch_nccl():
    """Point the library at the libnccl that torch bundles (HSLAM_NCCL_LIB), so that hslam_nccl_* / hslam_ba_allreduce and
    torch.distributed share one NCCL build.  Must run before the first NCCL call of the library; a no-op if the
    variable is already set or the wheel is not there (then the library falls back to the process / libnccl.so.2)."""
    if os.environ.get("HSLAM_NCCL_LIB"):
        return os.environ["HSLAM_NCCL_LIB"]
    try:
        import nvidia.nccl as _n
        cand = os.path.join(list(_n.__path__)[0], "lib", "libnccl.so.2")
        if os.path.exists(cand):
            os.environ["HSLAM_NCCL_LIB"] = cand
            return cand
    except Exception:
        pass
    return None


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    _chk(load().hslam_nccl_get_unique_id(buf))
    return bytes(buf.raw)


class NcclComm:
    """ncclComm_t created through the library's own NCCL handle (hslam_nccl_comm_init)."""

    def __init__(self, ctx: Context, unique_id: bytes, n_ranks: int, rank: int):
        self.L = ctx.L
        self.h = C.c_void_p()
        _chk(self.L.hslam_nccl_comm_init(ctx.h, C.c_char_p(unique_id), C.c_int(n_ranks), C.c_int(rank), C.byref(self.h)))

    def close(self):
        if self.h:
            self.L.hslam_nccl_comm_destroy(self.h)
            self.h = C.c_void_p()


class EnergyFunctionalBatch:
    """A batch of independent windows stepped together: 4 launches per GN iteration for all of them."""

    def __init__(self, ctx: Context, efs):
        self.ctx, self.L = ctx, ctx.L
        self.h = C.c_void_p()
        self.efs = list(efs)
        arr = (C.c_void_p * len(self.efs))(*[e.h for e in self.efs])
        _chk(self.L.hslam_ba_batch_create(ctx.h, arr, C.c_int(len(self.efs)), C.byref(self.h)))

    def iteration(self):
        _chk(self.L.hslam_ba_batch_iteration(self.h))

    def set_newest_frame(self, newest_frame):
        """Device-resident GN loop: iteration() also runs System::setNewFrameEnergyTH for this frame of every window (-1: off)."""
        _chk(self.L.hslam_ba_batch_set_newest_frame(self.h, C.c_int(int(newest_frame))))

    def make_io(self, wins, frames_per_window, pinned=True, newest_frame=None):
        """Persistent host-side I/O block for step_host(): per window the hslam_window struct, the per-point
        inputs and the output arrays.  pinned="bound": the outputs are views of the batch's own page-locked
        result block (hslam_ba_batch_io_bind, zero-copy); True: separate page-locked arrays (hslam_host_alloc,
        direct DMA); False: ordinary numpy arrays (staged + memcpy)."""
        n = len(self.efs)
        arr = (StepIO * n)()
        keep, outs = [], []
        for i, (ef, w, fr) in enumerate(zip(self.efs, wins, frames_per_window)):
            ws, ka = EnergyFunctional.window_struct(w, fr)
            inp = dict(idepth=np.ascontiguousarray(w.idepth, np.float32), deltaF=np.ascontiguousarray(w.deltaF, np.float32))
            arr[i].window = C.cast(C.pointer(ws), C.c_void_p)
            arr[i].idepth, arr[i].deltaF, arr[i].idepth_zero = _ptr(inp["idepth"]), _ptr(inp["deltaF"]), None
            keep.append((ws, ka, inp))
        names = ("energy", "newState", "newEnergy", "newEnergyWithOutlier", "Hb", "nres2")
        if pinned == "bound":
            steps = (C.c_void_p * n)()
            _chk(self.L.hslam_ba_batch_io_bind(self.h, arr, steps))

            def view(addr, count, dt):
                dt = np.dtype(dt)
                buf = (C.c_char * (count * dt.itemsize)).from_address(addr)
                return np.frombuffer(buf, dtype=dt, count=count)
            nmax = max(w.n for w in wins)
            th_all = np.zeros((n, nmax), np.float32)      # Frame::frameEnergyTH of every window, one row each
            for i, (ef, w) in enumerate(zip(self.efs, wins)):
                N = 8 * w.n + 4
                # the per-point INPUT slots are bound as well: the caller's idepth / deltaF arrays ARE the pinned block
                ws, ka, _old = keep[i]
                inp = dict(idepth=view(arr[i].idepth, ef.P, np.float32), deltaF=view(arr[i].deltaF, ef.P, np.float32))
                inp["idepth"][:] = w.idepth; inp["deltaF"][:] = w.deltaF
                # ... and the window tables (hslam_ba_batch_window_bind): setPrecalcValues / setAdjointsF write them in place
                _chk(self.L.hslam_ba_batch_window_bind(self.h, C.c_int(i), C.byref(ws)))
                nn = w.n * w.n
                tabs = dict(precalc=view(ws.precalc, nn * 28, np.float32), adHTdeltaF=view(ws.adHTdeltaF, nn * 8, np.float32),
                            adHost=view(ws.adHost, nn * 64, np.float64), adTarget=view(ws.adTarget, nn * 64, np.float64),
                            framePrior=view(ws.framePrior, w.n * 8, np.float64), frameDeltaPrior=view(ws.frameDeltaPrior, w.n * 8, np.float64))
                for k, v in tabs.items():
                    v[:] = np.asarray(ka[1][k]).reshape(-1)
                th_all[i, :w.n] = ka[1]["frameEnergyTH"]
                tabs["frameEnergyTH"] = th_all[i, :w.n]
                ws.frameEnergyTH = th_all[i].ctypes.data
                keep[i] = (ws, [ka[0], tabs], inp)
                outs.append(dict(energy=view(arr[i].energy, 1, np.float64), newState=view(arr[i].newState, ef.R, np.int8),
                                 newEnergy=view(arr[i].newEnergy, ef.R, np.float32),
                                 newEnergyWithOutlier=view(arr[i].newEnergyWithOutlier, ef.R, np.float32),
                                 Hb=view(arr[i].Hb, 3 * (N * N + N), np.float64), nres2=view(arr[i].nres2, 2, np.int32),
                                 frameEnergyTH=view(arr[i].frameEnergyTH_out, 1, np.float32),
                                 step=view(steps[i], ef.P, np.float32)))
                arr[i].newest_frame = (w.n - 1) if newest_frame is None else int(newest_frame)
        else:
            mk = (lambda shape, dt: self.ctx.host_empty(shape, dt)) if pinned else (lambda shape, dt: np.zeros(shape, dt))
            for i, (ef, w) in enumerate(zip(self.efs, wins)):
                N = 8 * w.n + 4
                o = dict(energy=mk(1, np.float64), newState=mk(ef.R, np.int8), newEnergy=mk(ef.R, np.float32),
                         newEnergyWithOutlier=mk(ef.R, np.float32), Hb=mk(3 * (N * N + N), np.float64), nres2=mk(2, np.int32),
                         step=mk(ef.P, np.float32), frameEnergyTH=mk(1, np.float32))
                for k in names:
                    setattr(arr[i], k, _ptr(o[k]))
                arr[i].frameEnergyTH_out = _ptr(o["frameEnergyTH"])
                arr[i].newest_frame = (w.n - 1) if newest_frame is None else int(newest_frame)
                outs.append(o)
        res = dict(arr=arr, keep=keep, out=outs)
        if pinned == "bound":
            res["th_all"] = th_all
            # the threshold slots of the result block are equally spaced when the windows have the same size: one strided view
            addr = [arr[i].frameEnergyTH_out for i in range(n)]
            step = addr[1] - addr[0] if n > 1 else 4
            if n > 1 and all(addr[i + 1] - addr[i] == step for i in range(n - 1)) and step > 0 and step % 4 == 0:
                buf = (C.c_char * (step * (n - 1) + 4)).from_address(addr[0])
                res["th_out"] = np.ndarray((n,), np.float32, buffer=buf, strides=(step,))
                res["th_col"] = np.array([arr[i].newest_frame for i in range(n)])
        return res

    def feed_back_thresholds(self, io):
        """The caller keeps the new frameEnergyTH in its Frame and hands it back in the next hslam_window (what the reference does
        with newFrame->frame->frameEnergyTH); for bound io blocks this is one strided copy."""
        if "th_out" in io:
            io["th_all"][np.arange(len(self.efs)), io["th_col"]] = io["th_out"]
        else:
            for i, o in enumerate(io["out"]):
                io["keep"][i][1][1]["frameEnergyTH"][io["arr"][i].newest_frame] = o["frameEnergyTH"][0]

    def step_host(self, io):
        """One GN iteration of every window with host buffers (one H2D block, one D2H block, one wait)."""
        _chk(self.L.hslam_ba_batch_step_host(self.h, io["arr"]))
        return io["out"]

    def resubstitute_host(self, io, xs):
        n = len(self.efs)
        if io.get("xs_id") != id(xs) or any(not (isinstance(x, np.ndarray) and x.dtype == np.float64 and x.flags.c_contiguous) for x in xs):
            io["xa"] = [np.ascontiguousarray(x, np.float64) for x in xs]
            io["xp"] = (C.c_void_p * n)(*[a.ctypes.data for a in io["xa"]])
            io["xs_id"] = id(xs)      # the same list of fp64 arrays next time: the pointer table is reused (values are read by the call)
        if "step_ptrs" not in io:
            io["step_ptrs"] = (C.c_void_p * n)(*[o["step"].ctypes.data for o in io["out"]])
        _chk(self.L.hslam_ba_batch_resubstitute_host(self.h, io["xp"], io["step_ptrs"]))
        return [o["step"] for o in io["out"]]

    def close(self):
        if self.h:
            self.L.hslam_ba_batch_destroy(self.h)
            self.h = C.c_void_p()
