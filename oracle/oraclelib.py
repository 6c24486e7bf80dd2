"""ctypes bindings of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (h-slam_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


class Params(C.Structure):
    _fields_ = [
        ("huberTH", C.c_float), ("coarseCutoffTH", C.c_float), ("outlierTHSumComponent", C.c_float),
        ("affineOptModeA", C.c_float), ("affineOptModeB", C.c_float), ("scale_xi_rot", C.c_float),
        ("scale_xi_trans", C.c_float), ("scale_a", C.c_float), ("scale_b", C.c_float), ("scale_f", C.c_float),
        ("scale_c", C.c_float), ("scale_idepth", C.c_float), ("max_tracker_iterations", C.c_int * 5),
        ("frameEnergyTHN", C.c_float), ("frameEnergyTHFacMedian", C.c_float), ("frameEnergyTHConstWeight", C.c_float),
        ("overallEnergyTHWeight", C.c_float),
    ]

    @staticmethod
    def default():
        p = Params(9.0, 20.0, 2500.0, 1e12, 1e8, 1.0, 0.5, 10.0, 1000.0, 50.0, 50.0, 1.0)
        for i, v in enumerate([10, 20, 50, 50, 50]):
            p.max_tracker_iterations[i] = v
        p.frameEnergyTHN, p.frameEnergyTHFacMedian, p.frameEnergyTHConstWeight, p.overallEnergyTHWeight = 0.7, 1.5, 0.5, 1.0
        return p


class TrkLevel(C.Structure):
    _fields_ = [
        ("w", C.c_int), ("h", C.c_int), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
        ("K", C.c_float * 9), ("Ki", C.c_float * 9), ("n", C.c_int),
        ("pc_u", C.c_void_p), ("pc_v", C.c_void_p), ("pc_idepth", C.c_void_p), ("pc_color", C.c_void_p),
    ]


class FramePair(C.Structure):
    _fields_ = [("ref_exposure", C.c_float), ("new_exposure", C.c_float), ("ref_aff_a", C.c_double), ("ref_aff_b", C.c_double)]


class OrcWindow(C.Structure):
    _fields_ = [
        ("n_frames", C.c_int), ("w", C.c_int), ("h", C.c_int),
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("fxi", C.c_float), ("fyi", C.c_float),
        ("dI", C.c_void_p), ("frameEnergyTH", C.c_void_p), ("precalc", C.c_void_p), ("adHost", C.c_void_p),
        ("adTarget", C.c_void_p), ("adHTdeltaF", C.c_void_p), ("cDeltaF", C.c_float * 4), ("cPrior", C.c_double * 4),
        ("framePrior", C.c_void_p), ("frameDeltaPrior", C.c_void_p),
    ]


class PointSoA(C.Structure):
    _fields_ = [("n_points", C.c_int)] + [(k, C.c_void_p) for k in (
        "u", "v", "idepth", "idepth_zero", "color", "weights", "priorF", "deltaF", "host_idx", "res_begin", "res_index")]


class ResidualSoA(C.Structure):
    _fields_ = [("n_residuals", C.c_int)] + [(k, C.c_void_p) for k in (
        "point_idx", "target_idx", "state_state", "is_linearized", "is_active", "state_energy")]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def build(fast=False):
    name = "libhslam_oracle_fast.so" if fast else "libhslam_oracle.so"
    path = os.path.join(_HERE, name)
    srcs = [os.path.join(_HERE, f) for f in ("oracle_tracker.cpp", "oracle_ba.cpp", "oracle_math.hpp", "hslam_oracle.h")]
    if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs if os.path.exists(s)):
        subprocess.check_call(["make", "-C", _HERE, name], stdout=subprocess.DEVNULL)
    return path


_libs = {}


def lib(fast=False):
    if fast not in _libs:
        L = C.CDLL(build(fast))
        L.orc_pyramid_offset.restype = C.c_int64
        L.orc_ba_create.restype = C.c_void_p
        L.orc_ba_create.argtypes = [C.c_int, C.POINTER(Params)]
        _libs[fast] = L
    return _libs[fast]


# ------------------------------------------------------------------ SE3
def se3_exp(xi):
    T = np.zeros(12)
    lib().orc_se3_exp(_ptr(np.ascontiguousarray(xi, np.float64)), _ptr(T))
    return T


def se3_log(T):
    xi = np.zeros(6)
    lib().orc_se3_log(_ptr(np.ascontiguousarray(T, np.float64)), _ptr(xi))
    return xi


def se3_mul(A, B):
    out = np.zeros(12)
    lib().orc_se3_mul(_ptr(np.ascontiguousarray(A, np.float64)), _ptr(np.ascontiguousarray(B, np.float64)), _ptr(out))
    return out


def se3_inv(A):
    out = np.zeros(12)
    lib().orc_se3_inv(_ptr(np.ascontiguousarray(A, np.float64)), _ptr(out))
    return out


def se3_adj(T):
    out = np.zeros(36)
    lib().orc_se3_adj(_ptr(np.ascontiguousarray(T, np.float64)), _ptr(out))
    return out.reshape(6, 6)


def ldlt_solve(A, b):
    A = np.asfortranarray(A, np.float64)
    n = A.shape[0]
    x = np.zeros(n)
    lib().orc_ldlt_solve(A.ctypes.data_as(C.c_void_p), _ptr(np.ascontiguousarray(b, np.float64)), C.c_int(n), _ptr(x))
    return x


# ------------------------------------------------------------------ pyramid
def pyramid(img, levels, gamma=None):
    """Returns (list of (h,w,3) float32 views, list of (h,w) absgrad, flat dI buffer)."""
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    L = lib()
    tot = int(L.orc_pyramid_offset(w, h, levels))
    dI = np.zeros(tot * 3, np.float32)
    ag = np.zeros(tot, np.float32)
    g = np.ascontiguousarray(gamma, np.float32) if gamma is not None else None
    rc = L.orc_pyramid(_ptr(img), w, h, levels, _ptr(g), _ptr(dI), _ptr(ag))
    assert rc == 0
    lv, la = [], []
    for l in range(levels):
        o = int(L.orc_pyramid_offset(w, h, l))
        wl, hl = w >> l, h >> l
        lv.append(dI[3 * o: 3 * (o + wl * hl)].reshape(hl, wl, 3))
        la.append(ag[o: o + wl * hl].reshape(hl, wl))
    return lv, la, dI


# ------------------------------------------------------------------ tracker
class TrackerOracle:
    """CoarseTracker restated on the CPU (float-faithful 'f' and double 'd')."""

    def __init__(self, cam, levels, pc, ref_exposure=1.0, new_exposure=1.0, ref_aff=(0.0, 0.0), params=None, fast=False):
        self.L = lib(fast)
        self.prm = params or Params.default()
        self.levels = levels
        self.lv = (TrkLevel * levels)()
        self.L.orc_make_K(levels, cam.w, cam.h, C.c_float(cam.fx), C.c_float(cam.fy), C.c_float(cam.cx), C.c_float(cam.cy), self.lv)
        self._keep = []
        for l in range(levels):
            arrs = [np.ascontiguousarray(a, np.float32) for a in pc[l]]
            self._keep.append(arrs)
            self.lv[l].n = len(arrs[0])
            self.lv[l].pc_u, self.lv[l].pc_v, self.lv[l].pc_idepth, self.lv[l].pc_color = [_ptr(a) for a in arrs]
        self.fp = FramePair(ref_exposure, new_exposure, ref_aff[0], ref_aff[1])

    def level_K(self, l):
        v = self.lv[l]
        return dict(w=v.w, h=v.h, fx=v.fx, fy=v.fy, cx=v.cx, cy=v.cy, Ki=np.array(list(v.Ki), np.float32))

    def calc_res(self, lvl, dI_lvl, T, aff, cutoff, prec="f", want_warped=True):
        n = self.lv[lvl].n
        dt = np.float32 if prec == "f" else np.float64
        warped = [np.zeros(n + 8, dt) for _ in range(8)] if want_warped else None
        wp = (C.c_void_p * 8)(*[_ptr(a) for a in warped]) if want_warped else None
        out6 = np.zeros(6)
        nw = C.c_int(0)
        mask = np.zeros(max(n, 1), np.uint8)
        fn = self.L.orc_trk_calc_res_f if prec == "f" else self.L.orc_trk_calc_res_d
        rc = fn(self.lv, lvl, _ptr(np.ascontiguousarray(dI_lvl, np.float32)), C.byref(self.fp), C.byref(self.prm),
                _ptr(np.ascontiguousarray(T, np.float64)), C.c_double(aff[0]), C.c_double(aff[1]), C.c_float(cutoff),
                _ptr(out6), wp, C.byref(nw), _ptr(mask))
        assert rc == 0
        return out6, warped, nw.value, mask[:n]

    def calc_gs(self, lvl, warped, nw, aff, prec="f"):
        H = np.zeros(64)
        b = np.zeros(8)
        wp = (C.c_void_p * 8)(*[_ptr(a) for a in warped])
        fn = self.L.orc_trk_calc_gs_f if prec == "f" else self.L.orc_trk_calc_gs_d
        rc = fn(C.byref(self.lv[lvl]), C.byref(self.fp), C.byref(self.prm), C.c_double(aff[0]), C.c_double(aff[1]), wp,
                C.c_int(nw), _ptr(H), _ptr(b))
        assert rc == 0
        return H.reshape(8, 8).T.copy(), b  # col-major -> numpy [r,c]

    def track(self, dI_flat, T0, aff0, coarsest, min_res=None, prec="f", trace_cap=2048):
        T = np.array(T0, np.float64).copy()
        aff = np.array(aff0, np.float64).copy()
        mra = np.full(5, np.nan) if min_res is None else np.asarray(min_res, np.float64).copy()
        lastRes = np.zeros(5)
        flow = np.zeros(3)
        ok = C.c_int(0)
        its = np.zeros(5, np.int32)
        trace = np.zeros(4 * trace_cap)
        tn = C.c_int(0)
        fn = self.L.orc_trk_track_f if prec == "f" else self.L.orc_trk_track_d
        rc = fn(self.lv, self.levels, _ptr(np.ascontiguousarray(dI_flat, np.float32)), C.byref(self.fp), C.byref(self.prm),
                _ptr(T), _ptr(aff), C.c_int(coarsest), _ptr(mra), _ptr(lastRes), _ptr(flow), C.byref(ok), _ptr(its),
                _ptr(trace), C.c_int(trace_cap), C.byref(tn))
        assert rc == 0, rc
        return dict(T=T, aff=aff, lastResiduals=lastRes, flow=flow, ok=bool(ok.value), iterations=its,
                    trace=trace.reshape(-1, 4)[: min(tn.value, trace_cap)])


def make_coarse_depth(levels, w, h, su, sv, sid, sw, dI_flat):
    L = lib()
    outs = [[np.zeros((w >> l) * (h >> l), np.float32) for _ in range(4)] for l in range(levels)]
    arr = lambda k: (C.c_void_p * levels)(*[_ptr(outs[l][k]) for l in range(levels)])
    pcn = np.zeros(levels, np.int32)
    rc = L.orc_make_coarse_depth(levels, w, h, len(su), _ptr(np.ascontiguousarray(su, np.int32)),
                                 _ptr(np.ascontiguousarray(sv, np.int32)), _ptr(np.ascontiguousarray(sid, np.float32)),
                                 _ptr(np.ascontiguousarray(sw, np.float32)), _ptr(np.ascontiguousarray(dI_flat, np.float32)),
                                 arr(0), arr(1), arr(2), arr(3), _ptr(pcn))
    assert rc == 0
    return [tuple(a[: pcn[l]] for a in outs[l]) for l in range(levels)]


class TraceSettings(C.Structure):
    _fields_ = [(k, C.c_float) for k in ("maxPixSearch", "slackInterval", "stepsize", "minImprovementFactor", "GNThreshold",
                                         "extraSlackOnTH", "huberTH")] + [("minTraceTestRadius", C.c_int), ("GNIterations", C.c_int)]

    @staticmethod
    def default():   # Src/Settings.cpp:68,85-93
        return TraceSettings(0.027, 1.5, 1.0, 2.0, 0.1, 1.2, 9.0, 2, 3)


def trace_immature(dI0, ims, settings=None):
    """ImmaturePoint::traceOn over every point of `ims` (synth.ImmatureSet) against the level-0 Vec3f image dI0;
    the in/out members of `ims` are updated in place."""
    L = lib()
    S = settings or TraceSettings.default()
    h, w = dI0.shape[:2]
    a = lambda x, dt: np.ascontiguousarray(x, dt)
    K, T, A = a(ims.KRKi, np.float32), a(ims.Kt, np.float32), a(ims.aff, np.float32)
    rc = L.orc_trace_immature(_ptr(a(dI0, np.float32)), w, h, C.byref(S), len(K), _ptr(K), _ptr(T), _ptr(A), len(ims.u),
                              _ptr(a(ims.host_idx, np.int32)), _ptr(a(ims.u, np.float32)), _ptr(a(ims.v, np.float32)),
                              _ptr(a(ims.gradH, np.float32)), _ptr(a(ims.color, np.float32)), _ptr(a(ims.weights, np.float32)),
                              _ptr(a(ims.energyTH, np.float32)), _ptr(ims.idepth_min), _ptr(ims.idepth_max), _ptr(ims.quality),
                              _ptr(ims.status), _ptr(ims.lastTraceUV), _ptr(ims.lastTracePixelInterval))
    assert rc == 0
    return ims


def activate_immature(dI0_list, cam, RtAff, ims, minObs=1, minIdepthH_act=100.0, GNIts=3, huber=9.0):
    """System::optimizeImmaturePoint over every point of `ims`: (ok, idepth, res_state[n, n_frames])."""
    L = lib()
    nf, n = len(dI0_list), len(ims.u)
    imgs = [np.ascontiguousarray(d, np.float32) for d in dI0_list]
    ptrs = (C.c_void_p * nf)(*[i.ctypes.data for i in imgs])
    a = lambda x, dt: np.ascontiguousarray(x, dt)
    ok = np.zeros(n, np.int8); idp = np.zeros(n, np.float32); st = np.zeros((n, nf), np.int8)
    fxi, fyi = np.float32(1.0) / np.float32(cam.fx), np.float32(1.0) / np.float32(cam.fy)
    h, w = imgs[0].shape[:2]
    rc = L.orc_activate_immature(nf, ptrs, w, h, C.c_float(cam.fx), C.c_float(cam.fy), C.c_float(cam.cx), C.c_float(cam.cy), C.c_float(fxi),
                                 C.c_float(fyi), _ptr(a(RtAff, np.float32)), C.c_float(huber), C.c_float(minIdepthH_act), GNIts, minObs, n,
                                 _ptr(a(ims.host_idx, np.int32)), _ptr(a(ims.u, np.float32)), _ptr(a(ims.v, np.float32)),
                                 _ptr(a(ims.color, np.float32)), _ptr(a(ims.weights, np.float32)), _ptr(a(ims.energyTH, np.float32)),
                                 _ptr(a(ims.idepth_min, np.float32)), _ptr(a(ims.idepth_max, np.float32)), _ptr(ok), _ptr(idp), _ptr(st))
    assert rc == 0
    return ok, idp, st


# ------------------------------------------------------------------ window BA
def prepare_window(cam, T_eval, state, state_zero, exposure, params=None):
    n = len(T_eval)
    prm = params or Params.default()
    precalc = np.zeros(n * n * 28, np.float32)
    adH = np.zeros(n * n * 64)
    adT = np.zeros(n * n * 64)
    adHTd = np.zeros(n * n * 8, np.float32)
    rc = lib().orc_ba_prepare_window(n, _ptr(np.ascontiguousarray(T_eval, np.float64)), _ptr(np.ascontiguousarray(state, np.float64)),
                                     _ptr(np.ascontiguousarray(state_zero, np.float64)), _ptr(np.ascontiguousarray(exposure, np.float32)),
                                     C.c_float(cam.fx), C.c_float(cam.fy), C.c_float(cam.cx), C.c_float(cam.cy), C.byref(prm),
                                     _ptr(precalc), _ptr(adH), _ptr(adT), _ptr(adHTd))
    assert rc == 0
    return precalc, adH, adT, adHTd


def fill_point_soa(win):
    keep = dict(
        u=np.ascontiguousarray(win.u, np.float32), v=np.ascontiguousarray(win.v, np.float32),
        idepth=np.ascontiguousarray(win.idepth, np.float32), idepth_zero=np.ascontiguousarray(win.idepth_zero, np.float32),
        color=np.ascontiguousarray(win.color, np.float32), weights=np.ascontiguousarray(win.weights, np.float32),
        priorF=np.ascontiguousarray(win.priorF, np.float32), deltaF=np.ascontiguousarray(win.deltaF, np.float32),
        host_idx=np.ascontiguousarray(win.host_idx, np.int32), res_begin=np.ascontiguousarray(win.res_begin, np.int32),
        res_index=np.ascontiguousarray(win.res_index, np.int32))
    s = PointSoA(len(keep["u"]))
    for k, a in keep.items():
        setattr(s, k, _ptr(a))
    return s, keep


def fill_residual_soa(win):
    keep = dict(
        point_idx=np.ascontiguousarray(win.point_idx, np.int32), target_idx=np.ascontiguousarray(win.target_idx, np.int32),
        state_state=np.ascontiguousarray(win.state_state, np.int8), is_linearized=np.ascontiguousarray(win.is_linearized, np.uint8),
        is_active=np.ascontiguousarray(win.is_active, np.uint8))
    s = ResidualSoA(len(keep["point_idx"]))
    for k, a in keep.items():
        setattr(s, k, _ptr(a))
    s.state_energy = None
    return s, keep


class BAOracle:
    """EnergyFunctional + Accumulated*Hessian + linearize restated on the CPU."""

    def __init__(self, win, dI_level0_list, prec="f", threads=1, params=None, fast=False):
        self.L = lib(fast)
        self.prm = params or Params.default()
        self.h = C.c_void_p(self.L.orc_ba_create(1 if prec == "d" else 0, C.byref(self.prm)))
        self.win = win
        self.n = win.n
        self._keep = []
        self.set_window(win, dI_level0_list)
        ps, k1 = fill_point_soa(win)
        rs, k2 = fill_residual_soa(win)
        self._keep += [k1, k2]
        assert self.L.orc_ba_set_topology(self.h, C.byref(ps), C.byref(rs)) == 0
        self.L.orc_ba_set_threads(self.h, C.c_int(threads))
        self.P, self.R = win.P, win.R

    def __del__(self):
        try:
            self.L.orc_ba_destroy(self.h)
        except Exception:
            pass

    def set_window(self, win, dI_list=None):
        if dI_list is not None:
            self._dI = [np.ascontiguousarray(d, np.float32) for d in dI_list]
        cam = win.cam
        ow = OrcWindow()
        ow.n_frames, ow.w, ow.h = win.n, cam.w, cam.h
        ow.fx, ow.fy, ow.cx, ow.cy = cam.fx, cam.fy, cam.cx, cam.cy
        fxf, fyf = np.float32(cam.fx), np.float32(cam.fy)
        ow.fxi, ow.fyi = np.float32(1.0) / fxf, np.float32(1.0) / fyf
        self._dIp = (C.c_void_p * win.n)(*[_ptr(d) for d in self._dI])
        ow.dI = C.cast(self._dIp, C.c_void_p)
        ka = dict(frameEnergyTH=np.ascontiguousarray(win.frameEnergyTH, np.float32), precalc=np.ascontiguousarray(win.precalc, np.float32),
                  adHost=np.ascontiguousarray(win.adHost, np.float64), adTarget=np.ascontiguousarray(win.adTarget, np.float64),
                  adHTdeltaF=np.ascontiguousarray(win.adHTdeltaF, np.float32), framePrior=np.ascontiguousarray(win.framePrior, np.float64),
                  frameDeltaPrior=np.ascontiguousarray(win.frameDeltaPrior, np.float64))
        for k, a in ka.items():
            setattr(ow, k, _ptr(a))
        for i in range(4):
            ow.cDeltaF[i] = float(win.cDeltaF[i])
            ow.cPrior[i] = float(win.cPrior[i])
        self._keep.append(ka)
        assert self.L.orc_ba_set_window(self.h, C.byref(ow)) == 0

    def update_points(self, idepth=None, deltaF=None, idepth_zero=None):
        a = [None if x is None else np.ascontiguousarray(x, np.float32) for x in (idepth, deltaF, idepth_zero)]
        assert self.L.orc_ba_update_points(self.h, _ptr(a[0]), _ptr(a[1]), _ptr(a[2])) == 0

    def linearize(self, fix=False):
        R = self.R
        e = C.c_double(0)
        ns = np.zeros(R, np.int8); ne = np.zeros(R, np.float32); nwo = np.zeros(R, np.float32)
        cpt = np.zeros(R * 3, np.float32); pto = np.zeros(R * 16, np.float32)
        assert self.L.orc_ba_linearize(self.h, C.c_int(int(fix)), C.byref(e), _ptr(ns), _ptr(ne), _ptr(nwo), _ptr(cpt), _ptr(pto)) == 0
        return dict(energy=e.value, newState=ns, newEnergy=ne, newEnergyWithOutlier=nwo, centerProjectedTo=cpt.reshape(R, 3),
                    projectedTo=pto.reshape(R, 8, 2))

    def apply_res(self):
        assert self.L.orc_ba_apply_res(self.h) == 0

    def fix_linearization(self, ids):
        ids = np.ascontiguousarray(ids, np.int32)
        assert self.L.orc_ba_fix_linearization(self.h, _ptr(ids), C.c_int(len(ids))) == 0

    def accumulate(self, which):
        N = 8 * self.n + 4
        H = np.zeros(N * N); b = np.zeros(N); nres = C.c_int(0)
        assert self.L.orc_ba_accumulate(self.h, C.c_int(which), _ptr(H), _ptr(b), C.byref(nres)) == 0
        return H.reshape(N, N).T.copy(), b, nres.value

    def point_results(self):
        P = self.P
        names = ["HdiF", "bdSumF", "idepth_hessian", "Hdd_accAF", "bd_accAF", "Hcd_accAF", "Hdd_accLF", "bd_accLF", "Hcd_accLF"]
        arrs = [np.zeros(P * (4 if "Hcd" in k else 1)) for k in names]
        assert self.L.orc_ba_point_results(self.h, *[_ptr(a) for a in arrs]) == 0
        d = dict(zip(names, arrs))
        d["Hcd_accAF"] = d["Hcd_accAF"].reshape(P, 4); d["Hcd_accLF"] = d["Hcd_accLF"].reshape(P, 4)
        return d

    def residual_results(self):
        R = self.R
        st = np.zeros(R, np.int8); act = np.zeros(R, np.uint8); jp = np.zeros(R * 8); j74 = np.zeros(R * 74); rtz = np.zeros(R * 8)
        assert self.L.orc_ba_residual_results(self.h, _ptr(st), _ptr(act), _ptr(jp), _ptr(j74), _ptr(rtz)) == 0
        return dict(state_state=st, is_active=act, JpJdF=jp.reshape(R, 8), J=j74.reshape(R, 74), res_toZeroF=rtz.reshape(R, 8))

    def raw_top(self, which):
        n = self.n
        out = np.zeros(n * n * 169)
        assert self.L.orc_ba_raw_top(self.h, C.c_int(which), _ptr(out)) == 0
        return out.reshape(n * n, 13, 13)

    def raw_sc(self):
        n = self.n
        D = np.zeros(n**3 * 64); E = np.zeros(n * n * 32); EB = np.zeros(n * n * 8); Hcc = np.zeros(16); bc = np.zeros(4)
        assert self.L.orc_ba_raw_sc(self.h, _ptr(D), _ptr(E), _ptr(EB), _ptr(Hcc), _ptr(bc)) == 0
        return dict(D=D.reshape(n**3, 8, 8).transpose(0, 2, 1), E=E.reshape(n * n, 4, 8).transpose(0, 2, 1), EB=EB.reshape(n * n, 8),
                    Hcc=Hcc.reshape(4, 4).T, bc=bc)

    def marginalize_points(self, point_ids, prior_scale):
        N = 8 * self.n + 4
        M = np.zeros(N * N); Mb = np.zeros(N); Ms = np.zeros(N * N); Mbs = np.zeros(N); nres = C.c_int(0)
        ids = np.ascontiguousarray(point_ids, np.int32)
        assert self.L.orc_ba_marginalize_points(self.h, _ptr(ids), C.c_int(len(ids)), C.c_float(prior_scale), _ptr(M), _ptr(Mb),
                                                _ptr(Ms), _ptr(Mbs), C.byref(nres)) == 0
        return M.reshape(N, N).T.copy(), Mb, Ms.reshape(N, N).T.copy(), Mbs, nres.value

    def calc_l_energy(self):
        e = C.c_double(0)
        assert self.L.orc_ba_calc_l_energy(self.h, C.byref(e)) == 0
        return e.value

    def new_frame_energy_th(self, newest):
        """System::setNewFrameEnergyTH (FullSystemOptimize.cpp:60-101); also stored as the window's frameEnergyTH[newest]."""
        th = C.c_float(0)
        assert self.L.orc_ba_new_frame_energy_th(self.h, C.c_int(int(newest)), C.byref(th)) == 0
        return np.float32(th.value)

    def resubstitute(self, x):
        step = np.zeros(self.P, np.float32)
        assert self.L.orc_ba_resubstitute(self.h, _ptr(np.ascontiguousarray(x, np.float64)), _ptr(step)) == 0
        return step
