"""ctypes front end of oracle/_ref/libhslam_ref.so: the UNMODIFIED reference translation units of the BA hot path
(Src/OptimizationClasses.cpp, Src/AccumulatedTopHessian.cpp, Src/AccumulatedSCHessian.cpp, Src/Settings.cpp + their headers)
compiled against the stand-ins of oracle/ref_shim/ with oracle/ref_driver.cpp around them (see oracle/Makefile, target
_ref).  TEST INFRASTRUCTURE: it pins the oracle (tests/test_oracle_vs_reference.py); nothing of the product loads it."""
import ctypes as C
import os

import numpy as np

from . import oraclelib as O

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libhslam_ref.so")


def available():
    return os.path.exists(LIB)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        L.ref_ba_create.restype = C.c_void_p
        L.ref_ba_destroy.argtypes = [C.c_void_p]
        _lib = L
    return _lib


_ptr = O._ptr


class RefBA:
    """Same call surface as oraclelib.BAOracle, answered by the reference's own code."""

    def __init__(self, win, dI_level0_list):
        self.L = lib()
        self.h = C.c_void_p(self.L.ref_ba_create())
        self.win, self.n = win, win.n
        self._keep = []
        self.set_window(win, dI_level0_list)
        ps, k1 = O.fill_point_soa(win)
        rs, k2 = O.fill_residual_soa(win)
        self._keep += [k1, k2]
        assert self.L.ref_ba_set_topology(self.h, C.byref(ps), C.byref(rs)) == 0
        self.P, self.R = win.P, win.R

    def __del__(self):
        try:
            self.L.ref_ba_destroy(self.h)
        except Exception:
            pass

    def set_window(self, win, dI_list=None):
        if dI_list is not None:
            self._dI = [np.ascontiguousarray(d, np.float32) for d in dI_list]
        cam = win.cam
        ow = O.OrcWindow()
        ow.n_frames, ow.w, ow.h = win.n, cam.w, cam.h
        ow.fx, ow.fy, ow.cx, ow.cy = cam.fx, cam.fy, cam.cx, cam.cy
        ow.fxi, ow.fyi = np.float32(1.0) / np.float32(cam.fx), np.float32(1.0) / np.float32(cam.fy)
        self._dIp = (C.c_void_p * win.n)(*[_ptr(d) for d in self._dI])
        ow.dI = C.cast(self._dIp, C.c_void_p)
        ka = dict(frameEnergyTH=np.ascontiguousarray(win.frameEnergyTH, np.float32), precalc=np.ascontiguousarray(win.precalc, np.float32),
                  adHost=np.ascontiguousarray(win.adHost, np.float64), adTarget=np.ascontiguousarray(win.adTarget, np.float64),
                  adHTdeltaF=np.ascontiguousarray(win.adHTdeltaF, np.float32), framePrior=np.ascontiguousarray(win.framePrior, np.float64),
                  frameDeltaPrior=np.ascontiguousarray(win.frameDeltaPrior, np.float64))
        for k, a in ka.items():
            setattr(ow, k, _ptr(a))
        for i in range(4):
            ow.cDeltaF[i] = float(win.cDeltaF[i]); ow.cPrior[i] = float(win.cPrior[i])
        self._keep.append(ka)
        assert self.L.ref_ba_set_window(self.h, C.byref(ow)) == 0

    def update_points(self, idepth=None, deltaF=None, idepth_zero=None):
        a = [None if x is None else np.ascontiguousarray(x, np.float32) for x in (idepth, deltaF, idepth_zero)]
        assert self.L.ref_ba_update_points(self.h, _ptr(a[0]), _ptr(a[1]), _ptr(a[2])) == 0

    def linearize(self, fix=False):
        R = self.R
        e = C.c_double(0)
        ns = np.zeros(R, np.int8); ne = np.zeros(R, np.float32); nwo = np.zeros(R, np.float32)
        cpt = np.zeros(R * 3, np.float32); pto = np.zeros(R * 16, np.float32)
        assert self.L.ref_ba_linearize(self.h, C.c_int(int(fix)), C.byref(e), _ptr(ns), _ptr(ne), _ptr(nwo), _ptr(cpt), _ptr(pto)) == 0
        return dict(energy=e.value, newState=ns, newEnergy=ne, newEnergyWithOutlier=nwo, centerProjectedTo=cpt.reshape(R, 3),
                    projectedTo=pto.reshape(R, 8, 2))

    def apply_res(self):
        assert self.L.ref_ba_apply_res(self.h) == 0

    def fix_linearization(self, ids):
        ids = np.ascontiguousarray(ids, np.int32)
        assert self.L.ref_ba_fix_linearization(self.h, _ptr(ids), C.c_int(len(ids))) == 0

    def accumulate(self, which):
        N = 8 * self.n + 4
        H = np.zeros(N * N); b = np.zeros(N); nres = C.c_int(0)
        assert self.L.ref_ba_accumulate(self.h, C.c_int(which), _ptr(H), _ptr(b), C.byref(nres)) == 0
        return H.reshape(N, N).T.copy(), b, nres.value

    def marginalize_points(self, point_ids, prior_scale):
        """EnergyFunctional::marginalizePointsF's accumulate + stitch part: addPoint<2>, SC addPoint(p, false), stitchDouble."""
        N = 8 * self.n + 4
        M = np.zeros(N * N); Mb = np.zeros(N); Ms = np.zeros(N * N); Mbs = np.zeros(N); nres = C.c_int(0)
        ids = np.ascontiguousarray(point_ids, np.int32)
        assert self.L.ref_ba_marginalize_points(self.h, _ptr(ids), C.c_int(len(ids)), C.c_float(prior_scale), _ptr(M), _ptr(Mb),
                                                _ptr(Ms), _ptr(Mbs), C.byref(nres)) == 0
        return M.reshape(N, N).T.copy(), Mb, Ms.reshape(N, N).T.copy(), Mbs, nres.value

    def system_linearize_all(self):
        """System::linearizeAll(false) of the reference (linearizeAll_Reductor + setNewFrameEnergyTH): (energy, new
        frameEnergyTH of the newest frame).  Per-residual results: residual_results() / linearize() fields as usual."""
        e = C.c_double(0); th = C.c_float(0)
        assert self.L.ref_ba_system_linearize_all(self.h, C.byref(e), C.byref(th)) == 0
        return e.value, np.float32(th.value)

    def resubstitute(self, x):
        """EnergyFunctional::resubstituteF_MT (single-thread branch) of the reference: per-point steps."""
        x = np.ascontiguousarray(x, np.float64)
        step = np.zeros(self.P, np.float32)
        assert self.L.ref_ba_resubstitute(self.h, _ptr(x), _ptr(step)) == 0
        return step

    def calc_l_energy(self):
        """EnergyFunctional::calcLEnergyF_MT of the reference (its own IndexThreadReduce, one worker)."""
        e = C.c_double(0)
        assert self.L.ref_ba_calc_l_energy(self.h, C.byref(e)) == 0
        return e.value

    def point_results(self):
        P = self.P
        names = ["HdiF", "bdSumF", "idepth_hessian", "Hdd_accAF", "bd_accAF", "Hcd_accAF", "Hdd_accLF", "bd_accLF", "Hcd_accLF"]
        arrs = [np.zeros(P * (4 if "Hcd" in k else 1)) for k in names]
        assert self.L.ref_ba_point_results(self.h, *[_ptr(a) for a in arrs]) == 0
        d = dict(zip(names, arrs))
        d["Hcd_accAF"] = d["Hcd_accAF"].reshape(P, 4); d["Hcd_accLF"] = d["Hcd_accLF"].reshape(P, 4)
        return d

    def residual_results(self):
        R = self.R
        st = np.zeros(R, np.int8); act = np.zeros(R, np.uint8); jp = np.zeros(R * 8); j74 = np.zeros(R * 74); rtz = np.zeros(R * 8)
        assert self.L.ref_ba_residual_results(self.h, _ptr(st), _ptr(act), _ptr(jp), _ptr(j74), _ptr(rtz)) == 0
        return dict(state_state=st, is_active=act, JpJdF=jp.reshape(R, 8), J=j74.reshape(R, 74), res_toZeroF=rtz.reshape(R, 8))


def precalc_set(cam, T_eval, state, state_zero, exposure):
    """FrameFramePrecalc::set (Src/OptimizationClasses.cpp:13-39) of every ordered pair -> n*n*28 floats."""
    n = len(T_eval)
    out = np.zeros(n * n * 28, np.float32)
    rc = lib().ref_precalc_set(C.c_int(n), _ptr(np.ascontiguousarray(T_eval, np.float64)), _ptr(np.ascontiguousarray(state, np.float64)),
                               _ptr(np.ascontiguousarray(state_zero, np.float64)), _ptr(np.ascontiguousarray(exposure, np.float32)),
                               C.c_float(cam.fx), C.c_float(cam.fy), C.c_float(cam.cx), C.c_float(cam.cy), C.c_int(cam.w), C.c_int(cam.h),
                               _ptr(out))
    assert rc == 0
    return out


# ------------------------------------------------------------------ Stage T + tracing (oracle/_ref/libhslam_ref_trk.so)
LIB_TRK = os.path.join(HERE, "_ref", "libhslam_ref_trk.so")
_lib_trk = None


def adjoints_deltas(T_eval, state, state_zero, exposure):
    """EnergyFunctional::setAdjointsF + setDeltaF of the reference: (adHost, adTarget, adHTdeltaF) flat like O.prepare_window."""
    n = len(T_eval)
    adH = np.zeros(n * n * 64); adT = np.zeros(n * n * 64); adHTd = np.zeros(n * n * 8, np.float32)
    a = lambda x, dt: np.ascontiguousarray(x, dt)
    rc = lib().ref_ef_adjoints_deltas(C.c_int(n), _ptr(a(T_eval, np.float64)), _ptr(a(state, np.float64)), _ptr(a(state_zero, np.float64)),
                                      _ptr(a(exposure, np.float32)), _ptr(adH), _ptr(adT), _ptr(adHTd))
    assert rc == 0
    return adH, adT, adHTd


def available_trk():
    return os.path.exists(LIB_TRK)


def lib_trk():
    global _lib_trk
    if _lib_trk is None:
        L = C.CDLL(LIB_TRK)
        L.ref_trk_create.restype = C.c_void_p
        L.ref_trk_destroy.argtypes = [C.c_void_p]
        _lib_trk = L
    return _lib_trk


class RefTracker:
    """CoarseTracker of the reference (Src/CoarseTracker.cpp, unmodified): makeK, calcRes, calcGSSSE, trackNewestCoarse."""

    def __init__(self, cam, levels, pc, dI_new_flat, ref_exposure=1.0, new_exposure=1.0, ref_aff=(0.0, 0.0)):
        self.L = lib_trk()
        self.levels = levels
        self.h = C.c_void_p(self.L.ref_trk_create(C.c_int(cam.w), C.c_int(cam.h), C.c_int(levels), C.c_float(cam.fx), C.c_float(cam.fy),
                                                  C.c_float(cam.cx), C.c_float(cam.cy)))
        self.n = []
        for l in range(levels):
            a = [np.ascontiguousarray(x, np.float32) for x in pc[l]]
            assert self.L.ref_trk_set_ref(self.h, C.c_int(l), _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]), C.c_int(len(a[0]))) == 0
            self.n.append(len(a[0]))
        self._dI = np.ascontiguousarray(dI_new_flat, np.float32)
        assert self.L.ref_trk_set_frames(self.h, C.c_float(ref_exposure), C.c_float(new_exposure), C.c_double(ref_aff[0]),
                                         C.c_double(ref_aff[1]), _ptr(self._dI)) == 0

    def __del__(self):
        try:
            self.L.ref_trk_destroy(self.h)
        except Exception:
            pass

    def level_K(self, l):
        k = np.zeros(4, np.float32); wh = np.zeros(2, np.int32); K9 = np.zeros(9, np.float32); Ki9 = np.zeros(9, np.float32)
        assert self.L.ref_trk_get_K(self.h, C.c_int(l), _ptr(k), _ptr(wh), _ptr(K9), _ptr(Ki9)) == 0
        return dict(w=int(wh[0]), h=int(wh[1]), fx=k[0], fy=k[1], cx=k[2], cy=k[3], Ki=Ki9)

    def calc_res(self, lvl, T, aff, cutoff):
        n = self.n[lvl]
        warped = [np.zeros(n + 8, np.float32) for _ in range(8)]
        wp = (C.c_void_p * 8)(*[_ptr(a) for a in warped])
        out6 = np.zeros(6); nw = C.c_int(0)
        assert self.L.ref_trk_calc_res(self.h, C.c_int(lvl), _ptr(np.ascontiguousarray(T, np.float64)), C.c_double(aff[0]), C.c_double(aff[1]),
                                       C.c_float(cutoff), _ptr(out6), wp, C.byref(nw)) == 0
        return out6, warped, nw.value

    def calc_gs(self, lvl, T, aff):
        H = np.zeros(64); b = np.zeros(8)
        assert self.L.ref_trk_calc_gs(self.h, C.c_int(lvl), _ptr(np.ascontiguousarray(T, np.float64)), C.c_double(aff[0]), C.c_double(aff[1]),
                                      _ptr(H), _ptr(b)) == 0
        return H.reshape(8, 8).T.copy(), b

    def make_coarse_depth(self, cpt3, hdif, dI_ref_flat):
        """CoarseTracker::setCoarseTrackingRef / makeCoarseDepthL0 of the reference; returns the per-level clouds."""
        c = np.ascontiguousarray(cpt3, np.float32).reshape(-1, 3); hd = np.ascontiguousarray(hdif, np.float32)
        self._dIref = np.ascontiguousarray(dI_ref_flat, np.float32)
        assert self.L.ref_trk_make_coarse_depth(self.h, C.c_int(len(hd)), _ptr(c), _ptr(hd), _ptr(self._dIref)) == 0
        out = []
        for l in range(self.levels):
            n = C.c_int(0)
            self.L.ref_trk_get_ref(self.h, C.c_int(l), None, None, None, None, C.byref(n))
            a = [np.zeros(n.value, np.float32) for _ in range(4)]
            self.L.ref_trk_get_ref(self.h, C.c_int(l), _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]), C.byref(n))
            out.append(tuple(a))
        return out

    def track(self, T0, aff0, coarsest, min_res=None):
        T = np.array(T0, np.float64).copy(); aff = np.array(aff0, np.float64).copy()
        mra = np.full(5, np.nan) if min_res is None else np.asarray(min_res, np.float64).copy()
        lastRes = np.zeros(5); flow = np.zeros(3); ok = C.c_int(0)
        assert self.L.ref_trk_track(self.h, _ptr(T), _ptr(aff), C.c_int(coarsest), _ptr(mra), _ptr(lastRes), _ptr(flow), C.byref(ok)) == 0
        return dict(T=T, aff=aff, lastResiduals=lastRes, flow=flow, ok=bool(ok.value))


def trace_immature(dI0, cam, ims):
    """ImmaturePoint::traceOn of the reference for every point of `ims` (synth.ImmatureSet), in place like O.trace_immature."""
    h, w = dI0.shape[0], dI0.shape[1]
    a = lambda x, dt: np.ascontiguousarray(x, dt)
    K, T, A = a(ims.KRKi, np.float32), a(ims.Kt, np.float32), a(ims.aff, np.float32)
    rc = lib_trk().ref_trace_immature(_ptr(a(dI0, np.float32)), C.c_int(w), C.c_int(h), C.c_float(cam.fx), C.c_float(cam.fy), C.c_float(cam.cx),
                                      C.c_float(cam.cy), C.c_int(len(K)), _ptr(K), _ptr(T), _ptr(A), C.c_int(len(ims.u)),
                                      _ptr(a(ims.host_idx, np.int32)), _ptr(a(ims.u, np.float32)), _ptr(a(ims.v, np.float32)),
                                      _ptr(a(ims.gradH, np.float32)), _ptr(a(ims.color, np.float32)), _ptr(a(ims.weights, np.float32)),
                                      _ptr(a(ims.energyTH, np.float32)), _ptr(ims.idepth_min), _ptr(ims.idepth_max), _ptr(ims.quality),
                                      _ptr(ims.status), _ptr(ims.lastTraceUV), _ptr(ims.lastTracePixelInterval))
    assert rc == 0
    return ims


def activate_immature(dI0_list, cam, RtAff, ims, minObs=1, minIdepthH_act=100.0, GNIts=3, huber=9.0):
    """System::optimizeImmaturePoint's loop around the reference's own ImmaturePoint::linearizeResidual for every point of
    `ims`: (ok, idepth, res_state[n, n_frames]) like O.activate_immature."""
    nf, n = len(dI0_list), len(ims.u)
    imgs = [np.ascontiguousarray(d, np.float32) for d in dI0_list]
    ptrs = (C.c_void_p * nf)(*[i.ctypes.data for i in imgs])
    a = lambda x, dt: np.ascontiguousarray(x, dt)
    ok = np.zeros(n, np.int8); idp = np.zeros(n, np.float32); st = np.zeros((n, nf), np.int8)
    h, w = imgs[0].shape[:2]
    rc = lib_trk().ref_activate_immature(C.c_int(nf), ptrs, C.c_int(w), C.c_int(h), C.c_float(cam.fx), C.c_float(cam.fy), C.c_float(cam.cx),
                                         C.c_float(cam.cy), _ptr(a(RtAff, np.float32)), C.c_float(huber), C.c_float(minIdepthH_act),
                                         C.c_int(GNIts), C.c_int(minObs), C.c_int(n), _ptr(a(ims.host_idx, np.int32)),
                                         _ptr(a(ims.u, np.float32)), _ptr(a(ims.v, np.float32)), _ptr(a(ims.color, np.float32)),
                                         _ptr(a(ims.weights, np.float32)), _ptr(a(ims.energyTH, np.float32)),
                                         _ptr(a(ims.idepth_min, np.float32)), _ptr(a(ims.idepth_max, np.float32)), _ptr(ok), _ptr(idp), _ptr(st))
    assert rc == 0
    return ok, idp, st


def available_frame():
    return os.path.exists(os.path.join(HERE, "_ref", "libhslam_ref_frame.so"))


def create_dir_pyrs(img, levels):
    """Frame::CreateDirPyrs of the reference (run by its Frame constructor) on a float image:
    (list of (h,w,3) {I,dx,dy}, list of (h,w) absSquaredGrad).  The first and the last row of dx / dy / absSquaredGrad are
    never written by the reference (uninitialised memory): compare rows 1 .. h-2 only."""
    L = C.CDLL(os.path.join(HERE, "_ref", "libhslam_ref_frame.so"))
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    n = sum((w >> l) * (h >> l) for l in range(levels))
    dI = np.zeros(3 * n, np.float32); ab = np.zeros(n, np.float32)
    assert L.ref_frame_create_dir_pyrs(_ptr(img), C.c_int(w), C.c_int(h), C.c_int(levels), _ptr(dI), _ptr(ab)) == 0
    lv, la, o = [], [], 0
    for l in range(levels):
        wl, hl = w >> l, h >> l
        lv.append(dI[3 * o:3 * (o + wl * hl)].reshape(hl, wl, 3)); la.append(ab[o:o + wl * hl].reshape(hl, wl))
        o += wl * hl
    return lv, la

