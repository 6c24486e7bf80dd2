"""Both stages of the hot path inside ONE sequence loop (BASELINE.json configs[2]): frame-to-keyframe tracking AND the
windowed photometric bundle adjustment of every new keyframe.

Host-side driver restated from the reference's control flow — System::trackNewCoarse (Src/System.cpp:333-489), System::
AddKeyframe (Src/Mapping.cpp:12-142), System::optimize / linearizeAll / setNewFrameEnergyTH / doStepFromBackup (Src/
FullSystemOptimize.cpp:60-165,171-264,362-546), EnergyFunctional::solveSystemF / resubstituteF_MT (Src/EnergyFunctional.cpp:
222-274,705-817) and CoarseTracker::makeCoarseDepthL0 (Src/CoarseTracker.cpp:105-263) — with the hot-path calls going
through a BACK END: the CUDA library behind the C ABI (CudaBackend), or the CPU oracle (tests only).  Both back ends run
the SAME host code and take the same decisions, so their trajectories can be compared (north_star: "trajectories on the
synthetic sequence agree in ATE within 1 %").

Per incoming frame: trackNewestCoarse against the newest keyframe (constant-motion hypothesis first, then "no motion").
Per keyframe (every `kf_every` frames): window of <= `max_kf` keyframes; new points in the new keyframe; a residual for
every (point, other keyframe); <= `max_iters` Gauss-Newton iterations of
    setPrecalcValues -> linearizeAll(false) + setNewFrameEnergyTH -> applyRes -> accumulateAF/LF/SCF -> solve (host, fp64,
    (8n+4)^2, FIX_LAMBDA 1e-5) -> resubstituteF -> doStepFromBackup (always accepted: setting_forceAceptStep),
then linearizeAll(true), setEvalPT of the newest frame and makeCoarseDepthL0 from the residuals into it — the new
tracking reference.

Out of scope here, replaced by documented stand-ins (they are host-side policy / algebra):
  * immature points + activation: new points start from the rendered depth with 5 % noise;
  * marginalisation (HM, bM): the oldest keyframe is dropped, and the TWO oldest keyframes of the window carry the strong
    first-frame prior (Frame.h:229-256) at their latest estimate, so that gauge and scale stay fixed;
  * nullspace orthogonalisation of x; the calibration is held by its prior (setting_initialCalibHessian).
"""
from __future__ import annotations

import dataclasses
import time

import numpy as np

from . import synth
from .sequence import trajectory, render_sequence  # noqa: F401  (re-exported)

RES_IN, RES_OOB, RES_OUT = 0, 1, 2


@dataclasses.dataclass
class KeyFrame:
    seq_idx: int
    img: np.ndarray
    handle: object               # back end's frame object
    T_eval: np.ndarray           # worldToCam_evalPT (12)
    state: np.ndarray            # 10
    state_zero: np.ndarray       # 10
    frameEnergyTH: float
    # points hosted here
    u: np.ndarray
    v: np.ndarray
    idepth: np.ndarray
    color: np.ndarray
    weights: np.ndarray

    def T_now(self):
        s = self.state
        sc = np.concatenate([synth.SCALE_XI_TRANS * s[0:3], synth.SCALE_XI_ROT * s[3:6]])
        R, t = synth.se3_exp(sc)
        return synth.se3_mul(synth.T12(R, t), self.T_eval)

    def aff_now(self):
        return (synth.SCALE_A * self.state[6], synth.SCALE_B * self.state[7])


@dataclasses.dataclass
class SequenceBAResult:
    T_w2c_est: list
    T_w2c_true: list
    ok: list
    ate_rmse: float
    n_keyframes: int
    ba_iterations: list          # GN iterations per keyframe
    ba_states: list              # per keyframe, per iteration: newState array (int8) — for back-end comparisons
    ba_energy: list              # per keyframe, per iteration: energy
    thresholds: list             # per keyframe, per iteration: frameEnergyTH of the newest frame
    kf_poses_after_ba: list      # per keyframe: world->camera of the newest keyframe after its optimisation
    wall_s: dict


def select_points(img, depth, n_points, rng, idepth_noise=0.05):
    """New points of a keyframe (stand-in for immature-point activation): high-gradient pixels, inverse depth from the
    rendered depth with noise, the 8-pixel pattern's colours and gradient weights (ImmaturePoint.cpp:27-38)."""
    pyr, ag = synth.build_pyramid_np(img, 1)
    h, w = img.shape
    ys, xs = np.mgrid[6:h - 6, 6:w - 6]
    ys, xs = ys.reshape(-1), xs.reshape(-1)
    good = (ag[0][ys, xs] > 40) & np.isfinite(depth[ys, xs])
    ys, xs = ys[good], xs[good]
    sel = rng.choice(len(xs), size=min(n_points, len(xs)), replace=False)
    sel.sort()
    x, y = xs[sel], ys[sel]
    idv = ((1.0 / depth[y, x]) * (1.0 + rng.normal(0, idepth_noise, len(sel)))).astype(np.float32)
    P = synth.PATTERN
    c = np.stack([pyr[0][y + P[k, 1], x + P[k, 0], 0] for k in range(8)], 1).astype(np.float32)
    g2 = np.stack([pyr[0][y + P[k, 1], x + P[k, 0], 1] ** 2 + pyr[0][y + P[k, 1], x + P[k, 0], 2] ** 2 for k in range(8)], 1).astype(np.float32)
    c_th = np.float32(50.0 * 50.0)
    wv = np.sqrt(c_th / (c_th + g2)).astype(np.float32)
    return x.astype(np.float32), y.astype(np.float32), idv, c, wv


def build_window(cam, kfs):
    """synth.Window of the current keyframes (oldest first): tables from the frame states, every point of every keyframe,
    one residual per (point, other keyframe)."""
    n = len(kfs)
    T_eval = [k.T_eval for k in kfs]
    state = np.array([k.state for k in kfs]); state_zero = np.array([k.state_zero for k in kfs])
    exposure = np.ones(n, np.float32)
    precalc, adH, adT, adHTd = synth.prepare_window_tables(cam, T_eval, state, state_zero, exposure)
    u = np.concatenate([k.u for k in kfs]); v = np.concatenate([k.v for k in kfs])
    idepth = np.concatenate([k.idepth for k in kfs])
    color = np.concatenate([k.color for k in kfs]); weights = np.concatenate([k.weights for k in kfs])
    host_idx = np.concatenate([np.full(len(k.u), i, np.int32) for i, k in enumerate(kfs)])
    P = len(u)
    others = np.arange(n - 1)[None, :] + (np.arange(n - 1)[None, :] >= host_idx[:, None])   # P x (n-1), skips the host
    R_ = P * (n - 1)
    prior = synth._frame_priors(n).reshape(n, 8)
    if n > 1:
        prior[1] = prior[0]            # the two oldest keyframes hold gauge and scale (stand-in for the marginalisation prior)
    return synth.Window(
        cam=cam, n=n, images=[k.img for k in kfs], T_w2c_eval=np.array(T_eval), state=state, state_zero=state_zero,
        ab_exposure=exposure, frameEnergyTH=np.array([k.frameEnergyTH for k in kfs], np.float32), precalc=precalc, adHost=adH,
        adTarget=adT, adHTdeltaF=adHTd, cDeltaF=np.zeros(4, np.float32), cPrior=np.full(4, 5e9), framePrior=prior.reshape(-1),
        frameDeltaPrior=state[:, :8].reshape(-1).copy(),   # delta_prior = state - getPriorZero() = state (Frame.h:259-266)
        u=u, v=v, idepth=idepth.copy(), idepth_zero=idepth.copy(),
        color=color, weights=weights, priorF=np.zeros(P, np.float32), deltaF=np.zeros(P, np.float32), host_idx=host_idx,
        res_begin=(np.arange(P + 1, dtype=np.int32) * (n - 1)), res_index=np.arange(R_, dtype=np.int32),
        point_idx=np.repeat(np.arange(P, dtype=np.int32), n - 1), target_idx=others.reshape(-1).astype(np.int32),
        state_state=np.zeros(R_, np.int8), is_linearized=np.zeros(R_, np.uint8), is_active=np.zeros(R_, np.uint8))


def solve_system(sys3, n, lam=1e-5):
    """EnergyFunctional::solveSystemF with setting_solverMode = FIX_LAMBDA (EnergyFunctional.cpp:705-802), HM = 0."""
    (HA, bA), (HL, bL), (HS, bS) = sys3
    HF = HL + HA
    bF = bL + bA - bS
    d = np.arange(8 * n + 4)
    HF[d, d] *= (1 + lam)
    HF = HF - HS * (1.0 / (1 + lam))
    sv = 1.0 / np.sqrt(np.diag(HF) + 10.0)
    Hs = sv[:, None] * HF * sv[None, :]
    return sv * np.linalg.solve(Hs, sv * bF)


class _Timed:
    """Forwards to the back end's BA object and accumulates the wall time spent inside it (the hot-path share of a
    keyframe, as opposed to the host code both back ends share)."""

    def __init__(self, ba, wall):
        self._ba, self._wall = ba, wall
        self.win = ba.win

    def __getattr__(self, name):
        fn = getattr(self._ba, name)
        if not callable(fn):
            return fn

        def call(*a, **k):
            t0 = time.perf_counter()
            try:
                return fn(*a, **k)
            finally:
                self._wall["ba_backend"] += time.perf_counter() - t0
        return call


def optimize_window(cam, kfs, ba, max_iters, log):
    """System::optimize (FullSystemOptimize.cpp:362-546) over the current window through the back end's BA object."""
    n = len(kfs)
    win = ba.win
    newest = n - 1
    states_log, energy_log, th_log = [], [], []

    def linearize(fix):
        lin = ba.linearize(fix)                       # linearizeAll (+ applyRes when fix)
        th = ba.new_frame_energy_th(newest)           # setNewFrameEnergyTH (FullSystemOptimize.cpp:124)
        kfs[newest].frameEnergyTH = float(th)
        win.frameEnergyTH[newest] = th
        states_log.append(lin["newState"].copy()); energy_log.append(float(lin["energy"])); th_log.append(float(th))
        return lin

    lin = linearize(False)
    ba.apply()
    iters = 0
    for it in range(max_iters):
        sys3 = ba.accumulate()
        x = solve_system([(H.copy(), b.copy()) for H, b in sys3], n)
        step_pts = ba.resubstitute(x)
        # doStepFromBackup (FullSystemOptimize.cpp:171-264): frames, points (idepth AND idepth_zero follow the step)
        sumA = sumB = sumT = sumR = 0.0
        for f, k in enumerate(kfs):
            st = -x[4 + 8 * f: 12 + 8 * f]
            k.state[:8] = k.state[:8] + st
            sumA += st[6] ** 2; sumB += st[7] ** 2; sumT += float((st[0:3] ** 2).sum()); sumR += float((st[3:6] ** 2).sum())
        win.idepth = (win.idepth + step_pts).astype(np.float32)
        win.idepth_zero = win.idepth.copy()
        sumNID = float(np.abs(win.idepth).mean()) if len(win.idepth) else 0.0
        # setPrecalcValues + setDeltaF at the new state
        win.state = np.array([k.state for k in kfs])
        win.precalc, win.adHost, win.adTarget, win.adHTdeltaF = synth.prepare_window_tables(
            cam, [k.T_eval for k in kfs], win.state, win.state_zero, win.ab_exposure)
        win.frameDeltaPrior = win.state[:, :8].reshape(-1).copy()
        ba.set_tables(win)
        ba.update_points(win.idepth, win.deltaF, win.idepth_zero)
        lin = linearize(False)
        ba.apply()                                    # setting_forceAceptStep: the step is always accepted
        iters += 1
        th_it = 1.2                                   # setting_thOptIterations
        canbreak = (np.sqrt(sumA / n) < 0.0005 * th_it and np.sqrt(sumB / n) < 0.00005 * th_it and np.sqrt(sumR / n) < 0.00005 * th_it and
                    np.sqrt(sumT / n) * sumNID < 0.00005 * th_it)
        if canbreak and it >= 1:                      # setting_minOptIterations = 1
            break
    lin = linearize(True)                             # final linearizeAll(true): applyRes inside
    log["ba_iterations"].append(iters); log["ba_states"].append(states_log); log["ba_energy"].append(energy_log)
    log["thresholds"].append(th_log)
    # points back into their keyframes
    off = 0
    for k in kfs:
        m = len(k.u)
        k.idepth = win.idepth[off:off + m].copy()
        off += m
    return lin


def run_sequence_ba(cam, Ts_true, frames, levels, backend, kf_every=8, pts_per_kf=400, max_kf=7, max_iters=6, seed=5):
    """Track every frame, optimise every keyframe's window; `backend` supplies the hot-path calls (CudaBackend / the oracle
    back end of the tests).  frames: list of (image, depth)."""
    rng = np.random.default_rng(seed)
    n = len(frames)
    log = dict(ba_iterations=[], ba_states=[], ba_energy=[], thresholds=[])
    wall = dict(track=0.0, ba=0.0, ba_backend=0.0, reference=0.0)
    est = [np.array(Ts_true[0], np.float64)]
    oks = [True]

    def new_keyframe(k, T_w2c, aff, th):
        img, depth = frames[k]
        u, v, idp, col, wts = select_points(img, depth, pts_per_kf, rng)
        st = np.zeros(10); st[6] = aff[0] / synth.SCALE_A; st[7] = aff[1] / synth.SCALE_B
        return KeyFrame(k, img, backend.new_frame(img, levels), np.array(T_w2c, np.float64), st.copy(), st.copy(), th, u, v, idp, col, wts)

    kfs = [new_keyframe(0, est[0], (0.0, 0.0), 8.0 * 8.0 * 8.0)]
    kf_poses = [kfs[0].T_now()]
    # the first keyframe has no window yet: its reference cloud comes from its own points (CoarseTracker.cpp:119-125 with
    # the points' own pixel as centerProjectedTo and a unit weight)
    t0 = time.perf_counter()
    trk = backend.tracker_from_points(cam, levels, kfs[0])
    wall["reference"] += time.perf_counter() - t0
    last = synth.T12(np.eye(3), np.zeros(3)); last2 = None
    aff = np.zeros(2)
    for k in range(1, n):
        T0 = last if last2 is None else synth.se3_mul(synth.se3_mul(last, synth.se3_inv(last2)), last)
        t0 = time.perf_counter()
        ok, T, aff_new = trk.track(frames[k][0], T0, aff, levels - 1)
        if not ok:
            ok, T, aff_new = trk.track(frames[k][0], last, aff, levels - 1)
        wall["track"] += time.perf_counter() - t0
        oks.append(bool(ok))
        if ok:
            last2, last, aff = last, np.array(T, np.float64), np.array(aff_new, np.float64)
        T_ref = kfs[-1].T_now()
        est.append(synth.se3_mul(last, T_ref))
        if k % kf_every == 0 and k + 1 < n:
            # ---- AddKeyframe (Mapping.cpp:12-142)
            kf = new_keyframe(k, est[-1], aff, kfs[-1].frameEnergyTH)     # :64 the new frame inherits the threshold
            kfs.append(kf)
            if len(kfs) > max_kf:
                old = kfs.pop(0)
                backend.free_frame(old.handle)
            # the two oldest keyframes are the gauge: they are held (strong prior, build_window) at their LATEST estimate,
            # i.e. their linearisation point is re-based on it (stand-in for the marginalisation prior, which would
            # carry that information in the reference)
            for q in kfs[:2]:
                q.T_eval = q.T_now(); q.state[:6] = 0.0; q.state_zero[:6] = 0.0
            t0 = time.perf_counter()
            win = build_window(cam, kfs)
            tb = time.perf_counter()
            ba = backend.open_ba(win, [q.handle for q in kfs])
            wall["ba_backend"] += time.perf_counter() - tb
            lin = optimize_window(cam, kfs, _Timed(ba, wall), max_iters, log)
            wall["ba"] += time.perf_counter() - t0
            # setEvalPT of the newest frame (FullSystemOptimize.cpp:498-506): its linearisation point becomes the estimate
            kf.T_eval = kf.T_now(); kf.state[:6] = 0.0; kf.state_zero = kf.state.copy()
            kf_poses.append(kf.T_now())
            est[-1] = kf.T_now()
            # ---- new tracking reference: makeCoarseDepthL0 from the residuals into the newest keyframe that are IN
            t0 = time.perf_counter()
            ids = np.nonzero((win.target_idx == len(kfs) - 1) & (lin["newState"] == RES_IN))[0].astype(np.int32)
            trk.close()
            trk = backend.tracker_from_ba(cam, levels, kf, getattr(ba, "_ba", ba), win, lin, ids)
            wall["reference"] += time.perf_counter() - t0
            ba.close()
            last, last2, aff = synth.T12(np.eye(3), np.zeros(3)), None, np.array(kf.aff_now())
    trk.close()
    for q in kfs:
        backend.free_frame(q.handle)
    if hasattr(backend, "close"):
        backend.close()

    def centre(T):
        R, t = synth.T12_split(T)
        return -R.T @ t
    err = np.array([centre(a) - centre(b) for a, b in zip(est, Ts_true)])
    return SequenceBAResult(est, list(Ts_true), oks, float(np.sqrt((err ** 2).sum(1).mean())), len(kf_poses), log["ba_iterations"],
                            log["ba_states"], log["ba_energy"], log["thresholds"], kf_poses, wall)


# ------------------------------------------------------------------------------------------------ CUDA back end
class _CudaTracker:
    """The tracker and the incoming-frame buffers live as long as the back end (the reference keeps two CoarseTracker
    instances and swaps them, System.cpp:127-133): a new keyframe only rebuilds the reference cloud."""

    def __init__(self, backend, trk, levels):
        self.backend, self.trk, self.levels = backend, trk, levels

    def track(self, img, T0, aff0, coarsest):
        b = self.backend
        if b._frame is None:
            b._frame = b.capi.Frame(b.ctx, img, self.levels)
        else:
            b._frame.update(img)
        return self.trk.trackNewestCoarse(b._frame, T0, aff0, coarsest)

    def close(self):
        pass


class _CudaBA:
    def __init__(self, capi, ctx, win, handles):
        self.win, self.handles = win, handles
        self.ef = capi.EnergyFunctional(ctx)
        self.ef.set_window(win, handles)
        self.ef.set_topology(win)

    def set_tables(self, win):
        self.ef.set_window(win, self.handles)

    def update_points(self, idepth, deltaF, idepth_zero):
        self.ef.update_points(idepth=idepth, deltaF=deltaF, idepth_zero=idepth_zero)

    def linearize(self, fix):
        return self.ef.linearizeAll(fix, outputs="states")

    def new_frame_energy_th(self, newest):
        return np.float32(self.ef.setNewFrameEnergyTH(newest))     # exact k-th select on the device

    def apply(self):
        self.ef.applyRes()

    def accumulate(self):
        return self.ef.accumulate_all()

    def resubstitute(self, x):
        return self.ef.resubstituteF_MT(x)

    def close(self):
        self.ef.close()


class CudaBackend:
    """The product path: every hot-path call goes through the C ABI (h-slam_b200/capi.py); there is no CPU fallback."""

    def __init__(self, capi, ctx):
        self.capi, self.ctx = capi, ctx
        self._trk = None          # one CoarseTracker, re-referenced at every keyframe
        self._frame = None        # one device frame for the incoming images

    def new_frame(self, img, levels):
        return self.capi.Frame(self.ctx, img, levels)

    def free_frame(self, handle):
        handle.close()

    def open_ba(self, win, handles):
        return _CudaBA(self.capi, self.ctx, win, handles)

    def _tracker(self, cam, levels):
        if self._trk is None:
            self._trk = self.capi.CoarseTracker(self.ctx, cam.w, cam.h)
            self._trk.makeK(cam, levels)
        return self._trk

    def tracker_from_points(self, cam, levels, kf):
        trk = self._tracker(cam, levels)
        cpt = np.stack([kf.u, kf.v, kf.idepth], 1).astype(np.float32)
        hdi = np.full(len(kf.u), 1e-3, np.float32)            # weight sqrt(1e-3 / (HdiF + 1e-12)) ~ 1
        trk.makeCoarseDepthL0(kf.handle, cpt, hdi, ref_aff=kf.aff_now())
        return _CudaTracker(self, trk, levels)

    def tracker_from_ba(self, cam, levels, kf, ba, win, lin, ids):
        trk = self._tracker(cam, levels)
        # device-resident: centerProjectedTo of the last linearisation and HdiF of the last Schur pass never leave HBM
        trk.makeCoarseDepthL0_from_ba(kf.handle, ba.ef, ids, ref_aff=kf.aff_now())
        return _CudaTracker(self, trk, levels)

    def close(self):
        if self._trk is not None:
            self._trk.close(); self._trk = None
        if self._frame is not None:
            self._frame.close(); self._frame = None
