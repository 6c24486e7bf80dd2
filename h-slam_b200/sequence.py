"""Frame-to-keyframe tracking over a rendered sequence (BASELINE.json configs[2]).

Host-side driver restated from the tracking part of the reference's System::ProcessFrame /
trackNewCoarse (Src/System.cpp:334-470): every frame is tracked against the latest keyframe with
CoarseTracker::trackNewestCoarse, initialised with the constant-motion hypothesis (the first entry of
lastF_2_fh_tries, System.cpp:367-372) and the previous affine brightness; a new keyframe is made every
`kf_every` frames and its reference cloud comes from CoarseTracker::makeCoarseDepthL0's rule
(synth.make_coarse_depth_np, Src/CoarseTracker.cpp:105-263) applied to the keyframe's depth.

Everything outside the hot path (keyframe policy, mapping, marginalisation) is deliberately minimal: the
driver exists to run the SAME decisions through two tracker back ends (the CUDA tracker and, in tests, the
CPU oracle) and compare their trajectories.  `tracker_factory(cam, levels, pc)` returns an object with
`track(img_new, T0, aff0, coarsest) -> (ok, T_ref2new(12), aff(2))`.
"""
from __future__ import annotations

import dataclasses
import numpy as np

from . import synth


@dataclasses.dataclass
class SequenceResult:
    T_w2c_est: list          # estimated world->camera (12) per frame
    T_w2c_true: list
    ok: list
    ate_rmse: float          # RMSE of camera centres vs ground truth (no alignment: both start at the origin)
    n_keyframes: int


def trajectory(n_frames, step=0.012, seed=0):
    """Smooth world->camera poses: forward/sideways drift plus slow rotation (~1.2 cm, ~0.25 deg per frame)."""
    rng = np.random.default_rng(seed)
    ph = rng.uniform(0, 2 * np.pi, 6)
    out = []
    for k in range(n_frames):
        s = k * step
        xi = np.array([0.6 * s + 0.03 * np.sin(0.9 * s * 8 + ph[0]), 0.02 * np.sin(0.7 * s * 8 + ph[1]), 0.9 * s,
                       0.02 * np.sin(0.5 * s * 8 + ph[3]), 0.25 * s + 0.02 * np.sin(0.6 * s * 8 + ph[4]), 0.01 * np.sin(0.4 * s * 8 + ph[5])])
        R, t = synth.se3_exp(xi)
        out.append(synth.T12(R, t))
    return out


def make_reference_cloud(cam, levels, img, depth, n_points, rng, idepth_noise=0.01):
    """Seed points on high-gradient pixels with (noisy) true inverse depth -> per-level pc_* by the reference rule."""
    pyr, ag = synth.build_pyramid_np(img, levels)
    h, w = img.shape
    ys, xs = np.mgrid[2:h - 2, 2:w - 2]
    ys, xs = ys.reshape(-1), xs.reshape(-1)
    good = (ag[0][ys, xs] > 50) & np.isfinite(depth[ys, xs])
    ys, xs = ys[good], xs[good]
    sel = rng.choice(len(xs), size=min(n_points, len(xs)), replace=False)
    sel.sort()
    su, sv = xs[sel].astype(np.int32), ys[sel].astype(np.int32)
    idep = (1.0 / depth[sv, su]) * (1.0 + rng.normal(0, idepth_noise, len(sel)))
    return synth.make_coarse_depth_np(levels, w, h, su, sv, idep.astype(np.float32), np.ones(len(sel), np.float32), pyr)


def render_sequence(w, h, n_frames, seed=1234, aff_drift=True):
    cam = synth.Camera.tum_like(w, h)
    scene = synth.Scene(seed)
    Ts = trajectory(n_frames, seed=seed)
    frames = []
    for k, T in enumerate(Ts):
        aff = (0.01 * np.sin(0.05 * k), 2.0 * np.sin(0.03 * k)) if aff_drift else (0.0, 0.0)
        img, depth = scene.render(cam, T, aff=aff)
        frames.append((img, depth))
    return cam, Ts, frames


def run_sequence(cam, Ts_true, frames, levels, tracker_factory, kf_every=8, n_points=2000, seed=5):
    """Track every frame against the latest keyframe; returns SequenceResult."""
    rng = np.random.default_rng(seed)
    n = len(frames)
    est = [np.array(Ts_true[0], np.float64)]
    oks = [True]
    kf = 0
    trk = tracker_factory(cam, levels, make_reference_cloud(cam, levels, frames[0][0], frames[0][1], n_points, rng))
    T_kf_est = est[0]
    last = synth.T12(np.eye(3), np.zeros(3))        # ref -> previous frame
    last2 = None                                    # ref -> the frame before
    aff = np.zeros(2)
    n_kf = 1
    for k in range(1, n):
        # constant-motion hypothesis (System.cpp:367-372): ref->new = (prev motion) * ref->prev
        T0 = last if last2 is None else synth.se3_mul(synth.se3_mul(last, synth.se3_inv(last2)), last)
        ok, T, aff_new = trk.track(frames[k][0], T0, aff, levels - 1)
        if not ok:                                   # fall back to "no motion" (second hypothesis of the reference)
            ok, T, aff_new = trk.track(frames[k][0], last, aff, levels - 1)
        oks.append(bool(ok))
        if ok:
            last2, last, aff = last, np.array(T, np.float64), np.array(aff_new, np.float64)
        est.append(synth.se3_mul(last, T_kf_est))
        if k % kf_every == 0 and k + 1 < n:          # new keyframe: reference cloud from this frame's depth
            kf, T_kf_est = k, est[-1]
            trk.close()
            trk = tracker_factory(cam, levels, make_reference_cloud(cam, levels, frames[k][0], frames[k][1], n_points, rng))
            last, last2, aff = synth.T12(np.eye(3), np.zeros(3)), None, np.zeros(2)
            n_kf += 1
    trk.close()

    def centre(T):
        R, t = synth.T12_split(T)
        return -R.T @ t
    err = np.array([centre(a) - centre(b) for a, b in zip(est, Ts_true)])
    return SequenceResult(est, list(Ts_true), oks, float(np.sqrt((err ** 2).sum(1).mean())), n_kf)


class CudaTracker:
    """tracker_factory product: the CUDA CoarseTracker behind the C ABI (no CPU fallback)."""

    def __init__(self, capi, ctx, cam, levels, pc):
        self.capi, self.ctx, self.levels = capi, ctx, levels
        self.trk = capi.CoarseTracker(ctx, cam.w, cam.h)
        self.trk.makeK(cam, levels)
        self.trk.setCoarseTrackingRef(pc)
        self.frame = None

    def track(self, img, T0, aff0, coarsest):
        if self.frame is None:
            self.frame = self.capi.Frame(self.ctx, img, self.levels)
        else:
            self.frame.update(img)
        return self.trk.trackNewestCoarse(self.frame, T0, aff0, coarsest)

    def close(self):
        self.trk.close()
        if self.frame is not None:
            self.frame.close()
