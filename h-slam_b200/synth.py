"""Synthetic rendered inputs for the hot path (no datasets offline).

Pure numpy, deterministic per seed.  Produces
  * analytic images + depth of a textured "room" (floor, ceiling, walls) at arbitrary poses,
  * config 1: a tracking frame pair with a reference point cloud built by the reference rule
    (CoarseTracker::makeCoarseDepthL0, Src/CoarseTracker.cpp:105-263, restated in numpy),
  * config 2/4/5: a keyframe window with points, residual topology and the per-pair tables the
    host side of the reference computes (FrameFramePrecalc::set, Src/OptimizationClasses.cpp:13-39;
    EnergyFunctional::setAdjointsF / setDeltaF, Src/EnergyFunctional.cpp:22-82,128-152).

This is workload generation for tests/bench/smoke; it is not the oracle and does not use it.
"""
from __future__ import annotations

import dataclasses
import numpy as np

PATTERN = np.array([[0, -2], [-1, -1], [1, -1], [-2, 0], [0, 0], [2, 0], [-1, 1], [0, 2]], dtype=np.int32)

SCALE_XI_ROT = 1.0
SCALE_XI_TRANS = 0.5
SCALE_A = 10.0
SCALE_B = 1000.0
SCALE_F = 50.0
SCALE_C = 50.0
PRECALC_STRIDE = 28


# ----------------------------------------------------------------------------- SE3 helpers (double)
def hat(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=np.float64)


def se3_exp(xi):
    """Sophus SE3::exp, tangent = [upsilon, omega] (sophus/se3.hpp:407-427). Returns (R, t)."""
    xi = np.asarray(xi, dtype=np.float64)
    ups, om = xi[:3], xi[3:6]
    th = np.linalg.norm(om)
    O = hat(om)
    O2 = O @ O
    if th < 1e-10:
        R = np.eye(3) + O + 0.5 * O2
        V = R
    else:
        R = np.eye(3) + np.sin(th) / th * O + (1 - np.cos(th)) / th**2 * O2
        V = np.eye(3) + (1 - np.cos(th)) / th**2 * O + (th - np.sin(th)) / th**3 * O2
    return R, V @ ups


def T12(R, t):
    return np.concatenate([np.asarray(R, np.float64).reshape(9), np.asarray(t, np.float64).reshape(3)])


def T12_split(T):
    T = np.asarray(T, np.float64)
    return T[:9].reshape(3, 3), T[9:12]


def se3_mul(A, B):
    Ra, ta = T12_split(A)
    Rb, tb = T12_split(B)
    return T12(Ra @ Rb, Ra @ tb + ta)


def se3_inv(A):
    R, t = T12_split(A)
    return T12(R.T, -R.T @ t)


def se3_adj(T):
    R, t = T12_split(T)
    Ad = np.zeros((6, 6))
    Ad[:3, :3] = R
    Ad[3:, 3:] = R
    Ad[:3, 3:] = hat(t) @ R
    return Ad


def aff_from_to(eF, eT, aF, bF, aT, bT):
    """AffLight::fromToVecExposure (Include/GlobalTypes.h:334-346)."""
    if eF == 0 or eT == 0:
        eF = eT = 1.0
    a = np.exp(aT - aF) * float(np.float32(eT)) / float(np.float32(eF))
    return a, bT - a * bF


# ----------------------------------------------------------------------------- camera / scene
@dataclasses.dataclass
class Camera:
    w: int
    h: int
    fx: float
    fy: float
    cx: float
    cy: float

    @staticmethod
    def tum_like(w, h):
        # Src/GeometricUndistorter.cpp:119-122 with Extras/Calib/TumMono*.yaml desiK
        return Camera(w, h, np.float32(0.4 * w), np.float32(0.53 * h), np.float32(0.5 * w - 0.5), np.float32(0.5 * h - 0.5))


class Scene:
    """Axis-aligned textured room around the origin; camera looks along +z, y down."""

    def __init__(self, seed=1234, n_waves=14):
        rng = np.random.default_rng(seed)
        # planes: (normal, d) with n.X = d ; in-plane axes e1,e2
        self.planes = []
        defs = [((0, 1, 0), 1.5), ((0, -1, 0), 1.3), ((0, 0, 1), 5.0), ((1, 0, 0), 3.2), ((-1, 0, 0), 3.0)]
        for nrm, d in defs:
            nrm = np.array(nrm, np.float64)
            e1 = np.array([1.0, 0, 0]) if abs(nrm[0]) < 0.5 else np.array([0, 0, 1.0])
            e2 = np.cross(nrm, e1)
            lam = rng.uniform(0.18, 1.1, n_waves)  # wavelengths in metres
            ang = rng.uniform(0, np.pi, n_waves)
            amp = rng.uniform(6, 16, n_waves)
            ph = rng.uniform(0, 2 * np.pi, n_waves)
            fa = np.cos(ang) / lam
            fb = np.sin(ang) / lam
            base = rng.uniform(105, 150)
            self.planes.append((nrm, float(d), e1, e2, fa, fb, amp, ph, base))

    def render(self, cam: Camera, T_w2c, aff=(0.0, 0.0)):
        """Returns (image float32 HxW in ~[0,255], depth float64 HxW) for world->camera pose T (12)."""
        R, t = T12_split(T_w2c)
        C = -R.T @ t
        ys, xs = np.mgrid[0 : cam.h, 0 : cam.w].astype(np.float64)
        dc = np.stack([(xs - cam.cx) / cam.fx, (ys - cam.cy) / cam.fy, np.ones_like(xs)], -1)  # camera rays, z=1
        dw = dc @ R  # = (R^T dc^T)^T
        best = np.full(xs.shape, np.inf)
        img = np.zeros(xs.shape)
        for nrm, d, e1, e2, fa, fb, amp, ph, base in self.planes:
            denom = dw @ nrm
            with np.errstate(divide="ignore", invalid="ignore"):
                s = (d - nrm @ C) / denom
            ok = (denom > 1e-9) & (s > 0.05) & (s < best)
            X = C[None, None, :] + s[..., None] * dw
            a = X @ e1
            b = X @ e2
            val = np.full(xs.shape, base)
            for k in range(len(fa)):
                val = val + amp[k] * np.sin(2 * np.pi * (fa[k] * a + fb[k] * b) + ph[k])
            img = np.where(ok, val, img)
            best = np.where(ok, s, best)
        img = np.exp(aff[0]) * img + aff[1]
        img = np.clip(img, 0.0, 255.0)
        return img.astype(np.float32), best


# ----------------------------------------------------------------------------- numpy pyramid (host mirror)
def pyramid_levels(w, h, max_levels):
    """Level count by the reference rule (Include/CalibData.h:110)."""
    lv, w_, h_ = 1, w, h
    while w_ % 2 == 0 and h_ % 2 == 0 and w_ * h_ > 5000 and lv < max_levels:
        w_ //= 2
        h_ //= 2
        lv += 1
    return lv


def build_pyramid_np(img, levels):
    """Frame::CreateDirPyrs (Src/Frame.cpp:104-181) in numpy float32; returns list of (h,w,3) + absgrad."""
    out, ag = [], []
    I = np.asarray(img, np.float32)
    for lvl in range(levels):
        if lvl > 0:
            p = out[-1][..., 0]
            h2, w2 = p.shape[0] // 2, p.shape[1] // 2
            p = p[: 2 * h2, : 2 * w2]
            I = np.float32(0.25) * (((p[0::2, 0::2] + p[0::2, 1::2]) + p[1::2, 0::2]) + p[1::2, 1::2])
        h, w = I.shape
        flat = I.reshape(-1)
        dI = np.zeros((h * w, 3), np.float32)
        dI[:, 0] = flat
        idx = np.arange(w, w * (h - 1))
        dx = np.float32(0.5) * (flat[idx + 1] - flat[idx - 1])
        dy = np.float32(0.5) * (flat[idx + w] - flat[idx - w])
        dx = np.where(np.isfinite(dx), dx, np.float32(0))
        dy = np.where(np.isfinite(dy), dy, np.float32(0))
        dI[idx, 1] = dx
        dI[idx, 2] = dy
        a = np.zeros(h * w, np.float32)
        a[idx] = dx * dx + dy * dy
        out.append(dI.reshape(h, w, 3))
        ag.append(a.reshape(h, w))
    return out, ag


def make_K_levels(cam: Camera, levels):
    """CoarseTracker::makeK (Src/CoarseTracker.cpp:72-101): list of (w,h,fx,fy,cx,cy)."""
    L = [(cam.w, cam.h, np.float32(cam.fx), np.float32(cam.fy), np.float32(cam.cx), np.float32(cam.cy))]
    for l in range(1, levels):
        p = L[-1]
        L.append(
            (
                cam.w >> l,
                cam.h >> l,
                np.float32(float(p[2]) * 0.5),
                np.float32(float(p[3]) * 0.5),
                np.float32((float(cam.cx) + 0.5) / (1 << l) - 0.5),
                np.float32((float(cam.cy) + 0.5) / (1 << l) - 0.5),
            )
        )
    return L


def make_coarse_depth_np(levels, w0, h0, su, sv, sid, sw, ref_pyr):
    """CoarseTracker::makeCoarseDepthL0 (Src/CoarseTracker.cpp:105-263) in numpy float32.

    Returns per level (pc_u, pc_v, pc_idepth, pc_color) float32 arrays in the reference's
    row-major scan order.
    """
    f32 = np.float32
    idepth = [np.zeros((h0 >> l, w0 >> l), f32) for l in range(levels)]
    wsum = [np.zeros((h0 >> l, w0 >> l), f32) for l in range(levels)]
    # sequential float accumulation at level 0 (duplicates are rare but keep the order)
    np.add.at(idepth[0], (sv, su), (sid.astype(f32) * sw.astype(f32)).astype(f32))
    np.add.at(wsum[0], (sv, su), sw.astype(f32))
    for l in range(1, levels):
        hl, wl = idepth[l].shape
        a, b = idepth[l - 1][: 2 * hl, : 2 * wl], wsum[l - 1][: 2 * hl, : 2 * wl]
        idepth[l] = ((a[0::2, 0::2] + a[0::2, 1::2]) + a[1::2, 0::2]) + a[1::2, 1::2]
        wsum[l] = ((b[0::2, 0::2] + b[0::2, 1::2]) + b[1::2, 0::2]) + b[1::2, 1::2]

    def dilate(l, offs):
        h, w = idepth[l].shape
        # the reference reads one element past the end for the last candidate (i+1+w == w*h);
        # a zero pad reproduces "weight 0 => skipped"
        ws = np.concatenate([wsum[l].reshape(-1), np.zeros(1, f32)])
        wb = ws.copy()
        idl = np.concatenate([idepth[l].reshape(-1), np.zeros(1, f32)])
        i = np.arange(w, w * h - w)
        cand = i[wb[i] <= 0]
        s = np.zeros(cand.shape, f32)
        num = np.zeros(cand.shape, f32)
        numn = np.zeros(cand.shape, f32)
        for o in offs:
            m = wb[cand + o] > 0
            s = np.where(m, s + idl[cand + o], s).astype(f32)
            num = np.where(m, num + wb[cand + o], num).astype(f32)
            numn = np.where(m, numn + f32(1), numn).astype(f32)
        ok = numn > 0
        with np.errstate(divide="ignore", invalid="ignore"):
            idl[cand[ok]] = (s[ok] / numn[ok]).astype(f32)
            ws[cand[ok]] = (num[ok] / numn[ok]).astype(f32)
        idepth[l] = idl[:-1].reshape(h, w)
        wsum[l] = ws[:-1].reshape(h, w)

    for l in range(min(2, levels)):
        w = w0 >> l
        dilate(l, [1 + w, -1 - w, w - 1, -w + 1])
    for l in range(2, levels):
        w = w0 >> l
        dilate(l, [1, -1, w, -w])
    out = []
    for l in range(levels):
        h, w = idepth[l].shape
        ys, xs = np.mgrid[2 : h - 2, 2 : w - 2]
        ys, xs = ys.reshape(-1), xs.reshape(-1)
        ws = wsum[l][ys, xs]
        m = ws > 0
        ys, xs = ys[m], xs[m]
        with np.errstate(divide="ignore", invalid="ignore"):
            idv = (idepth[l][ys, xs] / wsum[l][ys, xs]).astype(f32)
        col = ref_pyr[l][ys, xs, 0]
        ok = np.isfinite(col) & (idv > 0)
        out.append((xs[ok].astype(f32), ys[ok].astype(f32), idv[ok], col[ok].astype(f32)))
    return out


# ----------------------------------------------------------------------------- config 1: tracking pair
@dataclasses.dataclass
class TrackingPair:
    cam: Camera
    levels: int
    img_ref: np.ndarray
    img_new: np.ndarray
    depth_ref: np.ndarray
    pc: list  # per level (u, v, idepth, color)
    T_ref2new_true: np.ndarray  # 12
    aff_new_true: tuple
    ref_exposure: float = 1.0
    new_exposure: float = 1.0
    ref_aff: tuple = (0.0, 0.0)


def make_tracking_pair(w=640, h=480, levels=5, n_seed_points=6200, seed=1234, motion_scale=1.0):
    """Config 1: one frame pair ~2 cm / 0.5 deg apart, target brightened a=0.02,b=3."""
    rng = np.random.default_rng(seed)
    cam = Camera.tum_like(w, h)
    scene = Scene(seed)
    T_ref = T12(np.eye(3), np.zeros(3))
    xi = np.array([0.012, -0.006, 0.015, 0.004, -0.006, 0.003]) * motion_scale  # ~2 cm, ~0.45 deg
    R, t = se3_exp(xi)
    T_new = T12(R, t)  # world == ref camera, so ref->new == world->new
    aff_new = (0.02, 3.0)
    img_ref, depth_ref = scene.render(cam, T_ref)
    img_new, _ = scene.render(cam, T_new, aff=aff_new)
    pyr, ag = build_pyramid_np(img_ref, levels)
    ys, xs = np.mgrid[2 : h - 2, 2 : w - 2]
    ys, xs = ys.reshape(-1), xs.reshape(-1)
    good = ag[0][ys, xs] > 50
    ys, xs = ys[good], xs[good]
    sel = rng.choice(len(xs), size=min(n_seed_points, len(xs)), replace=False)
    sel.sort()
    su, sv = xs[sel].astype(np.int32), ys[sel].astype(np.int32)
    idep = (1.0 / depth_ref[sv, su]) * (1.0 + rng.normal(0, 0.02, len(sel)))
    pc = make_coarse_depth_np(levels, w, h, su, sv, idep.astype(np.float32), np.ones(len(sel), np.float32), pyr)
    return TrackingPair(cam, levels, img_ref, img_new, depth_ref, pc, T_new, aff_new)


# ----------------------------------------------------------------------------- config 2: BA window
@dataclasses.dataclass
class Window:
    cam: Camera
    n: int
    images: list           # n float32 HxW
    T_w2c_eval: np.ndarray  # n x 12 (worldToCam_evalPT)
    state: np.ndarray       # n x 10
    state_zero: np.ndarray  # n x 10
    ab_exposure: np.ndarray  # n float32
    frameEnergyTH: np.ndarray
    precalc: np.ndarray     # n*n*28 float32
    adHost: np.ndarray      # n*n*64 float64 (col-major blocks)
    adTarget: np.ndarray
    adHTdeltaF: np.ndarray  # n*n*8 float32
    cDeltaF: np.ndarray     # 4 float32
    cPrior: np.ndarray      # 4 float64
    framePrior: np.ndarray  # n*8
    frameDeltaPrior: np.ndarray
    # points
    u: np.ndarray
    v: np.ndarray
    idepth: np.ndarray
    idepth_zero: np.ndarray
    color: np.ndarray       # P x 8
    weights: np.ndarray     # P x 8
    priorF: np.ndarray
    deltaF: np.ndarray
    host_idx: np.ndarray
    res_begin: np.ndarray
    res_index: np.ndarray
    # residuals
    point_idx: np.ndarray
    target_idx: np.ndarray
    state_state: np.ndarray
    is_linearized: np.ndarray
    is_active: np.ndarray

    @property
    def P(self):
        return len(self.u)

    @property
    def R(self):
        return len(self.point_idx)


def prepare_window_tables(cam: Camera, T_eval, state, state_zero, exposure):
    """numpy restatement of FrameOptimizationData::setState (Include/Frame.h:184-197),
    FrameFramePrecalc::set (Src/OptimizationClasses.cpp:13-39), setAdjointsF/setDeltaF
    (Src/EnergyFunctional.cpp:22-82,128-152).  Index = host + target*n."""
    n = len(T_eval)
    f32 = np.float32
    pre_w2c, aff, aff0 = [], [], []
    for f in range(n):
        s = state[f]
        sc = np.concatenate([SCALE_XI_TRANS * s[0:3], SCALE_XI_ROT * s[3:6], [SCALE_A * s[6], SCALE_B * s[7]]])
        R, t = se3_exp(sc[:6])
        pre_w2c.append(se3_mul(T12(R, t), T_eval[f]))
        aff.append((sc[6], sc[7]))
        aff0.append((state_zero[f][6] * SCALE_A, state_zero[f][7] * SCALE_B))
    K = np.array([[cam.fx, 0, cam.cx], [0, cam.fy, cam.cy], [0, 0, 1]], f32)
    Ki = np.linalg.inv(K.astype(np.float64)).astype(f32)
    precalc = np.zeros((n * n, PRECALC_STRIDE), f32)
    adH = np.zeros((n * n, 64))
    adT = np.zeros((n * n, 64))
    adHTd = np.zeros((n * n, 8), f32)
    rs = np.array([SCALE_XI_TRANS] * 3 + [SCALE_XI_ROT] * 3 + [SCALE_A, SCALE_B])
    for h in range(n):
        for t in range(n):
            idx = h + t * n
            ll0 = se3_mul(T_eval[t], se3_inv(T_eval[h]))
            ll = se3_mul(pre_w2c[t], se3_inv(pre_w2c[h]))
            R = ll[:9].reshape(3, 3).astype(f32)
            tt = ll[9:].astype(f32)
            precalc[idx, 0:9] = ((K @ R).astype(f32) @ Ki).astype(f32).reshape(9)
            precalc[idx, 9:12] = (K @ tt).astype(f32)
            precalc[idx, 12:21] = ll0[:9].astype(f32)
            precalc[idx, 21:24] = ll0[9:].astype(f32)
            a, b = aff_from_to(exposure[h], exposure[t], aff[h][0], aff[h][1], aff[t][0], aff[t][1])
            precalc[idx, 24] = a
            precalc[idx, 25] = b
            precalc[idx, 26] = aff0[h][1]
            AH = np.eye(8)
            AT = np.eye(8)
            AH[:6, :6] = -se3_adj(ll0).T
            a0, _ = aff_from_to(exposure[h], exposure[t], aff0[h][0], aff0[h][1], aff0[t][0], aff0[t][1])
            a0 = float(f32(a0))
            AT[6, 6] = -a0
            AH[6, 6] = a0
            AT[7, 7] = -1
            AH[7, 7] = a0
            AH *= rs[:, None]
            AT *= rs[:, None]
            adH[idx] = AH.T.reshape(64)  # col-major
            adT[idx] = AT.T.reshape(64)
            dh = (state[h][:8] - state_zero[h][:8]).astype(f32)
            dt = (state[t][:8] - state_zero[t][:8]).astype(f32)
            adHTd[idx] = (dh @ AH.astype(f32) + dt @ AT.astype(f32)).astype(f32)
    return precalc.reshape(-1), adH.reshape(-1), adT.reshape(-1), adHTd.reshape(-1)


def _frame_priors(n):
    """FrameOptimizationData::getPrior (Include/Frame.h:229-256) with the default settings."""
    p = np.tile(np.array([0, 0, 0, 0, 0, 0, 1e12, 1e8], np.float64), (n, 1))
    p[0] = [1e10, 1e10, 1e10, 1e11, 1e11, 1e11, 1e14, 1e14]
    return p.reshape(-1)


@dataclasses.dataclass
class Keyframes:
    cam: Camera
    images: list
    depths: list
    T_true: list
    aff_true: list


def render_keyframes(w=640, h=480, n_frames=7, seed=4321):
    """n keyframes of the synthetic room on a smooth trajectory (5-10 cm steps), exact poses and depth."""
    cam = Camera.tum_like(w, h)
    scene = Scene(seed)
    n = n_frames
    kf = Keyframes(cam, [], [], [], [])
    for f in range(n):
        s = f - (n - 1) / 2
        xi = np.array([0.07 * s, 0.01 * np.sin(0.9 * f), 0.035 * s, 0.004 * s, -0.012 * s, 0.003 * np.cos(f)])
        R, t = se3_exp(xi)
        Tt = T12(R, t)
        a_f, b_f = 0.01 * f, 0.8 * f
        img, dep = scene.render(cam, Tt, aff=(a_f, b_f))
        kf.images.append(img); kf.depths.append(dep); kf.T_true.append(Tt); kf.aff_true.append((a_f, b_f))
    return kf


def make_window(w=640, h=480, n_frames=7, points_per_frame=2000, seed=4321, frac_linearized=0.0,
                pose_noise=4e-4, state_noise=3e-4, idepth_noise=0.003, keyframes=None):
    """Config 2: n keyframes on a smooth trajectory, points_per_frame points per host KF,
    residuals to every other KF, idepth_zero = idepth*(1+N(0,0.01)), state-state_zero ~ N(0,state_noise).
    `keyframes` re-uses rendered images (render_keyframes) so that many windows with different points,
    noise and states can be drawn from one rendering."""
    rng = np.random.default_rng(seed)
    kf = keyframes if keyframes is not None else render_keyframes(w, h, n_frames, seed)
    cam = kf.cam
    w, h = cam.w, cam.h
    n = len(kf.images)
    images, depths, aff_true = kf.images, kf.depths, kf.aff_true
    T_eval = []
    for f in range(n):
        Rn, tn = se3_exp(rng.normal(0, pose_noise, 6))
        T_eval.append(se3_mul(T12(Rn, tn), kf.T_true[f]))
    state_zero = np.zeros((n, 10))
    state = np.zeros((n, 10))
    for f in range(n):
        state_zero[f, 6] = aff_true[f][0] / SCALE_A
        state_zero[f, 7] = aff_true[f][1] / SCALE_B
        d = rng.normal(0, state_noise, 8)
        d[6] *= 0.1   # a is scaled x10, b x1000 when applied
        d[7] *= 0.3
        state[f, :8] = state_zero[f, :8] + d
    state[0, :6] = 0  # first frame fixed in practice
    exposure = np.ones(n, np.float32)
    precalc, adH, adT, adHTd = prepare_window_tables(cam, T_eval, state, state_zero, exposure)
    # points
    us, vs, idp, idz, col, wts, hosts = [], [], [], [], [], [], []
    c_th = np.float32(50.0 * 50.0)
    for f in range(n):
        pyr, ag = build_pyramid_np(images[f], 1)
        ys, xs = np.mgrid[6 : h - 6, 6 : w - 6]
        ys, xs = ys.reshape(-1), xs.reshape(-1)
        good = (ag[0][ys, xs] > 40) & np.isfinite(depths[f][ys, xs])
        ys, xs = ys[good], xs[good]
        sel = rng.choice(len(xs), size=min(points_per_frame, len(xs)), replace=False)
        sel.sort()
        x, y = xs[sel], ys[sel]
        idt = 1.0 / depths[f][y, x]
        idv = (idt * (1.0 + rng.normal(0, idepth_noise, len(sel)))).astype(np.float32)
        idzv = (idv * (1.0 + rng.normal(0, 0.01, len(sel)))).astype(np.float32)
        c = np.stack([pyr[0][y + PATTERN[k, 1], x + PATTERN[k, 0], 0] for k in range(8)], 1).astype(np.float32)
        g2 = np.stack(
            [pyr[0][y + PATTERN[k, 1], x + PATTERN[k, 0], 1] ** 2 + pyr[0][y + PATTERN[k, 1], x + PATTERN[k, 0], 2] ** 2
             for k in range(8)], 1).astype(np.float32)
        wv = np.sqrt(c_th / (c_th + g2)).astype(np.float32)
        us.append(x.astype(np.float32)); vs.append(y.astype(np.float32)); idp.append(idv); idz.append(idzv)
        col.append(c); wts.append(wv); hosts.append(np.full(len(sel), f, np.int32))
    u = np.concatenate(us); v = np.concatenate(vs); idepth = np.concatenate(idp); idepth_zero = np.concatenate(idz)
    color = np.concatenate(col); weights = np.concatenate(wts); host_idx = np.concatenate(hosts)
    P = len(u)
    # residual topology: each point -> every other frame, stored in a shuffled per-point order
    others = np.arange(n - 1)[None, :] + (np.arange(n - 1)[None, :] >= host_idx[:, None])   # P x (n-1), skips the host
    others = rng.permuted(others, axis=1)
    R_ = P * (n - 1)
    point_idx = np.repeat(np.arange(P, dtype=np.int32), n - 1)
    target_idx = others.reshape(-1).astype(np.int32)
    res_begin = np.arange(P + 1, dtype=np.int32) * (n - 1)
    res_index = np.arange(R_, dtype=np.int32)
    is_lin = (rng.random(R_) < frac_linearized).astype(np.uint8)
    priorF = np.where(rng.random(P) < 0.1, np.float32(50.0 * 50.0), np.float32(0)).astype(np.float32)
    return Window(
        cam=cam, n=n, images=images, T_w2c_eval=np.array(T_eval), state=state, state_zero=state_zero,
        ab_exposure=exposure, frameEnergyTH=np.full(n, 8 * 8 * 8, np.float32), precalc=precalc, adHost=adH,
        adTarget=adT, adHTdeltaF=adHTd, cDeltaF=np.array([2e-4, -1e-4, 1.5e-4, -2.5e-4], np.float32),
        cPrior=np.full(4, 5e9), framePrior=_frame_priors(n),
        frameDeltaPrior=state[:, :8].copy().reshape(-1), u=u, v=v, idepth=idepth, idepth_zero=idepth_zero, color=color,
        weights=weights, priorF=priorF, deltaF=(idepth - idepth_zero).astype(np.float32), host_idx=host_idx,
        res_begin=np.array(res_begin, np.int32), res_index=np.array(res_index, np.int32),
        point_idx=np.array(point_idx, np.int32), target_idx=np.array(target_idx, np.int32),
        state_state=np.zeros(R_, np.int8), is_linearized=np.zeros(R_, np.uint8), is_active=np.zeros(R_, np.uint8),
    ), is_lin


# ----------------------------------------------------------------------------- immature points (§8f rank 2)
PATTERN = np.array([[0, -2], [-1, -1], [1, -1], [-2, 0], [0, 0], [2, 0], [-1, 1], [0, 2]], np.int32)  # GlobalTypes.h:181


@dataclasses.dataclass
class ImmatureSet:
    """SoA of ImmaturePoint members (Include/ImmaturePoint.h:33-66) for the points of several host keyframes, plus the
    per-host tables System::traceNewCoarse forms for one new frame (Src/Mapping.cpp:503-517)."""
    cam: Camera
    host_idx: np.ndarray
    u: np.ndarray
    v: np.ndarray
    gradH: np.ndarray       # n x 4, Mat22f col-major
    color: np.ndarray       # n x 8
    weights: np.ndarray     # n x 8
    energyTH: np.ndarray
    idepth_min: np.ndarray
    idepth_max: np.ndarray
    quality: np.ndarray
    status: np.ndarray      # int8 HSLAM_IPS_*
    lastTraceUV: np.ndarray
    lastTracePixelInterval: np.ndarray
    idepth_true: np.ndarray
    KRKi: np.ndarray = None  # n_hosts x 9
    Kt: np.ndarray = None    # n_hosts x 3
    aff: np.ndarray = None   # n_hosts x 2


def make_immature_points(kf: Keyframes, hosts, points_per_host=500, seed=7, outlierTHSumComponent=2500.0, outlierTH=144.0):
    """ImmaturePoint::ImmaturePoint (Src/ImmaturePoint.cpp:8-38) at integer pixels of high gradient."""
    rng = np.random.default_rng(seed)
    cam = kf.cam
    cols = {k: [] for k in ("host_idx", "u", "v", "gradH", "color", "weights", "idt")}
    for hi, f in enumerate(hosts):
        pyr, ag = build_pyramid_np(kf.images[f], 1)
        dI = pyr[0]                                   # h x w x 3
        ys, xs = np.mgrid[6:cam.h - 6, 6:cam.w - 6]
        ys, xs = ys.reshape(-1), xs.reshape(-1)
        good = (ag[0][ys, xs] > 60) & np.isfinite(kf.depths[f][ys, xs])
        ys, xs = ys[good], xs[good]
        sel = np.sort(rng.choice(len(xs), size=min(points_per_host, len(xs)), replace=False))
        for x, y in zip(xs[sel], ys[sel]):
            px = dI[y + PATTERN[:, 1], x + PATTERN[:, 0]]          # integer positions: the bilinear tap is the pixel itself
            g = px[:, 1:3].astype(np.float32)
            G = np.zeros((2, 2), np.float32)
            for k in range(8):                                     # gradH += g g^T, sequential float sums
                G = (G + np.outer(g[k], g[k]).astype(np.float32)).astype(np.float32)
            sq = (g[:, 0] * g[:, 0] + g[:, 1] * g[:, 1]).astype(np.float32)
            wts = np.sqrt(np.float32(outlierTHSumComponent) / (np.float32(outlierTHSumComponent) + sq)).astype(np.float32)
            cols["host_idx"].append(hi); cols["u"].append(float(x)); cols["v"].append(float(y))
            cols["gradH"].append([G[0, 0], G[1, 0], G[0, 1], G[1, 1]]); cols["color"].append(px[:, 0]); cols["weights"].append(wts)
            cols["idt"].append(1.0 / kf.depths[f][y, x])
    n = len(cols["u"])
    return ImmatureSet(cam, np.array(cols["host_idx"], np.int32), np.array(cols["u"], np.float32), np.array(cols["v"], np.float32),
                       np.array(cols["gradH"], np.float32), np.array(cols["color"], np.float32), np.array(cols["weights"], np.float32),
                       np.full(n, 8 * outlierTH, np.float32), np.zeros(n, np.float32), np.full(n, np.nan, np.float32),
                       np.full(n, 10000.0, np.float32), np.full(n, 5, np.int8), np.full((n, 2), -1, np.float32), np.zeros(n, np.float32),
                       np.array(cols["idt"], np.float64))


def trace_tables(kf: Keyframes, hosts, new_frame, ims: ImmatureSet):
    """Per host: KRKi = K R K^-1, Kt = K t of host->new (float), aff = fromToVecExposure (Mapping.cpp:503-517)."""
    cam = kf.cam
    K = np.array([[cam.fx, 0, cam.cx], [0, cam.fy, cam.cy], [0, 0, 1]], np.float64)
    Ki = np.linalg.inv(K)
    KRKi, Kt, aff = [], [], []
    for f in hosts:
        T = se3_mul(kf.T_true[new_frame], se3_inv(kf.T_true[f]))     # host -> new
        R, t = T12_split(T)
        KRKi.append((K @ R @ Ki).astype(np.float32).reshape(-1)); Kt.append((K @ t).astype(np.float32))
        aff.append(np.array(aff_from_to(1.0, 1.0, kf.aff_true[f][0], kf.aff_true[f][1], kf.aff_true[new_frame][0], kf.aff_true[new_frame][1]), np.float32))
    ims.KRKi, ims.Kt, ims.aff = np.array(KRKi, np.float32), np.array(Kt, np.float32), np.array(aff, np.float32)
    return ims


def activation_tables(kf: Keyframes):
    """Per ordered pair (host + target*n): PRE_RTll row-major, PRE_tTll, PRE_aff_mode (OptimizationClasses.cpp:13-39) as float."""
    n = len(kf.images)
    out = np.zeros((n * n, 14), np.float32)
    for hst in range(n):
        for tg in range(n):
            T = se3_mul(kf.T_true[tg], se3_inv(kf.T_true[hst]))
            R, t = T12_split(T)
            a = aff_from_to(1.0, 1.0, kf.aff_true[hst][0], kf.aff_true[hst][1], kf.aff_true[tg][0], kf.aff_true[tg][1])
            out[hst + tg * n, :9] = R.astype(np.float32).reshape(-1)
            out[hst + tg * n, 9:12] = t.astype(np.float32)
            out[hst + tg * n, 12:14] = np.array(a, np.float32)
    return out
