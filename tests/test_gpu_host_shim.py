"""The C++ host mirror (h-slam_b200/host/hslam_host.hpp) driven like the reference's System would drive
it: a C++ program goes Frame::CreateDirPyrs -> EnergyFunctional::{setAdjointsF.., makeIDX, linearizeAll,
applyRes, accumulate*_MT, resubstituteF_MT} -> CoarseTracker::trackNewestCoarse; results are compared with
the oracle (fp64) and with the oracle's host tables (bit-exact)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import rel_fro

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_host_shim.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_host_shim")


def build_exe():
    lib_dir = os.path.join(ROOT, "h-slam_b200")
    if not os.path.exists(os.path.join(lib_dir, "libhslam_b200.so")):
        subprocess.check_call(["make", "-C", os.path.join(lib_dir, "csrc")], stdout=subprocess.DEVNULL)
    hdr = os.path.join(lib_dir, "host", "hslam_host.hpp")
    chdr = os.path.join(ROOT, "include", "hslam_b200.h")   # struct layouts: a stale binary would smash its stack
    if (not os.path.exists(EXE)) or max(os.path.getmtime(SRC), os.path.getmtime(hdr), os.path.getmtime(chdr)) > os.path.getmtime(EXE):
        subprocess.check_call(["g++", "-std=c++11", "-O2", "-Wall", "-o", EXE, SRC, f"-L{lib_dir}", "-lhslam_b200",
                               f"-Wl,-rpath,{lib_dir}"])
    return EXE


def test_host_shim_compiles_against_the_c_abi():
    """CPU-side: the C++ mirror compiles as C++11 and links against the C-ABI library."""
    build_exe()
    out = subprocess.run([EXE], capture_output=True, text=True)
    assert out.returncode == 2 and "usage" in out.stderr


@pytest.mark.gpu
def test_host_shim_end_to_end(tmp_path, O, synth):
    exe = build_exe()
    kf = synth.render_keyframes(320, 240, 4, seed=12)
    win, _ = synth.make_window(points_per_frame=120, seed=40, keyframes=kf)
    levels = 4
    n, P, R = win.n, win.P, win.R
    N = 8 * n + 4
    rng = np.random.default_rng(3)
    x = rng.normal(0, 1e-3, N)
    # reference cloud for the tracker: frame 0 as reference, tracked frame = frame 1
    pyr0, ag0 = synth.build_pyramid_np(win.images[0], levels)
    sel = np.nonzero(win.host_idx == 0)[0]
    su, sv = win.u[sel].astype(np.int32), win.v[sel].astype(np.int32)
    pc = synth.make_coarse_depth_np(levels, 320, 240, su, sv, win.idepth[sel], np.ones(len(sel), np.float32), pyr0)
    fin = tmp_path / "in.bin"
    with open(fin, "wb") as f:
        np.array([320, 240, n, P, R, levels, 0, 0], np.int32).tofile(f)
        np.array([win.cam.fx, win.cam.fy, win.cam.cx, win.cam.cy], np.float32).tofile(f)
        for im in win.images:
            im.astype(np.float32).tofile(f)
        fp = win.framePrior.reshape(n, 8); fd = win.frameDeltaPrior.reshape(n, 8)
        for i in range(n):
            win.T_w2c_eval[i].astype(np.float64).tofile(f)
            win.state[i].astype(np.float64).tofile(f)
            win.state_zero[i].astype(np.float64).tofile(f)
            fp[i].astype(np.float64).tofile(f)
            fd[i].astype(np.float64).tofile(f)
        win.cDeltaF.astype(np.float32).tofile(f)
        win.cPrior.astype(np.float64).tofile(f)
        for p in range(P):
            np.array([win.u[p], win.v[p], win.idepth[p], win.idepth_zero[p], win.priorF[p]], np.float32).tofile(f)
            win.color[p].astype(np.float32).tofile(f)
            win.weights[p].astype(np.float32).tofile(f)
            lst = win.res_index[win.res_begin[p]:win.res_begin[p + 1]]
            np.array([win.host_idx[p], len(lst)], np.int32).tofile(f)
            lst.astype(np.int32).tofile(f)
        np.stack([win.point_idx, win.target_idx], 1).astype(np.int32).tofile(f)
        x.tofile(f)
        for l in range(levels):
            np.array([len(pc[l][0])], np.int32).tofile(f)
            for a in pc[l]:
                a.astype(np.float32).tofile(f)
    fout = tmp_path / "out.bin"
    r = subprocess.run([exe, str(fin), str(fout)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    with open(fout, "rb") as f:
        energy = np.fromfile(f, np.float64, 1)[0]
        newState = np.fromfile(f, np.int8, R)
        sys3 = []
        for _ in range(3):
            H = np.fromfile(f, np.float64, N * N).reshape(N, N).T
            b = np.fromfile(f, np.float64, N)
            sys3.append((H, b))
        same, resInA, resInL, ok = np.fromfile(f, np.int32, 4)
        steps = np.fromfile(f, np.float32, P)
        T = np.fromfile(f, np.float64, 12)
        aff = np.fromfile(f, np.float64, 2)
        lastRes = np.fromfile(f, np.float64, 5)
        precalc = np.fromfile(f, np.float32, n * n * 28)
        adH = np.fromfile(f, np.float64, n * n * 64)
        adT = np.fromfile(f, np.float64, n * n * 64)
        adHTd = np.fromfile(f, np.float32, n * n * 8)
        newFrameTH = np.fromfile(f, np.float32, 1)[0]
    # host tables: same operation order as the oracle's restatement of the reference
    pc_o, adH_o, adT_o, adHTd_o = O.prepare_window(win.cam, win.T_w2c_eval, win.state, win.state_zero, win.ab_exposure)
    # (the mirror's SE3::exp is Rodrigues, the oracle's goes through the unit quaternion as Sophus does:
    #  equal to fp64 round-off, so the fp32 tables agree to the last ulp or so)
    assert np.allclose(precalc, pc_o, rtol=2e-6, atol=2e-5)
    assert np.allclose(adH, adH_o, rtol=1e-13, atol=1e-14) and np.allclose(adT, adT_o, rtol=1e-13, atol=1e-14)
    assert np.allclose(adHTd, adHTd_o, rtol=1e-6, atol=1e-9)
    # oracle on the shim's tables
    import copy
    w2 = copy.copy(win)
    w2.precalc, w2.adHost, w2.adTarget, w2.adHTdeltaF = precalc, adH, adT, adHTd
    dIs = [O.pyramid(im, 1)[0][0] for im in win.images]
    of = O.BAOracle(w2, dIs, prec="f")
    od = O.BAOracle(w2, dIs, prec="d")
    lf = of.linearize(False)
    assert newFrameTH == of.new_frame_energy_th(n - 1)      # System::setNewFrameEnergyTH through the mirror, bit-exact
    of.apply_res()
    od.linearize(False); od.apply_res()
    assert np.array_equal(newState, lf["newState"])
    assert energy == pytest.approx(lf["energy"], rel=1e-12)
    assert same == 1 and resInL == 0
    for k, (H, b) in enumerate(sys3):
        Hd, bd, nres = od.accumulate(k)
        Hf, bf, _ = of.accumulate(k)
        assert rel_fro(H, Hd) <= max(1e-5, 1.5 * rel_fro(Hf, Hd)), k
        assert rel_fro(b, bd) <= max(1e-5, 1.5 * rel_fro(bf, bd)), k
        if k == 0:
            assert resInA == nres
    assert np.array_equal(steps, of.resubstitute(x))
    # tracker: converges to the motion between keyframe 0 and 1
    assert ok == 1
    T01 = O.se3_mul(kf.T_true[1], O.se3_inv(kf.T_true[0]))
    assert np.allclose(O.se3_log(T), O.se3_log(T01), atol=2e-3)
    assert np.all(lastRes[:levels] < 15)
