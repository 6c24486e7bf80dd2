"""The sequence driver (h-slam_b200/sequence.py) on the CPU with the oracle as the tracker back end: float-faithful and
fp64 restatements take the same decisions and give the same trajectory (configs[2] logic without a GPU)."""
import importlib
import sys, os

import numpy as np

sys.path.insert(0, os.path.dirname(__file__))


def test_sequence_driver_float_vs_double_oracle(O):
    from test_gpu_sequence import OracleTracker
    seq = importlib.import_module("h-slam_b200.sequence")
    cam, Ts, frames = seq.render_sequence(160, 120, 14, seed=3)
    f = seq.run_sequence(cam, Ts, frames, 3, lambda c, l, pc: OracleTracker(O, c, l, pc, "f"), kf_every=5, n_points=800)
    d = seq.run_sequence(cam, Ts, frames, 3, lambda c, l, pc: OracleTracker(O, c, l, pc, "d"), kf_every=5, n_points=800)
    assert all(f.ok) and all(d.ok) and f.n_keyframes == 3
    path = np.linalg.norm(np.diff(np.array([T[9:] for T in Ts]), axis=0), axis=1).sum()
    assert f.ate_rmse < 0.05 * path
    assert abs(f.ate_rmse - d.ate_rmse) <= 0.01 * d.ate_rmse + 1e-6


def test_two_stage_sequence_driver_on_the_oracle(O):
    """Tracking AND windowed BA in one loop (h-slam_b200/sequence_ba.py) with the oracle as the back end: the window
    optimisation runs, converges (energy decreases inside a keyframe's GN loop), thresholds move, the trajectory stays
    close to the truth, and the float and fp64 restatements give the same ATE."""
    from seq_backends import OracleBackend
    sb = importlib.import_module("h-slam_b200.sequence_ba")
    cam, Ts, frames = sb.render_sequence(160, 120, 22, seed=3)
    f = sb.run_sequence_ba(cam, Ts, frames, 3, OracleBackend(O, "f"), kf_every=4, pts_per_kf=150, max_kf=4)
    d = sb.run_sequence_ba(cam, Ts, frames, 3, OracleBackend(O, "d"), kf_every=4, pts_per_kf=150, max_kf=4)
    assert all(f.ok) and all(d.ok)
    assert f.n_keyframes == 6 and len(f.ba_iterations) == 5 and all(1 <= i <= 6 for i in f.ba_iterations)
    for en in f.ba_energy:
        assert en[-2] <= en[0] * 1.0001          # GN did not make the window worse (last entry = the fix pass)
    assert any(abs(th[-1] - 512.0) > 1.0 for th in f.thresholds)
    path = np.linalg.norm(np.diff(np.array([T[9:] for T in Ts]), axis=0), axis=1).sum()
    assert f.ate_rmse < 0.05 * path, (f.ate_rmse, path)
    assert abs(f.ate_rmse - d.ate_rmse) <= 0.02 * d.ate_rmse + 1e-6, (f.ate_rmse, d.ate_rmse)
