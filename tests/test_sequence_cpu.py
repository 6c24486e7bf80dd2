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
