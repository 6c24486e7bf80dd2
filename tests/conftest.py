import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module("h-slam_b200.synth")


@pytest.fixture(scope="session")
def capi():
    return importlib.import_module("h-slam_b200.capi")


@pytest.fixture(scope="session")
def O():
    from oracle import oraclelib
    return oraclelib


@pytest.fixture(scope="session")
def tracking_pair(synth):
    """Config 1 at full size (640x480, 5 levels, ~28k level-0 points)."""
    return synth.make_tracking_pair()


@pytest.fixture(scope="session")
def tracking_pair_small(synth):
    return synth.make_tracking_pair(w=320, h=240, levels=4, n_seed_points=1500, seed=7)


@pytest.fixture(scope="session")
def window_small(synth):
    """Config 2 shape at reduced point count: 7 KF x 300 points."""
    return synth.make_window(points_per_frame=300, frac_linearized=0.3)


@pytest.fixture(scope="session")
def ctx(capi):
    c = capi.Context(0)
    yield c
    c.close()


def rel_fro(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
