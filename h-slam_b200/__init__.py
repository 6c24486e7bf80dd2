"""h-slam_b200 — B200-native photometric Gauss-Newton hot path of H-SLAM.

The directory name carries a hyphen (it follows the reference project's name), so import it with
    importlib.import_module("h-slam_b200")
Sub-modules:
    capi   ctypes binding of include/hslam_b200.h + Python mirrors of CoarseTracker / EnergyFunctional
    synth  synthetic rendered workloads (configs of BASELINE.json)
The compute path lives in csrc/ (hand-written sm_100a CUDA behind an extern "C" ABI, built in-tree
into libhslam_b200.so).  There is no CPU fallback.
"""
import importlib as _il

__all__ = ["capi", "synth"]


def __getattr__(name):
    if name in __all__:
        return _il.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
