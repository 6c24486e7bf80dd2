"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/hslam_b200.h declares, and fails LOUDLY (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "hslam_b200.h")
LIB = os.path.join(ROOT, "h-slam_b200", "libhslam_b200.so")


def declared():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hslam_[a-z_A-Z0-9]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "h-slam_b200", "csrc")], stdout=subprocess.DEVNULL)
    return C.CDLL(LIB)


def test_every_declared_symbol_is_exported(lib):
    names = declared()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_header_cites_reference_for_every_entry_point():
    """Each compute entry point names the reference file it replaces."""
    src = open(HDR).read()
    for key in ("Src/Frame.cpp:104-181", "Src/CoarseTracker.cpp:329-485", "Src/CoarseTracker.cpp:267-324",
                "Src/CoarseTracker.cpp:506-683", "OptimizationClasses.cpp:43-233", "EnergyFunctional.cpp:155-175",
                "EnergyFunctional.cpp:203-220", "AccumulatedSCHessian.cpp", "EnergyFunctional.cpp:222-274"):
        assert key in src, key


def test_defaults_match_reference_settings(lib, capi):
    p = capi.Params()
    lib.hslam_params_default(C.byref(p))
    assert (p.huberTH, p.coarseCutoffTH, p.outlierTHSumComponent) == (9.0, 20.0, 2500.0)      # Src/Settings.cpp:68,77,64
    assert (p.scale_xi_rot, p.scale_xi_trans, p.scale_a, p.scale_b) == (1.0, 0.5, 10.0, 1000.0)  # GlobalTypes.h:34-41
    assert list(p.max_tracker_iterations) == [10, 20, 50, 50, 50]                              # CoarseTracker.cpp:518
    assert lib.hslam_abi_version() == 2


def test_fails_loudly_without_a_gpu(lib, capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    rc = lib.hslam_create(None, 0, C.byref(h))
    assert rc == -2                     # HSLAM_ERR_CUDA
    lib.hslam_last_error.restype = C.c_char_p
    assert b"no usable CUDA device" in lib.hslam_last_error()
    with pytest.raises(capi.HslamError):
        capi.Context(0)


def test_product_never_imports_the_oracle():
    """The product path must not route through oracle/ (only tests, smoke and bench's CPU legs may)."""
    pkg = os.path.join(ROOT, "h-slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oraclelib" not in txt and "hslam_oracle" not in txt and "orc_" not in txt, os.path.join(dirpath, f)
