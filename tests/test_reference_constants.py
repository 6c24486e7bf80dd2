"""Constants that the restatements (oracle, synthetic inputs, CUDA defaults) copy from the reference are checked
against the reference's own SOURCE TEXT: the residual pattern, the state scales and the tracing / robust-cost
settings.  Runs wherever /root/reference is mounted (this container); skipped elsewhere (the GPU box has no copy).
This pins the constants, not the arithmetic — the oracle's algorithms stay "parity unpinned" (DESIGN.md §5)."""
import ctypes as C
import importlib
import os
import re

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not mounted")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _read(rel):
    return open(os.path.join(REF, rel), errors="replace").read()


def _setting(src, name):
    m = re.search(r"\b(?:float|int|double)\s+" + re.escape(name) + r"\s*=\s*([^;]+);", src)
    assert m, name
    expr = re.sub(r"(?<=[0-9.])f\b", "", m.group(1))
    assert re.fullmatch(r"[0-9.eE+\-*/ ()]+", expr), expr   # plain arithmetic only
    return float(eval(expr))


def test_residual_pattern_matches_GlobalTypes():
    gt = _read("Include/GlobalTypes.h")
    chosen = int(re.search(r"#define\s+ChosenPattern\s+(\d+)", gt).group(1))
    body = gt[gt.index("staticPattern {"):gt.index("staticPatternNum")]
    # every pattern is one "{{x,y}, ... }}" group; take the ChosenPattern-th and its staticPatternNum entries
    groups = re.findall(r"\{\s*\{.*?\}\s*\}", body, re.S)
    nums = [int(v) for v in re.findall(r"\d+", gt[gt.index("staticPatternNum[10]"):gt.index("staticPatternPadding")].split("=")[1])]
    assert len(groups) == 10 and len(nums) == 10 and nums[chosen] == 8
    pairs = re.findall(r"\{\s*([+-]?\d+)\s*,\s*([+-]?\d+)\s*\}", groups[chosen])
    pat = np.array(pairs[:8], dtype=np.int32)
    assert (np.array(pairs[8:], dtype=np.int32) == -100).all()   # exactly 8 used
    synth = importlib.import_module("h-slam_b200.synth")
    assert np.array_equal(pat, synth.PATTERN)
    cu = open(os.path.join(ROOT, "h-slam_b200", "csrc", "ba_kernels.cuh")).read()
    m = re.search(r"c_pat\[8\]\[2\]\s*=\s*\{(.*?)\};", cu, re.S)
    dev = np.array(re.findall(r"\{\s*(-?\d+)\s*,\s*(-?\d+)\s*\}", m.group(1)), dtype=np.int32)
    assert np.array_equal(pat, dev)


def test_state_scales_match_GlobalTypes():
    gt = _read("Include/GlobalTypes.h")
    synth = importlib.import_module("h-slam_b200.synth")
    for name in ("SCALE_XI_ROT", "SCALE_XI_TRANS", "SCALE_F", "SCALE_C", "SCALE_A", "SCALE_B"):
        ref = float(re.search(r"#define\s+" + name + r"\s+([0-9.]+)f", gt).group(1))
        assert getattr(synth, name) == ref, name
    assert float(re.search(r"#define\s+SCALE_IDEPTH\s+([0-9.]+)f", gt).group(1)) == 1.0


def test_trace_and_robust_settings_match_Settings_cpp():
    st = _read("Src/Settings.cpp")
    gt = _read("Include/GlobalTypes.h")
    capi = importlib.import_module("h-slam_b200.capi")
    L = capi.load()   # host-only entry points: no device needed
    s = capi.TraceSettings()
    L.hslam_trace_settings_default(C.byref(s))
    want = dict(maxPixSearch="setting_maxPixSearch", slackInterval="setting_trace_slackInterval", stepsize="setting_trace_stepsize",
                minImprovementFactor="setting_trace_minImprovementFactor", GNThreshold="setting_trace_GNThreshold",
                extraSlackOnTH="setting_trace_extraSlackOnTH", huberTH="setting_huberTH",
                minTraceTestRadius="setting_minTraceTestRadius", GNIterations="setting_trace_GNIterations")
    for field, name in want.items():
        assert np.float32(getattr(s, field)) == np.float32(_setting(st, name)), field
    # hslam_params_default: robust-cost thresholds and the state scales the kernels apply
    p = capi.Params()
    L.hslam_params_default(C.byref(p))
    assert p.huberTH == _setting(st, "setting_huberTH")
    assert p.coarseCutoffTH == _setting(st, "setting_coarseCutoffTH")
    assert p.outlierTHSumComponent == _setting(st, "setting_outlierTHSumComponent")
    for field, name in (("scale_xi_rot", "SCALE_XI_ROT"), ("scale_xi_trans", "SCALE_XI_TRANS"), ("scale_a", "SCALE_A"),
                        ("scale_b", "SCALE_B"), ("scale_f", "SCALE_F"), ("scale_c", "SCALE_C"), ("scale_idepth", "SCALE_IDEPTH")):
        ref = float(re.search(r"#define\s+" + name + r"\s+([0-9.]+)f", gt).group(1))
        assert getattr(p, field) == np.float32(ref), field
    # System::setNewFrameEnergyTH settings (FullSystemOptimize.cpp:80-91)
    for field, name in (("frameEnergyTHN", "setting_frameEnergyTHN"), ("frameEnergyTHFacMedian", "setting_frameEnergyTHFacMedian"),
                        ("frameEnergyTHConstWeight", "setting_frameEnergyTHConstWeight"), ("overallEnergyTHWeight", "setting_overallEnergyTHWeight")):
        assert np.float32(getattr(p, field)) == np.float32(_setting(st, name)), field
    fso = _read("Src/FullSystemOptimize.cpp")
    assert "12*12*patternNum" in fso and "26.0f*setting_frameEnergyTHConstWeight" in fso and "setNewFrameEnergyTH();" in fso
    # values the synthetic generators hard-code
    assert _setting(st, "setting_outlierTH") == 144.0 and _setting(st, "setting_idepthFixPrior") == 2500.0
