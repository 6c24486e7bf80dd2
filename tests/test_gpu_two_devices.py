"""Two hslam contexts on two GPUs inside ONE process (the opt-in shared-memory attributes of the top / Schur
kernels are per device, so they are kept per context): both run the whole iteration on the same window and must give
bit-identical systems.  Skipped on a one-GPU box."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_two_contexts_in_one_process(capi, synth, window_small):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    win, _ = window_small
    out = []
    keep = []
    for dev in (0, 1):
        ctx = capi.Context(device=dev)
        frames = [capi.Frame(ctx, im, 1) for im in win.images]
        ef = capi.EnergyFunctional(ctx)
        ef.set_window(win, frames)
        ef.set_topology(win)
        ef.iteration_device()
        Hb = ef.download_Hb()
        out.append([np.array(a) for pair in Hb for a in pair])
        keep.append((ctx, frames, ef))
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)
    for ctx, frames, ef in keep:
        ef.close()
        for f in frames:
            f.close()
        ctx.close()
