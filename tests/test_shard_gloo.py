"""N>1 host logic on CPU: point-sharded window + one all-reduce(sum) of the stitched fp64 systems
(world_size 2, gloo).  The per-shard systems come from the oracle here (no GPU in this container); the GPU
path uses the same shard_window()/priors-on-rank-0 logic with hslam_ba_* + NCCL (tests/test_gpu_shard.py)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    synth = importlib.import_module("h-slam_b200.synth")
    shard = importlib.import_module("h-slam_b200.shard")
    from oracle import oraclelib as O
    win, _ = synth.make_window(w=320, h=240, points_per_frame=90, n_frames=4, seed=321)
    dIs = [O.pyramid(im, 1)[0][0] for im in win.images]
    sub = shard.shard_window(win, rank, world)
    if rank != 0:   # priors only once (AccumulatedTopHessian.cpp:268-279)
        sub.cPrior = np.zeros(4); sub.framePrior = np.zeros_like(win.framePrior)
    ba = O.BAOracle(sub, dIs, prec="d")
    lin = ba.linearize(False); ba.apply_res()
    blocks = []
    for which in (0, 1, 2):
        H, b, _ = ba.accumulate(which)
        blocks += [H.reshape(-1), b]
    buf = torch.from_numpy(np.concatenate(blocks))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    e = torch.tensor([lin["energy"]], dtype=torch.float64)
    dist.all_reduce(e)
    if rank == 0:
        full = O.BAOracle(win, dIs, prec="d")
        lf = full.linearize(False); full.apply_res()
        ref = []
        for which in (0, 1, 2):
            H, b, _ = full.accumulate(which)
            ref += [H.reshape(-1), b]
        ref = np.concatenate(ref)
        err = np.linalg.norm(buf.numpy() - ref) / np.linalg.norm(ref)
        q.put((err, float(e.item()), lf["energy"], sub.P, win.P))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_window_allreduce_equals_whole_window():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    err, e_sum, e_full, subP, P = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-12, err
    assert e_sum == pytest.approx(e_full, rel=1e-12)
    assert 0 < subP < P


def test_partition_is_contiguous_and_complete():
    sys.path.insert(0, ROOT)
    shard = importlib.import_module("h-slam_b200.shard")
    host = np.array([2, 0, 1, 0, 2, 1, 1, 0, 2, 2, 0])
    for world in (1, 2, 3, 4, 8):
        order, bounds = shard.point_ranges(host, world)
        assert bounds[0] == 0 and bounds[-1] == len(host)
        assert sorted(order.tolist()) == list(range(len(host)))
        assert np.all(np.diff(host[order]) >= 0)
        sizes = np.diff(bounds)
        assert sizes.max() - sizes.min() <= 1
    assert shard.assign_windows(64, 3, 8) == list(range(24, 32))
    assert sum(len(shard.assign_windows(10, r, 4)) for r in range(4)) == 10
