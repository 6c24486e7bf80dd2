"""Point-sharded window on 2 GPUs: every rank runs the hot path on its shard through the C ABI, the stitched
fp64 systems are summed with one NCCL all-reduce, and the result must equal the un-sharded window's systems
(stitch is linear in the accumulators).  Needs >= 2 CUDA devices (gpurun --gpus 2); skipped otherwise."""
import importlib
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _DevArray:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3, "strides": None}


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    capi = importlib.import_module("h-slam_b200.capi")
    synth = importlib.import_module("h-slam_b200.synth")
    shard = importlib.import_module("h-slam_b200.shard")
    win, _ = synth.make_window(w=320, h=240, points_per_frame=400, n_frames=5, seed=77)
    ctx = capi.Context(rank)
    frames = [capi.Frame(ctx, im, 1) for im in win.images]
    sub = shard.shard_window(win, rank, world)
    ef = capi.EnergyFunctional(ctx)
    ef.set_window(sub, frames); ef.set_topology(sub); ef.set_add_priors(rank == 0)
    # the communicator is created through the library's own NCCL handle; the id travels over torch.distributed
    capi.use_torch_nccl()
    ids = [capi.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm = capi.NcclComm(ctx, ids[0], world, rank)
    ptr, cnt = ef.iteration_device()
    ef.allreduce(comm)                        # hslam_ba_allreduce: ONE ncclAllReduce(sum, fp64) on the context stream
    got = np.concatenate([np.concatenate([H.T.reshape(-1), b]) for H, b in ef.download_Hb()])   # synchronises
    stats = ef.download_stats()               # energy and resInA/L travelled with the systems
    ef.iteration_allreduce(comm)              # the same through the one-call form
    got2 = np.concatenate([np.concatenate([H.T.reshape(-1), b]) for H, b in ef.download_Hb()])
    assert np.array_equal(got, got2) and ef.download_stats() == stats
    if rank == 0:
        full = capi.EnergyFunctional(ctx)
        full.set_window(win, frames); full.set_topology(win)
        p2, c2 = full.iteration_device()
        ctx.synchronize()
        ref = torch.as_tensor(_DevArray(p2, c2), device=torch.device("cuda", 0)).cpu().numpy()
        fs = full.download_stats()
        assert stats[1:] == fs[1:] and abs(stats[0] - fs[0]) <= 1e-9 * abs(fs[0]), (stats, fs)
        q.put((float(np.linalg.norm(got - ref) / np.linalg.norm(ref)), int(sub.P), int(win.P)))
    comm.close()
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_window_nccl_allreduce():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    err, subP, P = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert 0 < subP < P
    assert err < 1e-6, err      # fp32 chunk partials are grouped differently per shard; fp64 stitch is linear
