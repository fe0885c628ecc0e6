"""The N>1 path on CPU: world_size-2 gloo.  Each rank renders its strided share of the
samples (here with the oracle's float mirror standing in for a GPU: same fixed-point
contract), the int64 accumulation buffers are summed with one reduce to rank 0 -- exactly the
collective bench.py issues over NCCL -- and the result equals the single-rank render bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _share(spp, rank, world):
    """samples of a frame owned by `rank`: global indices rank, rank+world, ... (bench.py)"""
    return spp // world + (1 if rank < spp % world else 0)


def _worker(rank, world, port, spp, out_path):
    sys.path.insert(0, ROOT)
    import oracle as orc
    from rtxplay_b200 import scenes
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spheres = scenes.book1(seed=1)
    tab, _ = scenes.table(spheres, "analytic")
    cam = orc.camera_f32((13, 2, 3), (0, 0, 0), (0, 1, 0), 20., 1.5, .1, 10.)
    w, h = 48, 32
    mine = orc.render(orc.F32_PCG, tab, cam, w, h, _share(spp, rank, world), 50, sample0=rank, sample_stride=world, threads=2)
    acc = np.concatenate([mine["fix"].astype(np.int64), mine["rpp"].astype(np.int64)[..., None]], axis=2)  # r,g,b,segments
    t = torch.from_numpy(acc.copy())
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(out_path, t.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("spp", [6, 7])
def test_two_rank_reduce_equals_single_render(tmp_path, spp):
    import oracle as orc
    from rtxplay_b200 import scenes
    orc.build()
    out = str(tmp_path / "sum.npy")
    port = 29500 + (os.getpid() % 1000) + spp
    mp.spawn(_worker, args=(2, port, spp, out), nprocs=2, join=True)
    got = np.load(out)
    spheres = scenes.book1(seed=1)
    tab, _ = scenes.table(spheres, "analytic")
    cam = orc.camera_f32((13, 2, 3), (0, 0, 0), (0, 1, 0), 20., 1.5, .1, 10.)
    full = orc.render(orc.F32_PCG, tab, cam, 48, 32, spp, 50, threads=2)
    assert np.array_equal(got[..., :3].astype(np.uint64), full["fix"])
    assert np.array_equal(got[..., 3].astype(np.uint32), full["rpp"])


def test_sample_shares_cover_the_frame():
    for spp in (1, 7, 500, 4096):
        for world in (1, 2, 4, 8):
            assert sum(_share(spp, r, world) for r in range(world)) == spp
            idx = sorted(r + k * world for r in range(world) for k in range(_share(spp, r, world)))
            assert idx == list(range(spp))
