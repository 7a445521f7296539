"""world_size-2 gloo test of the multi-GPU host logic (tile ownership, the gather / sum-reduce of the frame, the sample-range partition),
with the CPU kernel-body harness standing in for the device render of each rank."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_scene


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "hostsim")):
        sys.path.insert(0, p)
    import raytracing_cuda as rc
    import hostsim_py
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = load_scene("cb", 160, 130)
    A = rc.AovFlags
    st = rc.RaytracerSettings(outputs=A.BEAUTY | A.NORMALS, samples_per_pixel=2, light_sample_count=1)
    bs = rc.multi_gpu.backend_settings_for_rank(rank, world)
    out, _ = hostsim_py.render(sc, st, tile_rank=bs.tile_rank, tile_world=bs.tile_world)
    planes = {"beauty": torch.from_numpy(out.beauty.copy()), "normals": torch.from_numpy(out.normals.copy())}
    gathered = {k: v.clone() for k, v in planes.items()}
    rc.multi_gpu.reduce_planes(planes, dst=0)
    # the collective the tile partition uses: only the owned pixels travel
    idx = rc.multi_gpu.owned_pixel_indices(160, 130, world)
    rc.multi_gpu.gather_tiles(gathered, [torch.from_numpy(i) for i in idx], [len(i) for i in idx], rank, world, dst=0)
    # the alternate partition: every rank renders a sample range of every pixel, the un-normalised sums are reduced
    st4 = rc.RaytracerSettings(samples_per_pixel=5, light_sample_count=1)
    lo, hi = rc.multi_gpu.sample_range_for_rank(5, rank, world)
    part, _ = hostsim_py.render(sc, st4, sample_range=(lo, hi))
    sums = {"beauty": torch.from_numpy(part.beauty.copy())}
    rc.multi_gpu.reduce_planes(sums, dst=0)
    if rank == 0:
        whole, _ = hostsim_py.render(sc, st4)
        ret["samples_close"] = bool(np.allclose(sums["beauty"].numpy() * np.float32(1.0 / 5), whole.beauty, rtol=2e-6, atol=1e-7))
        one, _ = hostsim_py.render(sc, st4, sample_range=(0, 5))
        ret["full_range_exact"] = bool(np.array_equal(one.beauty * np.float32(1.0 / 5), whole.beauty))
    if rank == 0:
        full, _ = hostsim_py.render(sc, st)
        ret["beauty_equal"] = bool(np.array_equal(planes["beauty"].numpy(), full.beauty))
        ret["normals_equal"] = bool(np.array_equal(planes["normals"].numpy(), full.normals))
        ret["gather_equal"] = bool(np.array_equal(gathered["beauty"].numpy(), full.beauty) and np.array_equal(gathered["normals"].numpy(), full.normals))
        owner = rc.multi_gpu.tile_owner_map(160, 130, world)
        ret["mine_only"] = bool((out.beauty[owner != 0] == 0).all())
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_tile_partition_and_reduce(hostsim):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["beauty_equal"] and ret["normals_equal"] and ret["mine_only"] and ret["gather_equal"] and ret["samples_close"] and ret["full_range_exact"]
