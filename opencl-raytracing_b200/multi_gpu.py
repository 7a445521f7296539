"""Multi-GPU rendering: the scene is replicated on every GPU, square image tiles are dealt round-robin to the
ranks (tile i -> rank i % world, the tile grid of create_render_jobs, crates/raytracing-cpu/src/lib.rs:481-504;
64x64 like the reference's RenderTile unless `tile_size` asks for a finer deal), and the tile-disjoint frames are
combined with ONE sum-reduce of each plane to rank 0 over NCCL / NVLink. Pixels a rank does not own are exactly 0
in its frame, so the sum is a gather and is bit-exact.

Alternate partition (small rasters / high spp): `partition="samples"` gives rank r the sample range
[r * spp / world, (r + 1) * spp / world) of EVERY pixel; the un-normalised sums are reduced and rank 0 multiplies by
1 / spp. Same streams per (pixel, sample), but the float sum over samples is associated per rank, so the frame
matches the single-GPU one to rounding (not bit for bit).

The reference has no multi-device path; its analogue is the CPU thread pool popping tiles from a queue
(lib.rs:706-805). Every (pixel, sample) stream is a pure function of (seed, x, y, sample) (sample.rs:69-87),
so the partition does not change any pixel.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .backend import CudaBackendSettings, CudaRenderer
from .renderer import AovFlags, RaytracerSettings, RenderOutput


def tile_owner_map(width: int, height: int, world: int, tile: int = 64) -> np.ndarray:
    """[H, W] array of the rank that renders each pixel."""
    tiles_x = (width + tile - 1) // tile
    ty, tx = np.meshgrid(np.arange(height) // tile, np.arange(width) // tile, indexing="ij")
    return ((ty * tiles_x + tx) % max(1, world)).astype(np.int32)


def backend_settings_for_rank(rank: int, world: int, device_id: Optional[int] = None, **kw) -> CudaBackendSettings:
    return CudaBackendSettings(device_id=rank if device_id is None else device_id, tile_rank=rank, tile_world=world, **kw)


def sample_range_for_rank(spp: int, rank: int, world: int) -> tuple:
    """[lo, hi) of the samples rank `rank` traces under partition="samples" (contiguous, sizes differ by at most 1)."""
    return (rank * spp) // world, ((rank + 1) * spp) // world


PLANES = (("beauty", AovFlags.BEAUTY, 3), ("normals", AovFlags.NORMALS, 3), ("albedo", AovFlags.ALBEDO, 3),
          ("uv", AovFlags.UV_COORDS, 2), ("mip_level", AovFlags.MIP_LEVEL, 1), ("debug_depth", AovFlags.DEBUG_DEPTH, 1))


def reduce_planes(planes: dict, group=None, dst: int = 0) -> None:
    """Sum-reduce every float plane (torch tensors, CUDA for NCCL or CPU for gloo) to rank `dst`."""
    import torch.distributed as dist
    for name in sorted(planes):
        dist.reduce(planes[name], dst=dst, op=dist.ReduceOp.SUM, group=group)


class DistributedRenderer:
    """One process per GPU (torchrun): rank r owns tiles r, r+world, ...; `render()` returns the full
    RenderOutput on rank 0 (None elsewhere). Frames stay in HBM between the render and the NCCL reduce
    (rtcuda_render_device)."""

    def __init__(self, scene, rank: int, world: int, device_id: Optional[int] = None, partition: str = "tiles", **backend_kw):
        import torch
        assert partition in ("tiles", "samples")
        self.torch = torch
        self.rank, self.world, self.partition = rank, world, partition
        self.device = torch.device("cuda", rank if device_id is None else device_id)
        if partition == "samples":   # every rank owns every pixel
            bs = CudaBackendSettings(device_id=self.device.index, **backend_kw)
        else:
            bs = backend_settings_for_rank(rank, world, self.device.index, **backend_kw)
        self.renderer = CudaRenderer(scene, bs)
        self.width, self.height = self.renderer.width, self.renderer.height
        self._planes = {}

    def planes_for(self, outputs: AovFlags) -> dict:
        torch = self.torch
        planes = {}
        for name, flag, ch in PLANES:
            if outputs & flag:
                key = (name, ch)
                if key not in self._planes:
                    self._planes[key] = torch.zeros((self.height, self.width, ch), dtype=torch.float32, device=self.device)
                planes[name] = self._planes[key]
        return planes

    def render_local(self, settings: RaytracerSettings) -> dict:
        """This rank's share of the frame in HBM planes. partition="samples": the beauty plane holds this rank's sample
        range already multiplied by 1 / spp (so that the reduce yields the mean); AOVs come from rank 0 alone."""
        planes = self.planes_for(AovFlags(settings.outputs))
        if self.partition == "samples" and self.world > 1:
            lo, hi = sample_range_for_rank(settings.samples_per_pixel, self.rank, self.world)
            if "beauty" in planes:
                if hi > lo:
                    self.renderer.render_samples_device(settings, lo, hi, planes["beauty"].data_ptr())
                    planes["beauty"].mul_(self.torch.tensor(1.0 / settings.samples_per_pixel, dtype=self.torch.float32))
                else:
                    planes["beauty"].zero_()
            aov = {k: v for k, v in planes.items() if k != "beauty"}
            if aov:
                if self.rank == 0:
                    import dataclasses
                    st = dataclasses.replace(settings, outputs=AovFlags(settings.outputs) & ~AovFlags.BEAUTY)
                    self.renderer.render_device(st, {k: v.data_ptr() for k, v in aov.items()})
                else:
                    for v in aov.values():
                        v.zero_()
            return planes
        self.renderer.render_device(settings, {k: v.data_ptr() for k, v in planes.items()})
        return planes

    def render(self, settings: RaytracerSettings, group=None) -> Optional[RenderOutput]:
        planes = self.render_local(settings)
        if self.world > 1:
            reduce_planes(planes, group=group, dst=0)
        if self.rank != 0:
            return None
        out = RenderOutput(self.width, self.height)
        for name, t in planes.items():
            arr = t.cpu().numpy()
            setattr(out, name, arr[..., 0] if arr.shape[-1] == 1 else arr)
        return out

    def close(self):
        self.renderer.close()


def render_distributed(scene, settings: RaytracerSettings, rank: int, world: int, device_id: Optional[int] = None, group=None,
                       **kw) -> Optional[RenderOutput]:
    """The one-shot call of a torchrun job (the multi-GPU analogue of `render(scene, settings)`): every rank uploads the
    scene to its GPU and builds the BVH there, renders its tiles into HBM planes, ONE NCCL sum-reduce per plane, and rank 0
    copies the frame to the host (None on the other ranks)."""
    dr = DistributedRenderer(scene, rank, world, device_id=device_id, **kw)
    try:
        return dr.render(settings, group=group)
    finally:
        dr.close()
