"""Multi-GPU rendering: the scene is replicated on every GPU, square image tiles are dealt round-robin to the
ranks (tile i -> rank i % world, the tile grid of create_render_jobs, crates/raytracing-cpu/src/lib.rs:481-504;
64x64 like the reference's RenderTile unless `tile_size` asks for a finer deal), and the tile-disjoint frames are
combined with ONE collective per frame over NCCL / NVLink: every rank sends only the pixels it owns (all planes packed
side by side) and rank 0 copies them into its frame (`gather_tiles`) — bit-exact, 1/world of a frame per rank on the
wire. (`reduce_planes`, the sum of the zero-filled full frames, gives the same bits with world times the traffic; it
serves the sample partition, whose partial sums really have to be added.)

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


def tile_row_stride(tiles_x: int, world: int) -> int:
    """Tile (tx, ty) belongs to rank (ty * stride + tx) % world: the row-major round-robin of the tile index, with one idle
    slot per tile row when the row length is a multiple of the world size (else every rank would own whole tile columns)."""
    return tiles_x + (1 if world > 1 and tiles_x % world == 0 else 0)


def tile_owner_map(width: int, height: int, world: int, tile: int = 64) -> np.ndarray:
    """[H, W] array of the rank that renders each pixel."""
    tiles_x = (width + tile - 1) // tile
    ty, tx = np.meshgrid(np.arange(height) // tile, np.arange(width) // tile, indexing="ij")
    return ((ty * tile_row_stride(tiles_x, world) + tx) % max(1, world)).astype(np.int32)


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


def owned_pixel_indices(width: int, height: int, world: int, tile: int = 64) -> list:
    """Per rank, the flat (y * W + x) indices of the pixels it renders, ascending (numpy int64)."""
    own = tile_owner_map(width, height, world, tile).reshape(-1)
    return [np.flatnonzero(own == r) for r in range(max(1, world))]


def gather_tiles(planes: dict, owned: list, counts: list, rank: int, world: int, group=None, dst: int = 0) -> None:
    """Tile partition: every rank sends ONLY the pixels it owns (all planes packed side by side, one collective per frame)
    and `dst` copies them into its full-frame planes — the frames are tile-disjoint, so this is the same result as summing
    them (`reduce_planes`), bit for bit, with 1/world of the traffic per rank. `owned[r]` = index tensor (same device as the
    planes) of rank r's pixels, `counts[r]` its length; ranks other than `dst` only need their own `owned` entry."""
    import torch
    import torch.distributed as dist
    names = sorted(planes)
    if not names or world <= 1:
        return
    chans = [planes[n].shape[-1] for n in names]
    maxc = max(counts)
    first = planes[names[0]]
    send = torch.zeros((maxc, sum(chans)), dtype=first.dtype, device=first.device)
    c0 = 0
    for n, c in zip(names, chans):
        send[:counts[rank], c0:c0 + c] = planes[n].reshape(-1, c).index_select(0, owned[rank])
        c0 += c
    recv = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, recv, dst=dst, group=group)
    if rank != dst:
        return
    for r in range(world):
        if r == dst:
            continue
        c0 = 0
        for n, c in zip(names, chans):
            planes[n].reshape(-1, c).index_copy_(0, owned[r], recv[r][:counts[r], c0:c0 + c])
            c0 += c


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
        self._owned = None
        self.tile = backend_kw.get("tile_size", 0) or 64

    def planes_for(self, outputs: AovFlags) -> dict:
        torch = self.torch
        planes = {}
        for name, flag, ch in PLANES:
            if outputs & flag:
                key = (name, ch)
                if key not in self._planes:   # (the library zero-fills the pixels of other ranks itself)
                    self._planes[key] = torch.empty((self.height, self.width, ch), dtype=torch.float32, device=self.device)
                planes[name] = self._planes[key]
        return planes

    def render_local(self, settings: RaytracerSettings) -> dict:
        """This rank's share of the frame in HBM planes. partition="samples": the beauty plane holds this rank's sample
        range already multiplied by 1 / spp (so that the reduce yields the mean); AOVs come from rank 0 alone."""
        planes = self.planes_for(AovFlags(settings.outputs))
        # The library renders on its own non-blocking stream: whatever torch still has in flight on these planes (the previous
        # frame's index_select / gather reads, the allocation of a fresh plane) must be done before the library writes them;
        # the library itself synchronises its stream before it returns, so torch may read the planes right after.
        self.torch.cuda.current_stream(self.device).synchronize()
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

    def combine(self, planes: dict, group=None) -> None:
        """Bring the ranks' shares together on rank 0: a gather of the owned pixels for the tile partition (1/world of a
        frame per rank on the wire), a sum-reduce of the partial sums for the sample partition."""
        if self.world <= 1:
            return
        if self.partition == "samples":
            reduce_planes(planes, group=group, dst=0)
            return
        if self._owned is None:   # on the device: a one-shot call must not spend tens of ms of numpy on an index table
            t = self.torch
            tiles_x = (self.width + self.tile - 1) // self.tile
            ty = t.arange(self.height, device=self.device) // self.tile
            tx = t.arange(self.width, device=self.device) // self.tile
            owner = ((ty[:, None] * tile_row_stride(tiles_x, self.world) + tx[None, :]) % self.world).reshape(-1)
            self._counts = t.bincount(owner, minlength=self.world).tolist()
            self._owned = [t.nonzero(owner == r).reshape(-1) if (r == self.rank or self.rank == 0) else None for r in range(self.world)]
        gather_tiles(planes, self._owned, self._counts, self.rank, self.world, group=group, dst=0)

    def render(self, settings: RaytracerSettings, group=None) -> Optional[RenderOutput]:
        planes = self.render_local(settings)
        self.combine(planes, group=group)
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
    scene to its GPU and builds the BVH there, renders its tiles into HBM planes, ONE NCCL collective per frame (`combine`),
    and rank 0 copies the frame to the host (None on the other ranks)."""
    dr = DistributedRenderer(scene, rank, world, device_id=device_id, **kw)
    try:
        return dr.render(settings, group=group)
    finally:
        dr.close()
