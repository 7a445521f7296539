"""Renderer vocabulary — mirror of crates/raytracing/src/renderer/mod.rs and sampling/mod.rs.

Same names, defaults and meaning as the reference so the parity tests read like its own tests.
"""
from __future__ import annotations

import enum
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _ffi


class AovFlags(enum.IntFlag):
    """renderer/mod.rs:13-47"""
    BEAUTY = 1 << 0
    NORMALS = 1 << 1
    ALBEDO = 1 << 2
    UV_COORDS = 1 << 3
    MIP_LEVEL = 1 << 4
    DEBUG = NORMALS | ALBEDO | UV_COORDS | MIP_LEVEL
    FIRST_HIT_AOVS = NORMALS | ALBEDO | UV_COORDS | MIP_LEVEL
    # debug planes added by this backend (not in the reference): primary-hit ids and depth
    DEBUG_IDS = 1 << 16
    DEBUG_DEPTH = 1 << 17


@dataclass(frozen=True)
class Sampler:
    """sampling/mod.rs:2-10. kind 'independent' | 'stratified'."""
    kind: str = "independent"
    jitter: bool = True
    x_strata: int = 1
    y_strata: int = 1

    @staticmethod
    def independent() -> "Sampler":
        return Sampler("independent")

    @staticmethod
    def stratified(jitter: bool, x_strata: int, y_strata: int) -> "Sampler":
        return Sampler("stratified", jitter, x_strata, y_strata)


@dataclass
class RaytracerSettings:
    """renderer/mod.rs:84-117 (defaults: depth 8, accumulate, 4 light samples, 32 spp, seed None)."""
    max_ray_depth: int = 8
    accumulate_bounces: bool = True
    light_sample_count: int = 4
    samples_per_pixel: int = 32
    seed: Optional[int] = None
    sampler: Sampler = field(default_factory=Sampler.independent)
    outputs: AovFlags = AovFlags.BEAUTY
    antialias_primary_rays: bool = True
    antialias_secondary_rays: bool = True

    def to_c(self) -> _ffi.Settings:
        s = _ffi.Settings()
        s.max_ray_depth = self.max_ray_depth
        s.accumulate_bounces = int(self.accumulate_bounces)
        s.light_sample_count = self.light_sample_count
        s.samples_per_pixel = self.samples_per_pixel
        s.has_seed = 0 if self.seed is None else 1
        s.seed = 0 if self.seed is None else int(self.seed)
        s.sampler_kind = _ffi.SAMPLER_STRATIFIED if self.sampler.kind == "stratified" else _ffi.SAMPLER_INDEPENDENT
        s.stratified_jitter = int(self.sampler.jitter)
        s.x_strata = self.sampler.x_strata
        s.y_strata = self.sampler.y_strata
        s.outputs = int(self.outputs)
        s.antialias_primary_rays = int(self.antialias_primary_rays)
        s.antialias_secondary_rays = int(self.antialias_secondary_rays)
        return s


@dataclass
class RenderOutput:
    """renderer/mod.rs:49-73 — planes are None unless their AovFlags bit was requested; row-major
    [H, W, C] float32 arrays (idx = y*W + x)."""
    width: int
    height: int
    beauty: Optional[np.ndarray] = None
    normals: Optional[np.ndarray] = None
    albedo: Optional[np.ndarray] = None
    uv: Optional[np.ndarray] = None
    mip_level: Optional[np.ndarray] = None
    debug_ids: Optional[np.ndarray] = None    # [H, W, 2] uint32 (geom_id, prim_id), backend extension
    debug_depth: Optional[np.ndarray] = None  # [H, W] float32, backend extension

    _PLANES = (("beauty", AovFlags.BEAUTY, 3, np.float32), ("normals", AovFlags.NORMALS, 3, np.float32),
               ("albedo", AovFlags.ALBEDO, 3, np.float32), ("uv", AovFlags.UV_COORDS, 2, np.float32),
               ("mip_level", AovFlags.MIP_LEVEL, 1, np.float32), ("debug_ids", AovFlags.DEBUG_IDS, 2, np.uint32),
               ("debug_depth", AovFlags.DEBUG_DEPTH, 1, np.float32))

    @classmethod
    def allocate(cls, width: int, height: int, outputs: AovFlags, lib=None) -> "RenderOutput":
        """The planes the backend fills. With `lib` (the loaded library) they are page-locked buffers from rtcuda_host_alloc, which
        the GPU writes without a staging copy and which go back to the library's cache when the arrays are garbage-collected;
        without it, plain numpy arrays (zero-filled)."""
        out = cls(width, height)
        for name, flag, ch, dt in cls._PLANES:
            if outputs & flag:
                shape = (height, width, ch) if ch > 1 else (height, width)
                arr = _ffi.host_array(lib, shape, dt) if lib is not None else None
                setattr(out, name, arr if arr is not None else np.zeros(shape, dtype=dt))
        return out

    def to_c(self) -> _ffi.Outputs:
        o = _ffi.Outputs()
        o.width, o.height = self.width, self.height
        for name, _flag, _ch, _dt in self._PLANES:
            arr = getattr(self, name)
            setattr(o, name, arr.ctypes.data if arr is not None else None)
        return o


@dataclass
class SinglePixelOutput:
    """renderer/mod.rs:75-82"""
    sample_index: int
    hit: bool
    uv: tuple
    normal: tuple
    radiance: tuple
