"""Builtin test scenes — the data of crates/raytracing/src/scene/test_scenes/mod.rs restated, plus the
synthetic procedural mesh of BASELINE config C5.

`all_test_scenes()` keeps the reference's names and per-scene settings (test_scenes/mod.rs:605-692).
`environment_light` needs `lake_pier_1k.exr`, which the reference repo does not ship
(.MISSING_LARGE_BLOBS); a caller may pass its own lat-long f32 image instead.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

from . import _ffi
from .geometry import Transform, f32, vec3
from .renderer import AovFlags, RaytracerSettings, Sampler
from .scene import Camera, Light, Material, Mesh, Scene, SceneBuilder, Sphere, Texture, mesh_from_ply_bytes


def _radians(deg: float) -> float:
    return float(f32(f32(deg) * f32(math.pi / 180.0)))  # f32::to_radians


def make_plane(a, b, c, d, normal) -> Mesh:
    """test_scenes/mod.rs:29-54"""
    v = np.array([a, b, c, d], dtype=f32)
    n = np.array([normal] * 4, dtype=f32)
    return Mesh(v, np.array([[0, 1, 2], [2, 3, 0]], dtype=np.uint32), n, None)


def make_cube(side: float) -> Mesh:
    """test_scenes/mod.rs:56-147: 6 faces (+X,-X,+Y,-Y,+Z,-Z), 4 duplicated vertices each."""
    h = side / 2.0
    faces = [
        ([(h, -h, -h), (h, h, -h), (h, h, h), (h, -h, h)], (1, 0, 0)),
        ([(-h, h, -h), (-h, -h, -h), (-h, -h, h), (-h, h, h)], (-1, 0, 0)),
        ([(h, h, -h), (-h, h, -h), (-h, h, h), (h, h, h)], (0, 1, 0)),
        ([(-h, -h, -h), (h, -h, -h), (h, -h, h), (-h, -h, h)], (0, -1, 0)),
        ([(-h, -h, h), (h, -h, h), (h, h, h), (-h, h, h)], (0, 0, 1)),
        ([(h, -h, -h), (-h, -h, -h), (-h, h, -h), (h, h, -h)], (0, 0, -1)),
    ]
    verts, normals, tris = [], [], []
    for quad, n in faces:
        base = len(verts)
        verts += quad
        normals += [n] * 4
        tris += [(base, base + 1, base + 2), (base, base + 2, base + 3)]
    return Mesh(np.array(verts, dtype=f32), np.array(tris, dtype=np.uint32), np.array(normals, dtype=f32), None)


def _white_diffuse(b: SceneBuilder) -> int:
    white = b.add_constant_texture((1.0, 1.0, 1.0, 1.0))
    return b.add_material(Material(_ffi.MATERIAL_DIFFUSE, albedo=white))


def sphere_scene() -> Scene:
    """test_scenes/mod.rs:150-176 — BASELINE config C1."""
    b = SceneBuilder()
    mat = _white_diffuse(b)
    b.add_shape_at_position(Sphere((0.0, 0.0, 0.0), 1.0), mat, (0.0, 0.0, -3.0))
    b.add_camera(Camera.lookat_camera_perspective((0, 0, 0), (0, 0, -3), (0, 1, 0), False, _radians(45.0), 400, 400))
    return b.build()


def cube_scene() -> Scene:
    b = SceneBuilder()
    mat = _white_diffuse(b)
    b.add_shape_at_position(make_cube(1.0), mat, (0.0, 0.0, -3.0))
    b.add_camera(Camera.lookat_camera_perspective((1.0, 0.75, -1.0), (0, 0, -3), (0, 1, 0), False, _radians(45.0), 400, 400))
    return b.build()


def cube_orthographic_scene() -> Scene:
    b = SceneBuilder()
    mat = _white_diffuse(b)
    b.add_shape_at_position(make_cube(1.0), mat, (0.0, 0.0, -3.0))
    b.add_camera(Camera.lookat_camera_orthographic((1.0, 0.75, -1.0), (0, 0, -3), (0, 1, 0), False, 400, 400,
                                                   float(f32(2.5) / f32(400.0))))
    return b.build()


def checkered_plane_scene() -> Scene:
    b = SceneBuilder()
    plane = make_plane((-100, -100, 0.1), (100, -100, 0.1), (100, 100, 0.1), (-100, 100, 0.1), (0, 0, 1))
    plane.uvs = np.array([(-500, -500), (500, -500), (500, 500), (-500, 500)], dtype=f32)
    tex = b.add_texture(Texture(_ffi.TEXTURE_CHECKER, value=(0, 0, 0, 1), value2=(1, 1, 1, 1)))
    mat = b.add_material(Material(_ffi.MATERIAL_DIFFUSE, albedo=tex))
    b.add_shape_at_position(plane, mat, (0, 0, 0))
    b.add_light(Light(_ffi.LIGHT_DIRECTION, a=(0.0, 0.0, -1.0), b=(1000.0, 1000.0, 1000.0)))
    y_angle = f32(_radians(10.0))
    b.add_camera(Camera.lookat_camera_perspective(
        (0.0, 0.0, 0.22),
        (0.0, float(f32(math.cos(y_angle)) * f32(1.0)), float(f32(0.22) - f32(f32(math.sin(y_angle)) * f32(1.0)))),
        (0, 0, 1), False, _radians(40.0), 480, 270))
    return b.build()


def cornell_box() -> SceneBuilder:
    """test_scenes/mod.rs:277-377"""
    b = SceneBuilder()
    w, h, d = 2.0, 1.5, 2.0
    left, right, bottom, top, back, front = w / 2, -w / 2, 0.0, h, -d / 2, d / 2
    up, down, leftn, rightn, backn = (0, 0, 1), (0, 0, -1), (-1, 0, 0), (1, 0, 0), (0, 1, 0)
    floor = make_plane((right, front, bottom), (right, back, bottom), (left, back, bottom), (left, front, bottom), up)
    ceiling = make_plane((left, front, top), (left, back, top), (right, back, top), (right, front, top), down)
    left_wall = make_plane((left, front, bottom), (left, back, bottom), (left, back, top), (left, front, top), leftn)
    right_wall = make_plane((right, front, top), (right, back, top), (right, back, bottom), (right, front, bottom), rightn)
    back_wall = make_plane((right, back, top), (left, back, top), (left, back, bottom), (right, back, bottom), backn)
    white = b.add_constant_texture((0.6, 0.6, 0.6, 1.0))
    red = b.add_constant_texture((0.6, 0.2, 0.2, 1.0))
    blue = b.add_constant_texture((0.2, 0.2, 0.6, 1.0))
    wd = b.add_material(Material(_ffi.MATERIAL_DIFFUSE, albedo=white))
    rd = b.add_material(Material(_ffi.MATERIAL_DIFFUSE, albedo=red))
    bd = b.add_material(Material(_ffi.MATERIAL_DIFFUSE, albedo=blue))
    for mesh, mat in ((floor, wd), (ceiling, wd), (left_wall, rd), (right_wall, bd), (back_wall, wd)):
        b.add_shape_at_position(mesh, mat, (0, 0, 0))
    b.add_camera(Camera.lookat_camera_perspective((0.0, front + 3.4, 0.4), (0, 0, h / 2), (0, 0, 1), False,
                                                  _radians(37.8), 500, 500))
    b.add_point_light((0.0, 0.0, top - 0.1), (1000.0, 1000.0, 1000.0))
    return b


def _cornell_with_sphere(material: Material, extra_textures) -> Scene:
    b = cornell_box()
    ids = [b.add_constant_texture(v) for v in extra_textures]
    mat = b.add_material(material(ids))
    b.add_shape_at_position(Sphere((0, 0, 0), 0.5), mat, (0.0, 0.0, 0.75))
    return b.build()


def dielectric_scene() -> Scene:
    return _cornell_with_sphere(lambda t: Material(_ffi.MATERIAL_SMOOTH_DIELECTRIC, eta=t[0]), [(1.5, 0, 0, 0)])


def metal_scene() -> Scene:
    return _cornell_with_sphere(lambda t: Material(_ffi.MATERIAL_SMOOTH_CONDUCTOR, eta=t[0], kappa=t[1]),
                                [(0.13, 0.43, 1.38, 0.0), (4.10, 2.46, 1.91, 0.0)])


def rough_metal_scene() -> Scene:
    return _cornell_with_sphere(
        lambda t: Material(_ffi.MATERIAL_ROUGH_CONDUCTOR, eta=t[0], kappa=t[1], roughness=t[2], remap_roughness=True),
        [(0.13, 0.43, 1.38, 0.0), (4.10, 2.46, 1.91, 0.0), (0.5, 0.5, 0.0, 0.0)])


def rough_dielectric_scene() -> Scene:
    return _cornell_with_sphere(
        lambda t: Material(_ffi.MATERIAL_ROUGH_DIELECTRIC, eta=t[0], roughness=t[1], remap_roughness=True),
        [(1.5, 0, 0, 0), (0.5, 0.5, 0.0, 0.0)])


def out_of_focus_sphere_scene() -> Scene:
    b = SceneBuilder()
    mat = _white_diffuse(b)
    b.add_shape_at_position(Sphere((0, 0, 0), 1.0), mat, (0.0, 0.0, -5.0))
    b.add_light(Light(_ffi.LIGHT_DIRECTION, a=(0.0, 0.0, -1.0), b=(1.0, 1.0, 1.0)))
    b.add_camera(Camera.lookat_camera_thin_lens_perspective((0, 0, 0), (0, 0, -5), (0, 1, 0), False, _radians(45.0),
                                                            400, 400, 0.1, 3.0))
    return b.build()


def coated_diffuse_bunny_scene(bunny: Optional[Mesh] = None, bunny_ply: Optional[bytes] = None) -> Scene:
    """test_scenes/mod.rs:520-553. The bunny mesh is an asset (assets/bunny.ply): pass it in."""
    if bunny is None:
        if bunny_ply is None:
            raise ValueError("coated_diffuse_bunny needs the bunny mesh (fixture tests/golden/scenes/bunny_mesh.npz)")
        bunny = mesh_from_ply_bytes(bunny_ply, False)
    b = cornell_box()
    diffuse_albedo = b.add_constant_texture((0.8, 0.2, 0.2, 1.0))
    eta = b.add_constant_texture((1.5, 0, 0, 0))
    rough = b.add_constant_texture((0.1, 0.1, 0, 0))
    thick = b.add_constant_texture((0.5, 0, 0, 0))
    coat = b.add_constant_texture((1, 1, 1, 1))
    mat = b.add_material(Material(_ffi.MATERIAL_COATED_DIFFUSE, albedo=diffuse_albedo, eta=eta, roughness=rough,
                                  thickness=thick, coat_albedo=coat, remap_roughness=True))
    b.add_shape_at_position(bunny, mat, (0.0, 0.0, 0.25))
    return b.build()


def environment_lighting_scene(environment_map: np.ndarray) -> Scene:
    """test_scenes/mod.rs:555-603 with a caller-supplied lat-long image (the reference's EXR is a
    missing large blob)."""
    b = SceneBuilder()
    img = b.add_image(environment_map)
    tex = b.add_texture(Texture(_ffi.TEXTURE_IMAGE, image=img, filter=_ffi.FILTER_NEAREST, wrap=_ffi.WRAP_REPEAT))
    b.add_environment_light(tex)
    mat = _white_diffuse(b)
    b.add_shape_at_position(make_cube(1.0), mat, (0.0, 15.0, 0.0))
    b.add_camera(Camera.lookat_camera_perspective((0, 0, 0), (0, 1, 0), (0, 0, 1), False, _radians(37.8), 500, 500))
    return b.build()


def synthetic_environment_map(w: int = 64, h: int = 32) -> np.ndarray:
    """A small procedural f32 RGB sky used where the reference would load lake_pier_1k.exr."""
    y = np.linspace(0.0, 1.0, h, dtype=f32)[:, None]
    x = np.linspace(0.0, 1.0, w, dtype=f32)[None, :]
    r = (0.2 + 0.8 * (1.0 - y)) * np.ones_like(x)
    g = (0.3 + 0.5 * (1.0 - y)) * (0.75 + 0.25 * np.cos(2 * np.pi * x))
    b_ = 0.4 + 0.6 * (1.0 - y) * np.ones_like(x)
    return np.stack([r, g, b_], axis=2).astype(f32)


@dataclass
class TestScene:
    name: str
    scene_func: Callable[[], Scene]
    settings_func: Callable[[], RaytracerSettings]
    __test__ = False


def _normals_settings() -> RaytracerSettings:
    return RaytracerSettings(outputs=AovFlags.NORMALS)


def all_test_scenes():
    """test_scenes/mod.rs:618-692 (environment_light / coated_diffuse_bunny need their assets)."""
    return [
        TestScene("sphere", sphere_scene, _normals_settings),
        TestScene("cube", cube_scene, _normals_settings),
        TestScene("cube_orthographic", cube_orthographic_scene, _normals_settings),
        TestScene("checkered_plane", checkered_plane_scene, lambda: RaytracerSettings(samples_per_pixel=1)),
        TestScene("dielectric", dielectric_scene, RaytracerSettings),
        TestScene("metal", metal_scene, RaytracerSettings),
        TestScene("rough_metal", rough_metal_scene, RaytracerSettings),
        TestScene("rough_dielectric", rough_dielectric_scene, RaytracerSettings),
        TestScene("out_of_focus_sphere", out_of_focus_sphere_scene,
                  lambda: RaytracerSettings(sampler=Sampler.stratified(True, 6, 6), samples_per_pixel=36)),
    ]


# --------------------------------------------------------------------------------------------
# BASELINE config C5: synthetic procedural mesh. Deterministic, seed 42.
# --------------------------------------------------------------------------------------------
def _hash_lattice(ix, iy, iz, seed: int) -> np.ndarray:
    """PCG-style integer hash of a lattice point -> [0,1) f32."""
    h = (ix.astype(np.uint64) * np.uint64(0x9E3779B1) ^ iy.astype(np.uint64) * np.uint64(0x85EBCA77)
         ^ iz.astype(np.uint64) * np.uint64(0xC2B2AE3D) ^ np.uint64(seed)) & np.uint64(0xFFFFFFFF)
    h = (h * np.uint64(747796405) + np.uint64(2891336453)) & np.uint64(0xFFFFFFFF)
    h = (((h >> ((h >> np.uint64(28)) + np.uint64(4))) ^ h) * np.uint64(277803737)) & np.uint64(0xFFFFFFFF)
    h = ((h >> np.uint64(22)) ^ h) & np.uint64(0xFFFFFFFF)
    return (h >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def value_noise3(p: np.ndarray, seed: int = 42) -> np.ndarray:
    """Trilinear value noise over the integer lattice, smoothstep-faded. p: [...,3] f32."""
    p = p.astype(np.float64)
    i = np.floor(p).astype(np.int64)
    f = p - i
    f = f * f * (3.0 - 2.0 * f)
    out = np.zeros(p.shape[:-1], dtype=np.float64)
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                wgt = ((f[..., 0] if dx else 1 - f[..., 0]) * (f[..., 1] if dy else 1 - f[..., 1])
                       * (f[..., 2] if dz else 1 - f[..., 2]))
                out += wgt * _hash_lattice((i[..., 0] + dx) & 0xFFFF, (i[..., 1] + dy) & 0xFFFF, (i[..., 2] + dz) & 0xFFFF, seed)
    return out.astype(np.float32)


def procedural_sphere_mesh(n_lon: int, n_lat: int, radius: float = 0.5, displacement: float = 0.1,
                           frequency: float = 8.0, seed: int = 42) -> Mesh:
    """UV-sphere of n_lon x n_lat quads (2*n_lon*n_lat triangles, (n_lon+1)*(n_lat+1) vertices) with
    radial value-noise displacement and smooth normals from the displaced surface; no UVs."""
    lon = np.linspace(0.0, 2.0 * np.pi, n_lon + 1, dtype=np.float64)
    lat = np.linspace(0.0, np.pi, n_lat + 1, dtype=np.float64)
    st, ct = np.sin(lat)[:, None], np.cos(lat)[:, None]
    d = np.stack([st * np.cos(lon)[None, :], st * np.sin(lon)[None, :], ct * np.ones_like(lon)[None, :]], axis=2)
    r = radius + displacement * (value_noise3((d * frequency).astype(np.float32), seed).astype(np.float64) - 0.5)
    pos = d * r[..., None]
    # smooth normals by central differences of the displaced grid (wrap in longitude, clamp in latitude)
    dlon = np.roll(pos, -1, axis=1) - np.roll(pos, 1, axis=1)
    dlon[:, 0] = pos[:, 1] - pos[:, -2]
    dlon[:, -1] = dlon[:, 0]
    dlat = np.empty_like(pos)
    dlat[1:-1] = pos[2:] - pos[:-2]
    dlat[0] = pos[1] - pos[0]
    dlat[-1] = pos[-1] - pos[-2]
    n = np.cross(dlat, dlon)
    ln = np.linalg.norm(n, axis=2, keepdims=True)
    n = np.where(ln > 1e-20, n / np.maximum(ln, 1e-20), d)
    n = np.where((np.sum(n * d, axis=2, keepdims=True) < 0), -n, n)
    W = n_lon + 1
    j, i = np.meshgrid(np.arange(n_lat, dtype=np.uint32), np.arange(n_lon, dtype=np.uint32), indexing="ij")
    v00 = j * W + i
    v01 = v00 + 1
    v10 = v00 + W
    v11 = v10 + 1
    tris = np.stack([np.stack([v00, v10, v11], axis=-1), np.stack([v00, v11, v01], axis=-1)], axis=2).reshape(-1, 3)
    # drop the degenerate triangles at the two poles (zero area: two coincident vertices)
    keep = np.ones(len(tris), dtype=bool)
    tri_grid = keep.reshape(n_lat, n_lon, 2)
    tri_grid[0, :, 1] = False      # (v00, v11, v01) collapses at the north pole row
    tri_grid[-1, :, 0] = False     # (v00, v10, v11) collapses at the south pole row
    tris = tris[tri_grid.reshape(-1)]
    return Mesh(pos.reshape(-1, 3).astype(f32), tris.astype(np.uint32), n.reshape(-1, 3).astype(f32), None)


def synthetic_mesh_scene(base: Scene, n_lon: int = 4096, n_lat: int = 2048, center=(0.0, 0.3255, 0.0)) -> Scene:
    """Config C5: the procedural mesh (grey Diffuse 0.5) placed inside the Cornell box `base`
    (cb.glb with its area light and camera)."""
    mesh = procedural_sphere_mesh(n_lon, n_lat)
    b = SceneBuilder()
    b.scene = base
    grey = b.add_constant_texture((0.5, 0.5, 0.5, 1.0))
    mat = b.add_material(Material(_ffi.MATERIAL_DIFFUSE, albedo=grey))
    b.add_shape_with_transform(mesh, mat, Transform.translate(center), None)
    return b.build()
