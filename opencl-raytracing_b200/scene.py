"""Scene description — host-side mirror of crates/raytracing/src/scene/{scene,primitive,camera}.rs,
lights/light.rs, materials/{mod,texture,image}.rs, flattened the way `rtcuda_scene_desc`
(include/rtcuda.h) wants it.

The reference keeps a primitive graph (Basic / Transform / Aggregate, scene/primitive.rs:119-145) whose
only shape in practice is "one root Aggregate of Transform -> Basic" (scene.rs:483-509, 594-602,
634-662). This mirror stores exactly that flattened form:

  shapes[]     BasicPrimitive in creation order (index = position among Basic primitives)
  instances[]  the root Aggregate's children in order; index = the `geom_id` of PrimPtr (bvh2.rs:278-283)

Everything here is host-side *description* (no per-ray work): the data-parallel path lives in
libraytracing_cuda.so.
"""
from __future__ import annotations

import ctypes as C
import io
import json
import math
import struct
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _ffi
from .geometry import Transform, f32, mat_identity, mat_invert, matmul, unit, cross, vec3

NONE = _ffi.NONE


# --------------------------------------------------------------------------------------------
# Camera (scene/camera.rs)
# --------------------------------------------------------------------------------------------
def _screen_to_raster(width: int, height: int, tl, br) -> Transform:
    """camera.rs:41-59"""
    tl = np.asarray(tl, dtype=f32)
    br = np.asarray(br, dtype=f32)
    screen_to_zero = Transform.translate(-tl)
    sx = f32(br[0] - tl[0])
    sy = f32(br[1] - tl[1])
    screen_to_ndc = screen_to_zero.compose(Transform.scale([f32(1.0) / sx, f32(1.0) / sy, f32(1.0)]))
    return screen_to_ndc.compose(Transform.scale([f32(width), f32(height), f32(1.0)]))


def _perspective_transform(far_clip, near_clip, yfov, width: int, height: int) -> Transform:
    """camera.rs:63-107 (note the `yfov * (w/h)` horizontal FOV quirk for wide rasters)."""
    far_clip, near_clip, yfov = f32(far_clip), f32(near_clip), f32(yfov)
    persp = mat_identity()
    persp[2, 2] = f32(far_clip / f32(far_clip - near_clip))
    persp[2, 3] = f32(-f32(f32(far_clip * near_clip) / f32(far_clip - near_clip)))
    persp[3, 2] = f32(1.0)
    persp[3, 3] = f32(0.0)
    persp_t = Transform(persp)
    wide = width >= height
    fov = f32(yfov * f32(f32(width) / f32(height))) if wide else yfov
    invt = f32(f32(1.0) / f32(math.tan(f32(fov / f32(2.0)))))
    fov_scale = Transform.scale([-invt, -invt, f32(1.0)])
    if wide:
        k = f32(f32(height) / f32(width))
        tl, br = [-1.0, -k, 0.0], [1.0, k, 0.0]
    else:
        k = f32(f32(width) / f32(height))
        tl, br = [-k, -1.0, 0.0], [k, 1.0, 0.0]
    return persp_t.compose(fov_scale).compose(_screen_to_raster(width, height, tl, br))


def _orthographic_transform(far_clip, near_clip, width, height, ssw, ssh) -> Transform:
    """camera.rs:109-131"""
    far_clip, near_clip, ssw, ssh = f32(far_clip), f32(near_clip), f32(ssw), f32(ssh)
    translate = Transform.translate([0.0, 0.0, -near_clip])
    scale = Transform.scale([1.0, 1.0, f32(f32(1.0) / f32(far_clip - near_clip))])
    tl = [f32(-ssw / f32(2.0)), f32(-ssh / f32(2.0)), 0.0]
    br = [f32(ssw / f32(2.0)), f32(ssh / f32(2.0)), 0.0]
    return translate.compose(scale).compose(_screen_to_raster(width, height, tl, br))


@dataclass
class Camera:
    """scene/camera.rs:5-36. The three transforms are kept verbatim and uploaded as-is."""
    kind: int
    raster_width: int
    raster_height: int
    near_clip: float
    far_clip: float
    world_to_raster: Transform
    camera_to_world: Transform
    raster_to_camera: Transform
    yfov: float = 0.0
    aperture_radius: float = 0.0
    focal_distance: float = 0.0
    screen_space_width: float = 0.0
    screen_space_height: float = 0.0
    # glTF provenance, kept so the raster size can be re-derived (512x512 / 1080p / 4K overrides)
    _gltf: Optional[dict] = None

    @staticmethod
    def lookat_camera_perspective(pos, target, up, swap_handedness, yfov, w, h) -> "Camera":
        """camera.rs:206-243"""
        c2r = _perspective_transform(1000.0, 0.01, yfov, w, h)
        c2w = Transform.look_at(pos, target, up, swap_handedness)
        return Camera(_ffi.CAMERA_PINHOLE, w, h, 0.01, 1000.0, c2w.invert().compose(c2r), c2w, c2r.invert(),
                      yfov=float(f32(yfov)))

    @staticmethod
    def lookat_camera_orthographic(pos, target, up, swap_handedness, w, h, raster_to_screen_ratio) -> "Camera":
        """camera.rs:245-289"""
        ssw = f32(f32(w) * f32(raster_to_screen_ratio))
        ssh = f32(f32(h) * f32(raster_to_screen_ratio))
        c2r = _orthographic_transform(1000.0, 0.01, w, h, ssw, ssh)
        c2w = Transform.look_at(pos, target, up, swap_handedness)
        return Camera(_ffi.CAMERA_ORTHOGRAPHIC, w, h, 0.01, 1000.0, c2w.invert().compose(c2r), c2w, c2r.invert(),
                      screen_space_width=float(ssw), screen_space_height=float(ssh))

    @staticmethod
    def lookat_camera_thin_lens_perspective(pos, target, up, swap_handedness, yfov, w, h, aperture_radius,
                                            focal_distance) -> "Camera":
        """camera.rs:292-335"""
        c2r = _perspective_transform(1000.0, 0.01, yfov, w, h)
        c2w = Transform.look_at(pos, target, up, swap_handedness)
        return Camera(_ffi.CAMERA_THIN_LENS, w, h, 0.01, 1000.0, c2w.invert().compose(c2r), c2w, c2r.invert(),
                      yfov=float(f32(yfov)), aperture_radius=float(f32(aperture_radius)),
                      focal_distance=float(f32(focal_distance)))

    @staticmethod
    def from_gltf_camera_node(node_matrix: np.ndarray, cam: dict, raster_height: int,
                              raster_width: Optional[int] = None) -> "Camera":
        """camera.rs:133-203. `raster_width` overrides `(height * aspect) as usize` (BASELINE config
        C2 asks for 512x512, which the reference CLI cannot express: SURVEY §8a)."""
        m = np.asarray(node_matrix, dtype=f32).reshape(4, 4)
        flip_y = Transform.scale([1.0, -1.0, 1.0])
        camera_to_world = flip_y.compose(Transform(m))
        world_to_camera = Transform(mat_invert(m))
        if cam["type"] == "perspective":
            p = cam["perspective"]
            width = raster_width if raster_width is not None else int(f32(f32(raster_height) * f32(p["aspectRatio"])))
            zfar = p.get("zfar", 1000.0)
            c2r = _perspective_transform(-f32(zfar), -f32(p["znear"]), p["yfov"], width, raster_height)
            kind, extra = _ffi.CAMERA_PINHOLE, dict(yfov=float(f32(p["yfov"])))
        else:
            o = cam["orthographic"]
            ssw, ssh = f32(o["xmag"]), f32(o["ymag"])
            width = raster_width if raster_width is not None else int(f32(f32(f32(raster_height) * ssw) / ssh))
            c2r = _orthographic_transform(-f32(o["zfar"]), -f32(o["znear"]), width, raster_height, ssw, -ssh)
            kind, extra = _ffi.CAMERA_ORTHOGRAPHIC, dict(screen_space_width=float(ssw), screen_space_height=float(ssh))
        return Camera(kind, width, raster_height, 0.01, 1000.0, world_to_camera.compose(c2r), camera_to_world,
                      c2r.invert(), _gltf={"matrix": m.reshape(-1).tolist(), "camera": cam}, **extra)

    def with_raster_size(self, width: int, height: int) -> "Camera":
        """Re-derive the camera for another raster size (glTF cameras only)."""
        if self._gltf is None:
            raise ValueError("raster override is only defined for glTF cameras")
        return Camera.from_gltf_camera_node(np.array(self._gltf["matrix"], dtype=f32), self._gltf["camera"], height, width)

    def to_c(self) -> _ffi.Camera:
        c = _ffi.Camera()
        c.kind = self.kind
        c.raster_width, c.raster_height = self.raster_width, self.raster_height
        c.near_clip, c.far_clip = self.near_clip, self.far_clip
        c.yfov, c.aperture_radius, c.focal_distance = self.yfov, self.aperture_radius, self.focal_distance
        c.screen_space_width, c.screen_space_height = self.screen_space_width, self.screen_space_height
        for name in ("world_to_raster", "camera_to_world", "raster_to_camera"):
            _fill_transform(getattr(c, name), getattr(self, name))
        return c


def _fill_mat(dst: _ffi.Mat4, m: np.ndarray) -> None:
    flat = np.asarray(m, dtype=f32).reshape(16)
    for i in range(16):
        dst.m[i] = float(flat[i])


def _fill_transform(dst: _ffi.Transform, t: Transform) -> None:
    _fill_mat(dst.forward, t.forward)
    _fill_mat(dst.inverse, t.inverse)


# --------------------------------------------------------------------------------------------
# Shapes, lights, materials, textures, images
# --------------------------------------------------------------------------------------------
@dataclass
class Mesh:
    """geometry/shapes/mesh.rs:71-76"""
    vertices: np.ndarray                      # [V,3] f32
    tris: np.ndarray                          # [T,3] u32
    normals: Optional[np.ndarray] = None      # [V,3] f32 or None (empty Vec)
    uvs: Optional[np.ndarray] = None          # [V,2] f32 or None


@dataclass
class Sphere:
    center: tuple
    radius: float


@dataclass
class BasicPrimitive:
    """scene/primitive.rs:127-132"""
    shape: object            # Mesh | Sphere
    material: int
    area_light: Optional[int] = None


@dataclass
class Light:
    """lights/light.rs:7-28"""
    kind: int
    a: tuple = (0.0, 0.0, 0.0)          # position (point) / direction (direction)
    b: tuple = (0.0, 0.0, 0.0)          # intensity / radiance
    shape: int = NONE                   # DiffuseAreaLight::prim_id
    light_to_world: Optional[np.ndarray] = None


@dataclass
class Material:
    """materials/mod.rs:2-56; unused texture slots are NONE."""
    kind: int
    albedo: int = NONE
    eta: int = NONE
    kappa: int = NONE
    roughness: int = NONE
    thickness: int = NONE
    coat_albedo: int = NONE
    remap_roughness: bool = False


@dataclass
class Texture:
    """materials/texture.rs:9-112"""
    kind: int
    image: int = 0
    filter: int = 0
    wrap: int = 0
    a: int = NONE
    b: int = NONE
    c: int = NONE
    value: tuple = (0.0, 0.0, 0.0, 0.0)
    value2: tuple = (0.0, 0.0, 0.0, 0.0)


_NP_FORMAT = {np.dtype(np.uint8): _ffi.IMAGE_U8, np.dtype(np.uint16): _ffi.IMAGE_U16, np.dtype(np.float32): _ffi.IMAGE_F32}


@dataclass
class Scene:
    """scene/scene.rs:13-27 in flattened form."""
    camera: Camera
    shapes: List[BasicPrimitive] = field(default_factory=list)
    instances: List[tuple] = field(default_factory=list)     # (shape index, Transform)
    lights: List[Light] = field(default_factory=list)
    materials: List[Material] = field(default_factory=list)
    textures: List[Texture] = field(default_factory=list)
    images: List[np.ndarray] = field(default_factory=list)   # [H,W,C] u8 / u16 / f32
    environment_light: Optional[int] = None                   # texture id (EnvironmentLight::radiance)

    # -- flattening into the C ABI -----------------------------------------------------------
    def to_desc(self, own_arrays: bool = False) -> "SceneDescHolder":
        """Flatten into an rtcuda_scene_desc. own_arrays=True points every mesh at its own numpy arrays (rtcuda_shape.vertices /
        tris / normals / uvs) instead of concatenating them into scene-wide arrays — what the CUDA backend uploads from; the
        oracle and the CPU harness read the concatenated form."""
        return SceneDescHolder(self, own_arrays)

    def triangle_count(self) -> int:
        return sum(len(self.shapes[s].shape.tris) if isinstance(self.shapes[s].shape, Mesh) else 0 for s, _ in self.instances)

    # -- npz round trip (test fixtures) ------------------------------------------------------
    def save_npz(self, path: str) -> None:
        arrays = {}
        meta = {"camera": _camera_to_json(self.camera), "shapes": [], "instances": [], "lights": [], "materials": [],
                "textures": [], "n_images": len(self.images), "environment_light": self.environment_light}
        for i, bp in enumerate(self.shapes):
            s = bp.shape
            if isinstance(s, Mesh):
                arrays[f"s{i}_v"] = s.vertices
                arrays[f"s{i}_t"] = s.tris
                if s.normals is not None:
                    arrays[f"s{i}_n"] = s.normals
                if s.uvs is not None:
                    arrays[f"s{i}_uv"] = s.uvs
                meta["shapes"].append({"kind": "mesh", "material": bp.material, "area_light": bp.area_light})
            else:
                meta["shapes"].append({"kind": "sphere", "center": [float(c) for c in s.center], "radius": float(s.radius),
                                       "material": bp.material, "area_light": bp.area_light})
        for j, (si, tf) in enumerate(self.instances):
            meta["instances"].append(si)
            arrays[f"i{j}_f"] = tf.forward
            arrays[f"i{j}_i"] = tf.inverse
        for k, l in enumerate(self.lights):
            meta["lights"].append({"kind": l.kind, "a": [float(x) for x in l.a], "b": [float(x) for x in l.b], "shape": l.shape})
            if l.light_to_world is not None:
                arrays[f"l{k}_m"] = l.light_to_world
        for m in self.materials:
            meta["materials"].append(m.__dict__)
        for t in self.textures:
            d = dict(t.__dict__)
            d["value"] = [float(x) for x in t.value]
            d["value2"] = [float(x) for x in t.value2]
            meta["textures"].append(d)
        for k, im in enumerate(self.images):
            arrays[f"img{k}"] = im
        arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        np.savez_compressed(path, **arrays)

    @staticmethod
    def load_npz(path: str) -> "Scene":
        z = np.load(path)
        meta = json.loads(bytes(z["meta"]).decode())
        sc = Scene(camera=_camera_from_json(meta["camera"]))
        for i, s in enumerate(meta["shapes"]):
            if s["kind"] == "mesh":
                shape = Mesh(z[f"s{i}_v"], z[f"s{i}_t"], z[f"s{i}_n"] if f"s{i}_n" in z else None,
                             z[f"s{i}_uv"] if f"s{i}_uv" in z else None)
            else:
                shape = Sphere(tuple(s["center"]), s["radius"])
            sc.shapes.append(BasicPrimitive(shape, s["material"], s["area_light"]))
        for j, si in enumerate(meta["instances"]):
            sc.instances.append((si, Transform(z[f"i{j}_f"], z[f"i{j}_i"])))
        for k, l in enumerate(meta["lights"]):
            sc.lights.append(Light(l["kind"], tuple(l["a"]), tuple(l["b"]), l["shape"], z[f"l{k}_m"] if f"l{k}_m" in z else None))
        for m in meta["materials"]:
            sc.materials.append(Material(**m))
        for t in meta["textures"]:
            t["value"], t["value2"] = tuple(t["value"]), tuple(t["value2"])
            sc.textures.append(Texture(**t))
        for k in range(meta["n_images"]):
            sc.images.append(z[f"img{k}"])
        sc.environment_light = meta["environment_light"]
        return sc


def _camera_to_json(c: Camera) -> dict:
    d = {k: getattr(c, k) for k in ("kind", "raster_width", "raster_height", "near_clip", "far_clip", "yfov",
                                     "aperture_radius", "focal_distance", "screen_space_width", "screen_space_height")}
    for name in ("world_to_raster", "camera_to_world", "raster_to_camera"):
        t = getattr(c, name)
        # f32 -> python float -> f32 is exact
        d[name] = [t.forward.reshape(-1).astype(float).tolist(), t.inverse.reshape(-1).astype(float).tolist()]
    d["_gltf"] = c._gltf
    return d


def _camera_from_json(d: dict) -> Camera:
    tfs = {name: Transform(np.array(d[name][0], dtype=f32), np.array(d[name][1], dtype=f32))
           for name in ("world_to_raster", "camera_to_world", "raster_to_camera")}
    return Camera(d["kind"], d["raster_width"], d["raster_height"], d["near_clip"], d["far_clip"], tfs["world_to_raster"],
                  tfs["camera_to_world"], tfs["raster_to_camera"], d["yfov"], d["aperture_radius"], d["focal_distance"],
                  d["screen_space_width"], d["screen_space_height"], d.get("_gltf"))


class SceneDescHolder:
    """Owns the host arrays an `rtcuda_scene_desc` points into (keep alive until upload returns)."""

    def __init__(self, scene: Scene, own_arrays: bool = False):
        self.scene = scene
        shapes = (_ffi.Shape * max(1, len(scene.shapes)))()
        verts, tris, normals, uvs = [], [], [], []
        keep_arrays = []
        nv = nt = nn = nuv = 0
        for i, bp in enumerate(scene.shapes):
            s = shapes[i]
            s.material = bp.material
            s.area_light = NONE if bp.area_light is None else bp.area_light
            s.normal_offset = s.uv_offset = NONE
            if isinstance(bp.shape, Mesh):
                m = bp.shape
                s.kind = _ffi.SHAPE_TRIANGLE_MESH
                s.vertex_offset, s.vertex_count = nv, len(m.vertices)
                s.tri_offset, s.tri_count = nt, len(m.tris)
                mv = np.ascontiguousarray(m.vertices, dtype=f32).reshape(-1, 3)     # no copy when already f32 / u32 and contiguous
                mt = np.ascontiguousarray(m.tris, dtype=np.uint32).reshape(-1, 3)
                mn = np.ascontiguousarray(m.normals, dtype=f32).reshape(-1, 3) if m.normals is not None and len(m.normals) else None
                mu = np.ascontiguousarray(m.uvs, dtype=f32).reshape(-1, 2) if m.uvs is not None and len(m.uvs) else None
                if (mn is not None and len(mn) != len(mv)) or (mu is not None and len(mu) != len(mv)):   # the library copies vertex_count entries of each
                    raise ValueError(f"shape {i}: normals / uvs must have one entry per vertex ({len(mv)})")
                nv += len(mv)
                nt += len(mt)
                if mn is not None:
                    s.normal_offset = nn
                    nn += len(mn)
                if mu is not None:
                    s.uv_offset = nuv
                    nuv += len(mu)
                if own_arrays:
                    keep_arrays += [mv, mt, mn, mu]
                    # an empty mesh still needs a non-NULL vertices pointer to say "own arrays"
                    if not len(mv):
                        mv = np.zeros((1, 3), dtype=f32)
                        keep_arrays.append(mv)
                    s.vertices = mv.ctypes.data_as(C.POINTER(C.c_float))
                    s.tris = mt.ctypes.data_as(C.POINTER(C.c_uint32)) if len(mt) else None
                    s.normals = mn.ctypes.data_as(C.POINTER(C.c_float)) if mn is not None else None
                    s.uvs = mu.ctypes.data_as(C.POINTER(C.c_float)) if mu is not None else None
                else:
                    verts.append(mv)
                    tris.append(mt)
                    if mn is not None:
                        normals.append(mn)
                    if mu is not None:
                        uvs.append(mu)
            else:
                s.kind = _ffi.SHAPE_SPHERE
                s.center[0], s.center[1], s.center[2] = [float(f32(c)) for c in bp.shape.center]
                s.radius = float(f32(bp.shape.radius))
        cat = lambda lst, w, dt: np.ascontiguousarray(np.concatenate(lst, axis=0) if lst else np.zeros((0, w), dtype=dt))
        self.vertices, self.tris = cat(verts, 3, f32), cat(tris, 3, np.uint32)
        self.normals, self.uvs = cat(normals, 3, f32), cat(uvs, 2, f32)

        instances = (_ffi.Instance * max(1, len(scene.instances)))()
        for j, (si, tf) in enumerate(scene.instances):
            instances[j].shape = si
            _fill_transform(instances[j].object_to_world, tf)

        lights = (_ffi.Light * max(1, len(scene.lights)))()
        for k, l in enumerate(scene.lights):
            lights[k].kind = l.kind
            lights[k].shape = l.shape
            for q in range(3):
                lights[k].position_or_direction[q] = float(f32(l.a[q]))
                lights[k].intensity_or_radiance[q] = float(f32(l.b[q]))
            _fill_mat(lights[k].light_to_world, l.light_to_world if l.light_to_world is not None else mat_identity())

        materials = (_ffi.Material * max(1, len(scene.materials)))()
        for k, m in enumerate(scene.materials):
            mm = materials[k]
            mm.kind, mm.remap_roughness = m.kind, int(m.remap_roughness)
            mm.albedo, mm.eta, mm.kappa, mm.roughness = m.albedo, m.eta, m.kappa, m.roughness
            mm.thickness, mm.coat_albedo = m.thickness, m.coat_albedo

        textures = (_ffi.Texture * max(1, len(scene.textures)))()
        for k, t in enumerate(scene.textures):
            tt = textures[k]
            tt.kind, tt.image, tt.filter, tt.wrap, tt.a, tt.b, tt.c = t.kind, t.image, t.filter, t.wrap, t.a, t.b, t.c
            for q in range(4):
                tt.value[q] = float(f32(t.value[q]))
                tt.value2[q] = float(f32(t.value2[q]))

        images = (_ffi.Image * max(1, len(scene.images)))()
        blobs, off = [], 0
        for k, im in enumerate(scene.images):
            im = np.ascontiguousarray(im)
            if im.ndim == 2:
                im = im[:, :, None]
            images[k].height, images[k].width, images[k].channels = im.shape
            images[k].format = _NP_FORMAT[im.dtype]
            images[k].byte_offset = off
            b = im.tobytes()
            pad = (-len(b)) % 16
            blobs.append(b + b"\0" * pad)
            off += len(b) + pad
        self.image_bytes = np.frombuffer(b"".join(blobs) if blobs else b"\0" * 16, dtype=np.uint8).copy()

        self._keep = (shapes, instances, lights, materials, textures, images, keep_arrays)
        self.geometry_bytes = sum(a.nbytes for a in keep_arrays if isinstance(a, np.ndarray)) if own_arrays else \
            int(self.vertices.nbytes + self.tris.nbytes + self.normals.nbytes + self.uvs.nbytes)
        d = _ffi.SceneDesc()
        d.abi_version = _ffi.ABI_VERSION
        d.camera = scene.camera.to_c()
        d.shapes, d.shape_count = shapes, len(scene.shapes)
        d.instances, d.instance_count = instances, len(scene.instances)
        d.lights, d.light_count = lights, len(scene.lights)
        d.materials, d.material_count = materials, len(scene.materials)
        d.textures, d.texture_count = textures, len(scene.textures)
        d.images, d.image_count = images, len(scene.images)
        d.environment_light_texture = NONE if scene.environment_light is None else scene.environment_light
        d.vertices = self.vertices.ctypes.data_as(C.POINTER(C.c_float))
        d.vertex_count = len(self.vertices)
        d.tris = self.tris.ctypes.data_as(C.POINTER(C.c_uint32))
        d.tri_count = len(self.tris)
        d.normals = self.normals.ctypes.data_as(C.POINTER(C.c_float))
        d.normal_count = len(self.normals)
        d.uvs = self.uvs.ctypes.data_as(C.POINTER(C.c_float))
        d.uv_count = len(self.uvs)
        d.image_bytes = self.image_bytes.ctypes.data_as(C.POINTER(C.c_uint8))
        d.image_byte_count = len(self.image_bytes) if blobs else 0
        self.desc = d


# --------------------------------------------------------------------------------------------
# SceneBuilder (scene/scene.rs:525-675)
# --------------------------------------------------------------------------------------------
class SceneBuilder:
    def __init__(self):
        self.scene = Scene(camera=None)

    def add_camera(self, camera: Camera):
        self.scene.camera = camera

    def add_environment_light(self, texture_id: int):
        self.scene.environment_light = texture_id

    def add_texture(self, tex: Texture) -> int:
        self.scene.textures.append(tex)
        return len(self.scene.textures) - 1

    def add_constant_texture(self, value) -> int:
        return self.add_texture(Texture(_ffi.TEXTURE_CONSTANT, value=tuple(float(f32(v)) for v in value)))

    def add_material(self, material: Material) -> int:
        self.scene.materials.append(material)
        return len(self.scene.materials) - 1

    def add_image(self, image: np.ndarray) -> int:
        self.scene.images.append(image)
        return len(self.scene.images) - 1

    def add_shape_at_position(self, shape, material_id: int, position):
        self.add_shape_with_transform(shape, material_id, Transform.translate(position), None)

    def add_shape_with_transform(self, shape, material_id: int, transform: Transform, area_light_radiance=None):
        idx = len(self.scene.shapes)
        area = None
        if area_light_radiance is not None:
            area = len(self.scene.lights)
            self.scene.lights.append(Light(_ffi.LIGHT_DIFFUSE_AREA, b=tuple(area_light_radiance), shape=idx,
                                           light_to_world=transform.forward.copy()))
        self.scene.shapes.append(BasicPrimitive(shape, material_id, area))
        self.scene.instances.append((idx, transform))

    def add_instance(self, shape_index: int, transform: Transform):
        """A second Transform primitive over an existing Basic primitive (glTF mesh reuse, scene.rs:430-443)."""
        self.scene.instances.append((shape_index, transform))

    def add_light(self, light: Light):
        self.scene.lights.append(light)

    def add_point_light(self, position, intensity):
        self.add_light(Light(_ffi.LIGHT_POINT, a=tuple(position), b=tuple(intensity)))

    def build(self) -> Scene:
        assert self.scene.camera is not None, "scene description incomplete"
        return self.scene


# --------------------------------------------------------------------------------------------
# glTF / GLB importer (scene/scene.rs:227-523, geometry/shapes/mesh.rs:172-262, lights/light.rs:41-83)
# --------------------------------------------------------------------------------------------
HEIGHT = 600  # scene.rs:247

_COMPONENT = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_NCOMP = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}


def _quat_to_mat(q) -> np.ndarray:
    x, y, z, w = [f32(v) for v in q]
    two = f32(2.0)
    x2, y2, z2 = f32(x + x), f32(y + y), f32(z + z)
    xx2, yy2, zz2 = f32(x2 * x), f32(y2 * y), f32(z2 * z)
    xy2, xz2, yz2 = f32(x2 * y), f32(x2 * z), f32(y2 * z)
    sx2, sy2, sz2 = f32(w * x2), f32(w * y2), f32(w * z2)
    one = f32(1.0)
    m = mat_identity()
    m[0, 0], m[0, 1], m[0, 2] = one - yy2 - zz2, xy2 - sz2, xz2 + sy2
    m[1, 0], m[1, 1], m[1, 2] = xy2 + sz2, one - xx2 - zz2, yz2 - sx2
    m[2, 0], m[2, 1], m[2, 2] = xz2 - sy2, yz2 + sx2, one - xx2 - yy2
    del two
    return m.astype(f32)


def _node_matrix(node: dict) -> np.ndarray:
    """gltf::scene::Transform::matrix(): T * R * S (row-major here; the reference transposes the
    column-major glTF matrix, scene.rs:423-426)."""
    if "matrix" in node:
        return np.array(node["matrix"], dtype=f32).reshape(4, 4).T.copy()
    t = node.get("translation", [0, 0, 0])
    r = node.get("rotation", [0, 0, 0, 1])
    s = node.get("scale", [1, 1, 1])
    m = _quat_to_mat(r)
    for c in range(3):
        m[:3, c] = (m[:3, c] * f32(s[c])).astype(f32)
    m[:3, 3] = np.array(t, dtype=f32)
    return m


def _quat_rotate(q, v) -> np.ndarray:
    """Quaternion(w, xyz).rotate(v) (geometry/quaternion.rs) = q v q*."""
    return (_quat_to_mat(q)[:3, :3] @ np.asarray(v, dtype=f32)).astype(f32)


class _Glb:
    def __init__(self, data: bytes):
        magic, _ver, _ln = struct.unpack("<III", data[:12])
        if magic != 0x46546C67:
            raise ValueError("not a GLB file")
        off = 12
        self.json = None
        self.bin = b""
        while off < len(data):
            clen, ctype = struct.unpack("<II", data[off:off + 8])
            chunk = data[off + 8: off + 8 + clen]
            if ctype == 0x4E4F534A:
                self.json = json.loads(chunk)
            elif ctype == 0x004E4942:
                self.bin = chunk
            off += 8 + clen + ((-clen) % 4)

    def view(self, idx: int) -> bytes:
        bv = self.json["bufferViews"][idx]
        o = bv.get("byteOffset", 0)
        return self.bin[o:o + bv["byteLength"]]

    def accessor(self, idx: int) -> np.ndarray:
        acc = self.json["accessors"][idx]
        bv = self.json["bufferViews"][acc["bufferView"]]
        dt = np.dtype(_COMPONENT[acc["componentType"]])
        n = _NCOMP[acc["type"]]
        base = bv.get("byteOffset", 0) + acc.get("byteOffset", 0)
        stride = bv.get("byteStride", 0) or dt.itemsize * n
        count = acc["count"]
        if stride == dt.itemsize * n:
            arr = np.frombuffer(self.bin, dtype=dt, count=count * n, offset=base).reshape(count, n)
        else:
            raw = np.frombuffer(self.bin, dtype=np.uint8, count=stride * (count - 1) + dt.itemsize * n, offset=base)
            arr = np.stack([raw[i * stride:i * stride + dt.itemsize * n].view(dt) for i in range(count)])
        return arr.copy()


def _decode_image(data: bytes) -> np.ndarray:
    """gltf::import decodes with the `image` crate and keeps the source sample type; here PIL decodes
    (decoder LSB differences against the Rust JPEG decoder are possible — SURVEY appendix C)."""
    from PIL import Image as PILImage
    im = PILImage.open(io.BytesIO(data))
    if im.mode == "P":
        im = im.convert("RGBA" if "transparency" in im.info else "RGB")
    if im.mode in ("I;16", "I;16L", "I;16B"):
        return np.asarray(im, dtype=np.uint16)[:, :, None]
    if im.mode not in ("L", "LA", "RGB", "RGBA"):
        im = im.convert("RGB")
    arr = np.asarray(im, dtype=np.uint8)
    return arr if arr.ndim == 3 else arr[:, :, None]


def scene_from_gltf_file(path: str, raster_height: int = HEIGHT, raster_width: Optional[int] = None) -> Scene:
    """scene.rs:249-522. `raster_height` replaces the hard-coded HEIGHT=600; `raster_width` overrides
    the aspect-derived width (needed for BASELINE's 512x512 / 1080p / 4K configurations)."""
    with open(path, "rb") as fh:
        glb = _Glb(fh.read())
    js = glb.json
    b = SceneBuilder()
    sc = b.scene

    for img in js.get("images", []):
        sc.images.append(_decode_image(glb.view(img["bufferView"])))

    samplers = js.get("samplers", [])
    wrap_of = {33071: _ffi.WRAP_CLAMP, 33648: _ffi.WRAP_MIRROR, 10497: _ffi.WRAP_REPEAT}
    for tex in js.get("textures", []):
        smp = samplers[tex["sampler"]] if "sampler" in tex else {}
        wrap = wrap_of[smp.get("wrapS", 10497)]
        mn, mg = smp.get("minFilter"), smp.get("magFilter")
        if mn is None:
            filt = _ffi.FILTER_BILINEAR if mg == 9729 else _ffi.FILTER_NEAREST
        elif mn == 9728:
            filt = _ffi.FILTER_NEAREST
        elif mn == 9729:
            filt = _ffi.FILTER_BILINEAR
        elif mn == 9987:
            filt = _ffi.FILTER_TRILINEAR
        else:
            filt = _ffi.FILTER_NEAREST
        sc.textures.append(Texture(_ffi.TEXTURE_IMAGE, image=tex["source"], filter=filt, wrap=wrap))

    emissions = []
    for mat in js.get("materials", []):
        pbr = mat.get("pbrMetallicRoughness", {})
        fac = [float(f32(v)) for v in pbr.get("baseColorFactor", [1.0, 1.0, 1.0, 1.0])]
        bct = pbr.get("baseColorTexture")
        if bct is not None:
            base_id = bct["index"]
            if fac != [1.0, 1.0, 1.0, 1.0]:
                factor_id = b.add_constant_texture(fac)
                albedo = b.add_texture(Texture(_ffi.TEXTURE_SCALE, a=base_id, b=factor_id))
            else:
                albedo = base_id
        else:
            albedo = b.add_constant_texture(fac)
        # metallic-roughness textures are created (and ignored) exactly as the reference does, so
        # texture ids line up (scene.rs:368-404)
        metallic, roughness = pbr.get("metallicFactor", 1.0), pbr.get("roughnessFactor", 1.0)
        mrt = pbr.get("metallicRoughnessTexture")
        if mrt is not None:
            if metallic != 1.0 or roughness != 1.0:
                fid = b.add_constant_texture([0.0, roughness, metallic, 0.0])
                b.add_texture(Texture(_ffi.TEXTURE_SCALE, a=mrt["index"], b=fid))
        else:
            b.add_constant_texture([0.0, roughness, metallic, 0.0])
        b.add_material(Material(_ffi.MATERIAL_DIFFUSE, albedo=albedo))
        emissions.append([float(f32(v)) for v in mat.get("emissiveFactor", [0.0, 0.0, 0.0])])

    instancing = {}
    camera = None
    scene_gltf = js["scenes"][js.get("scene", 0)]
    plights = js.get("extensions", {}).get("KHR_lights_punctual", {}).get("lights", [])
    for ni in scene_gltf["nodes"]:
        node = js["nodes"][ni]
        m = _node_matrix(node)
        if "camera" in node:
            camera = Camera.from_gltf_camera_node(m, js["cameras"][node["camera"]], raster_height, raster_width)
        if "mesh" in node:
            tf = Transform(m)
            mi = node["mesh"]
            if mi in instancing:
                for si in instancing[mi]:
                    b.add_instance(si, tf)
            else:
                ids = []
                for prim in js["meshes"][mi]["primitives"]:
                    material_idx = prim.get("material", 0)
                    attrs = prim["attributes"]
                    verts = glb.accessor(attrs["POSITION"]).astype(f32)
                    idx = glb.accessor(prim["indices"]).astype(np.uint32).reshape(-1)
                    tris = idx[: (len(idx) // 3) * 3].reshape(-1, 3)
                    normals = glb.accessor(attrs["NORMAL"]).astype(f32)
                    uvs = None
                    if "TEXCOORD_0" in attrs:
                        raw = glb.accessor(attrs["TEXCOORD_0"])
                        if raw.dtype == np.uint8:
                            uvs = (raw.astype(f32) / f32(255.0)).astype(f32)
                        elif raw.dtype == np.uint16:
                            uvs = (raw.astype(f32) / f32(65535.0)).astype(f32)
                        else:
                            uvs = raw.astype(f32)
                    mesh = Mesh(verts, tris, normals, uvs)
                    em = emissions[material_idx] if emissions else [0.0, 0.0, 0.0]
                    ids.append(len(sc.shapes))
                    b.add_shape_with_transform(mesh, material_idx, tf, em if em != [0.0, 0.0, 0.0] else None)
                instancing[mi] = ids
        ext = node.get("extensions", {}).get("KHR_lights_punctual")
        if ext is not None:
            l = plights[ext["light"]]
            color = np.array(l.get("color", [1, 1, 1]), dtype=f32)
            inten = (color * f32(l.get("intensity", 1.0))).astype(f32)
            if l["type"] == "directional":
                # decomposed() rotation of the node (TRS nodes carry it directly)
                d = _quat_rotate(node.get("rotation", [0, 0, 0, 1]), [0.0, 0.0, -1.0])
                b.add_light(Light(_ffi.LIGHT_DIRECTION, a=tuple(float(x) for x in d), b=tuple(float(x) for x in inten)))
            elif l["type"] == "point":
                pos = m[:3, 3]
                b.add_light(Light(_ffi.LIGHT_POINT, a=tuple(float(x) for x in pos), b=tuple(float(x) for x in inten)))
    if camera is None:
        raise ValueError("Scene must have camera")
    b.add_camera(camera)
    return b.build()


# --------------------------------------------------------------------------------------------
# PLY (geometry/shapes/mesh.rs:86-170): binary little-endian / ascii, fan triangulation, degenerate
# triangles dropped.
# --------------------------------------------------------------------------------------------
def mesh_from_ply_bytes(data: bytes, swap_handedness: bool = False) -> Mesh:
    marker = data.find(b"end_header")
    if marker < 0:
        raise ValueError("not a PLY file: no end_header")
    end = data.index(b"\n", marker) + 1                  # tolerant of CRLF line ends ("end_header\r\n")
    header = data[:end].decode("ascii").splitlines()
    fmt = "ascii"
    elements = []
    for line in header:
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            elements.append({"name": tok[1], "count": int(tok[2]), "props": []})
        elif tok[0] == "property":
            elements[-1]["props"].append(tok[1:])
    ty = {"float": "<f4", "float32": "<f4", "double": "<f8", "uchar": "u1", "uint8": "u1", "int": "<i4", "uint": "<u4",
          "int32": "<i4", "uint32": "<u4", "short": "<i2", "ushort": "<u2", "char": "i1", "int8": "i1", "int16": "<i2", "uint16": "<u2",
          "float64": "<f8"}
    if fmt == "ascii":   # the bunny of the builtin scenes is binary; exported meshes referenced from .pbrt files are often ascii
        tokens = data[end:].split()
        pos = 0
        verts = faces = None
        for el in elements:
            if el["name"] == "vertex":
                dt = np.dtype([(p[1], ty[p[0]]) for p in el["props"]])
                nprop = len(el["props"])
                vals = np.array(tokens[pos:pos + nprop * el["count"]], dtype=np.float64).reshape(el["count"], nprop)
                pos += nprop * el["count"]
                verts = np.zeros(el["count"], dtype=dt)
                for k, prop in enumerate(el["props"]):
                    verts[prop[1]] = vals[:, k]
            elif el["name"] == "face":
                faces = []
                extra = len(el["props"]) - 1          # scalar properties after the index list are skipped
                for _ in range(el["count"]):
                    n = int(tokens[pos]); pos += 1
                    faces.append(np.array(tokens[pos:pos + n], dtype=np.int64).astype(np.uint32)); pos += n + extra
            else:
                for _ in range(el["count"]):
                    for prop in el["props"]:
                        if prop[0] == "list":
                            pos += 1 + int(tokens[pos])
                        else:
                            pos += 1
        fmt = None
    elif fmt != "binary_little_endian":
        raise NotImplementedError("big-endian PLY files are not supported")
    off = end
    if fmt is not None:
        verts = faces = None
    for el in (elements if fmt is not None else []):
        if el["name"] == "vertex":
            dt = np.dtype([(p[1], ty[p[0]]) for p in el["props"]])
            verts = np.frombuffer(data, dtype=dt, count=el["count"], offset=off)
            off += dt.itemsize * el["count"]
        else:   # face, or any other element: walk its properties so that offsets stay right; only the first list of `face` is kept
            is_face = el["name"] == "face"
            if is_face:
                faces = []
            fixed = all(p[0] != "list" for p in el["props"])
            if fixed and not is_face:
                off += sum(np.dtype(ty[p[0]]).itemsize for p in el["props"]) * el["count"]
                continue
            for _ in range(el["count"]):
                first_list = True
                for p in el["props"]:
                    if p[0] == "list":
                        cdt, idt = np.dtype(ty[p[1]]), np.dtype(ty[p[2]])
                        n = int(np.frombuffer(data, dtype=cdt, count=1, offset=off)[0])
                        off += cdt.itemsize
                        if is_face and first_list:
                            faces.append(np.frombuffer(data, dtype=idt, count=n, offset=off).astype(np.uint32))
                            first_list = False
                        off += idt.itemsize * n
                    else:
                        off += np.dtype(ty[p[0]]).itemsize
    names = verts.dtype.names
    v = np.stack([verts["x"], verts["y"], verts["z"]], axis=1).astype(f32)
    normals = np.stack([verts["nx"], verts["ny"], verts["nz"]], axis=1).astype(f32) if "nx" in names else None
    un = "u" if "u" in names else ("s" if "s" in names else None)      # mesh.rs:37-38 accepts (s, t) as well as (u, v)
    vn = "v" if "v" in names else ("t" if "t" in names else None)
    uvs = np.stack([verts[un], verts[vn]], axis=1).astype(f32) if un and vn else None
    tris = []
    for idx in faces:
        for i in range(1, len(idx) - 1):
            tri = (idx[0], idx[i + 1], idx[i]) if swap_handedness else (idx[0], idx[i], idx[i + 1])
            ab, ac = v[tri[1]] - v[tri[0]], v[tri[2]] - v[tri[0]]
            cr = cross(ab.astype(f32), ac.astype(f32))
            area = f32(0.5) * f32(np.sqrt(f32(cr[0] * cr[0]) + f32(cr[1] * cr[1]) + f32(cr[2] * cr[2])))
            if area == 0.0 or np.isnan(area):
                continue
            tris.append(tri)
    return Mesh(v, np.array(tris, dtype=np.uint32).reshape(-1, 3), normals, uvs)
