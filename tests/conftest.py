import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "hostsim"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def rc():
    import raytracing_cuda
    return raytracing_cuda


@pytest.fixture(scope="session")
def oracle():
    import oracle_py
    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def hostsim():
    import hostsim_py
    hostsim_py.build()
    return hostsim_py


def load_scene(name, width=None, height=None):
    import raytracing_cuda as rc
    sc = rc.Scene.load_npz(os.path.join(GOLDEN, "scenes", name + ".npz"))
    if width is not None:
        sc.camera = sc.camera.with_raster_size(width, height)
    return sc


def bunny_mesh():
    import numpy as np
    import raytracing_cuda as rc
    z = np.load(os.path.join(GOLDEN, "scenes", "bunny_mesh.npz"))
    return rc.Mesh(z["vertices"], z["tris"], z["normals"], None)


def texture_zoo_scene(width=160, height=160):
    """Test-only scene for the texture variants no importer of the reference produces together (texture.rs:235-459,
    materials/texture.rs:45-68, image.rs:56-121): a Cornell box whose five walls carry, with uvs running over [-1.5, 2.5]
    so that the wrap modes matter,
      floor      Mix(a = u16 RGB image Mirror / Bilinear, b = f32 RGB image Clamp / Nearest, c = Checker amount)
      ceiling    f32 RGBA image, Clamp / Bilinear
      left wall  u16 2-channel image (missing channels read 0), Mirror / Nearest, scaled by a constant
      right wall u8 RGB image, Mirror / Trilinear (non power-of-two: Lanczos3 pyramid) -> mip-level AOV
      back wall  Mix(Constant, Scale(u8 image Clamp / Bilinear, Constant), Constant amount 0.25): nesting depth 2
    lit by a point light."""
    import numpy as np
    import raytracing_cuda as rc
    from raytracing_cuda import _ffi
    T, M = rc.Texture, rc.Material
    b = rc.test_scenes.cornell_box()
    sc = b.scene
    rng = np.random.default_rng(11)
    yy, xx = np.meshgrid(np.linspace(0, 1, 24), np.linspace(0, 1, 40), indexing="ij")
    img_u16 = (np.stack([xx, yy, 0.5 + 0.5 * np.sin(9 * xx * yy)], axis=2) * 65535).astype(np.uint16)
    img_f32 = rng.random((16, 16, 3)).astype(np.float32)
    img_f32a = rng.random((20, 12, 4)).astype(np.float32)
    img_la16 = (rng.random((8, 8, 2)) * 65535).astype(np.uint16)
    img_u8 = (np.stack([np.abs(np.sin(7 * xx)), np.abs(np.cos(5 * yy)), xx * yy], axis=2) * 255).astype(np.uint8)   # 24 x 40: not a power of two
    ids = [b.add_image(i) for i in (img_u16, img_f32, img_f32a, img_la16, img_u8)]
    t_u16 = b.add_texture(T(_ffi.TEXTURE_IMAGE, image=ids[0], filter=_ffi.FILTER_BILINEAR, wrap=_ffi.WRAP_MIRROR))
    t_f32 = b.add_texture(T(_ffi.TEXTURE_IMAGE, image=ids[1], filter=_ffi.FILTER_NEAREST, wrap=_ffi.WRAP_CLAMP))
    t_chk = b.add_texture(T(_ffi.TEXTURE_CHECKER, value=(0, 0, 0, 0), value2=(1, 1, 1, 1)))
    t_floor = b.add_texture(T(_ffi.TEXTURE_MIX, a=t_u16, b=t_f32, c=t_chk))
    t_ceil = b.add_texture(T(_ffi.TEXTURE_IMAGE, image=ids[2], filter=_ffi.FILTER_BILINEAR, wrap=_ffi.WRAP_CLAMP))
    t_la = b.add_texture(T(_ffi.TEXTURE_IMAGE, image=ids[3], filter=_ffi.FILTER_NEAREST, wrap=_ffi.WRAP_MIRROR))
    t_k = b.add_constant_texture((0.9, 0.7, 0.5, 1.0))
    t_left = b.add_texture(T(_ffi.TEXTURE_SCALE, a=t_la, b=t_k))
    t_right = b.add_texture(T(_ffi.TEXTURE_IMAGE, image=ids[4], filter=_ffi.FILTER_TRILINEAR, wrap=_ffi.WRAP_MIRROR))
    t_u8c = b.add_texture(T(_ffi.TEXTURE_IMAGE, image=ids[4], filter=_ffi.FILTER_BILINEAR, wrap=_ffi.WRAP_CLAMP))
    t_sc = b.add_texture(T(_ffi.TEXTURE_SCALE, a=t_u8c, b=t_k))
    t_q = b.add_constant_texture((0.25, 0.25, 0.25, 0.25))
    t_white = b.add_constant_texture((0.8, 0.8, 0.8, 1.0))
    t_back = b.add_texture(T(_ffi.TEXTURE_MIX, a=t_white, b=t_sc, c=t_q))
    uv = np.array([(-1.5, -1.5), (2.5, -1.5), (2.5, 2.5), (-1.5, 2.5)], dtype=np.float32)
    for shape, tex in zip(sc.shapes[:5], (t_floor, t_ceil, t_left, t_right, t_back)):
        shape.shape.uvs = uv.copy()
        shape.material = b.add_material(M(_ffi.MATERIAL_DIFFUSE, albedo=tex))
    out = b.build()
    import math
    out.camera = rc.Camera.lookat_camera_perspective((0.0, 4.4, 0.4), (0, 0, 0.75), (0, 0, 1), False,
                                                     float(np.float32(37.8) * np.float32(math.pi / 180)), width, height)
    return out


def _trs(translate, axis, angle, scale):
    """object-to-world = T * R(axis, angle) * S as a Transform whose inverse is the matrix inverse (float64, rounded once)"""
    import numpy as np
    import raytracing_cuda as rc
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    R = np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)
    M = np.eye(4)
    M[:3, :3] = R @ np.diag(np.asarray(scale, dtype=np.float64))
    M[:3, 3] = translate
    f = M.astype(np.float32)
    return rc.test_scenes.Transform(f, np.linalg.inv(f.astype(np.float64)).astype(np.float32))


def instanced_bunnies_scene(width=128, height=96):
    """Test-only scene for mesh reuse (glTF: several nodes referencing one mesh -> several Transform primitives over ONE Basic
    primitive, scene/scene.rs:430-443, geometry.rs:92-136): the Cornell box of cb.glb (x, y in [-1, 1], z in [0, 1.5]) with the
    bunny mesh used by three instances — scaled + translated, rotated + scaled, and non-uniformly scaled + rotated about a
    tilted axis."""
    import raytracing_cuda as rc
    base = load_scene("cb", width, height)
    b = rc.SceneBuilder()
    b.scene = base
    tex = b.add_constant_texture((0.7, 0.55, 0.3, 1.0))
    mat = b.add_material(rc.Material(rc._ffi.MATERIAL_DIFFUSE, albedo=tex))
    shape_index = len(base.shapes)
    b.add_shape_with_transform(bunny_mesh(), mat, _trs((-0.5, -0.2, 0.0), (0, 0, 1), 0.0, (0.6, 0.6, 0.6)), None)
    b.add_instance(shape_index, _trs((0.45, 0.1, 0.0), (0, 0, 1), 1.1, (0.5, 0.5, 0.5)))
    b.add_instance(shape_index, _trs((0.0, 0.45, 0.1), (0.3, 0.2, 1.0), -0.7, (0.4, 0.7, 0.5)))
    return b.build()
