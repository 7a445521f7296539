"""PBRT importer (crates/raytracing/src/scene/pbrt.rs): tokens, parameter lists, directive semantics,
and that an imported scene IS the scene a SceneBuilder produces by hand (same pixels through the oracle / the CUDA backend)."""
import math
import os

import numpy as np
import pytest

import raytracing_cuda as rc
from raytracing_cuda import _ffi
from raytracing_cuda.geometry import Transform
from raytracing_cuda.pbrt import _Tokens, _param_list

HEADER = '''
Film "rgb" "integer xresolution" [ 96 ] "integer yresolution" [ 64 ]   # comment after a directive
LookAt 0 1 5   0 0.5 0   0 1 0
Camera "perspective" "float fov" 40
Sampler "halton" "integer pixelsamples" 16
Integrator "path"
WorldBegin
'''

SCENE = HEADER + '''
LightSource "point" "rgb I" [ 10 20 30 ] "float scale" 2 "point3 from" [ 1 3 1 ]
LightSource "distant" "rgb L" [ 1 1 1 ] "point3 from" [ 0 5 0 ] "point3 to" [ 0 0 0 ]
LightSource "spot" "rgb I" [ 5 5 5 ] "point3 from" [ 0 4 0 ]
LightSource "infinite" "rgb L" [ 1 1 1 ]
Texture "chk" "spectrum" "checkerboard" "rgb tex1" [ 0.1 0.1 0.1 ] "rgb tex2" [ 0.9 0.9 0.9 ]
Texture "half" "float" "scale" "float scale" 0.5
MakeNamedMaterial "gold" "string type" "conductor" "rgb eta" [ 0.2 0.9 1.1 ] "rgb k" [ 3.9 2.4 2.2 ] "float roughness" 0.2
MakeNamedMaterial "glass" "string type" "dielectric" "float eta" 1.33
AttributeBegin
    Translate 0 -0.5 0
    Scale 4 1 4
    Material "diffuse" "texture reflectance" "chk"
    Shape "trianglemesh" "point3 P" [ -1 0 -1  1 0 -1  1 0 1  -1 0 1 ] "integer indices" [ 0 1 2 0 2 3 ]
        "normal N" [ 0 1 0 0 1 0 0 1 0 0 1 0 ] "point2 uv" [ 0 0 4 0 4 4 0 4 ]
AttributeEnd
AttributeBegin
    NamedMaterial "gold"
    Translate -1 0.5 0
    Rotate 90 0 1 0
    Shape "sphere" "float radius" 0.5
    TransformBegin
        Translate 0 0 -2
        NamedMaterial "glass"
        Shape "sphere" "float radius" [ 0.25 ]
    TransformEnd
    Shape "disk" "float radius" 0.1
AttributeEnd
AttributeBegin
    AreaLightSource "diffuse" "rgb L" [ 4 4 4 ] "float scale" 0.5
    ConcatTransform [ 1 0 0 0  0 1 0 0  0 0 1 0  0 3 0 1 ]
    Material "coateddiffuse" "rgb reflectance" [ 0.7 0.1 0.1 ] "float uroughness" 0.1 "float vroughness" 0.3 "bool remaproughness" false
    Shape "trianglemesh" "point3 P" [ -0.5 0 -0.5  0.5 0 -0.5  0.5 0 0.5 ]
    Shape "trianglemesh" "point3 P" [ -0.5 0 -0.5  0.5 0 0.5  -0.5 0 0.5 ]
AttributeEnd
Shape "cone"
Shape "sphere"
WorldEnd
Shape "sphere"
'''


def test_tokens_and_parameter_lists():
    t = _Tokens('Foo "a b"[1 2]# c\n  -3.5e1 ] "x"')
    assert [t.next() for _ in range(8)] == ["Foo", '"a b"', "[", "1", "2", "]", "-3.5e1", "]"]
    assert t.peek() == '"x"' and t.next() == '"x"' and t.next() is None
    p = _param_list(_Tokens('"float a" 1 "float b" [ 1 2 ] "integer n" [ 7 ] "point3 P" [ 1 2 3 ] "rgb c" [ .1 .2 .3 ] '
                            '"bool f" "true" "bool g" false "string s" "hi" "texture t" "tex" "spectrum sp" [ 1 2 3 ] "blackbody bb" [ 6500 ] '
                            '"vector v" 1 2 3 Next'))
    assert p.get_float("a") == 1 and p.get_float("b") == 1 and p.get_floats("b") == [1, 2] and p.get_float("n") == 7
    assert p.get_integer_or("n", 0) == 7 and p.get_integer_or("zz", 5) == 5 and p.get_integers("n") is None
    assert p.get_point3("P") == (1, 2, 3) and p.get_point3s("P") == [(1, 2, 3)]
    assert np.allclose(p.get_rgb("c"), (0.1, 0.2, 0.3)) and p.get_rgb("sp") == (1, 2, 3) and p.get_rgb("a") is None
    assert p.get_bool("f") is True and p.get_bool("g") is False and p.get_string("s") == "hi" and p.get_texture("t") == "tex"
    assert p.get_float("bb") == 0 and p.get("v")[0] == "vector3" and p.get_string("t") is None
    with pytest.raises(rc.PbrtParseError):
        _param_list(_Tokens('"float" 1'))
    with pytest.raises(rc.PbrtParseError):
        _param_list(_Tokens('"float a" [ 1 2'))
    with pytest.raises(rc.PbrtParseError):
        _param_list(_Tokens('"integer a" [ 1.5 ]'))


def test_directives_build_the_expected_scene():
    sc = rc.scene_from_pbrt_string(SCENE)
    cam = sc.camera
    assert (cam.raster_width, cam.raster_height) == (96, 64)
    ref = rc.Camera.lookat_camera_perspective((0, 1, 5), (0, 0.5, 0), (0, 1, 0), False, float(np.float32(40) * np.float32(math.pi / 180)), 96, 64)
    assert np.allclose(cam.camera_to_world.forward, ref.camera_to_world.forward, atol=1e-6)
    assert np.allclose(cam.raster_to_camera.forward, ref.raster_to_camera.forward, atol=1e-6)

    # lights in file order: point (I * scale), distant (unit direction), spot -> point (no scale), then the two area lights
    L = sc.lights
    assert [l.kind for l in L] == [_ffi.LIGHT_POINT, _ffi.LIGHT_DIRECTION, _ffi.LIGHT_POINT, _ffi.LIGHT_DIFFUSE_AREA]
    assert L[0].a == (1, 3, 1) and L[0].b == (20, 40, 60)
    assert np.allclose(L[1].a, (0, -1, 0)) and L[1].b == (1, 1, 1)
    assert L[2].a == (0, 4, 0) and L[2].b == (5, 5, 5)
    # AreaLightSource applies to the NEXT shape only (pbrt.rs:1153-1158)
    assert L[3].b == (2, 2, 2) and L[3].shape == 4
    assert [s.area_light for s in sc.shapes] == [None, None, None, None, 3, None, None]

    kinds = [type(s.shape).__name__ for s in sc.shapes]
    assert kinds == ["Mesh", "Sphere", "Sphere", "Sphere", "Mesh", "Mesh", "Sphere"]   # "cone" skipped, the shape after WorldEnd never read
    floor = sc.shapes[0].shape
    assert floor.tris.tolist() == [[0, 1, 2], [0, 2, 3]] and floor.normals.shape == (4, 3) and floor.uvs[2].tolist() == [4, 4]
    assert sc.shapes[4].shape.tris.tolist() == [[0, 1, 2]] and sc.shapes[4].shape.normals is None
    assert sc.shapes[3].shape.radius == np.float32(0.1)   # disk placeholder

    # transforms: "current first, then new" (Transform::compose) and the attribute stack
    t_floor = Transform.identity().compose(Transform.translate((0, -0.5, 0))).compose(Transform.scale((4, 1, 4)))
    assert np.array_equal(sc.instances[0][1].forward, t_floor.forward)
    t_gold = Transform.translate((-1, 0.5, 0)).compose(Transform.rotate(float(np.float32(90) * np.float32(math.pi / 180)), (0, 1, 0)))
    assert np.array_equal(sc.instances[1][1].forward, t_gold.forward)
    assert np.array_equal(sc.instances[2][1].forward, t_gold.compose(Transform.translate((0, 0, -2))).forward)
    assert np.array_equal(sc.instances[3][1].forward, t_gold.forward)           # TransformEnd restored the CTM
    assert np.allclose(sc.instances[4][1].forward[:3, 3], (0, 3, 0))            # column-major ConcatTransform
    assert np.array_equal(sc.instances[6][1].forward, np.eye(4, dtype=np.float32))

    M, T = sc.materials, sc.textures
    m_floor = M[sc.shapes[0].material]
    assert m_floor.kind == _ffi.MATERIAL_DIFFUSE and T[m_floor.albedo].kind == _ffi.TEXTURE_CHECKER
    assert T[m_floor.albedo].value[:3] == pytest.approx((0.1, 0.1, 0.1)) and T[m_floor.albedo].value2[:3] == pytest.approx((0.9, 0.9, 0.9))
    gold = M[sc.shapes[1].material]
    assert gold.kind == _ffi.MATERIAL_ROUGH_CONDUCTOR and gold.remap_roughness and T[gold.roughness].value[0] == pytest.approx(0.2)
    assert T[gold.eta].value[:3] == pytest.approx((0.2, 0.9, 1.1)) and T[gold.kappa].value[:3] == pytest.approx((3.9, 2.4, 2.2))
    glass = M[sc.shapes[2].material]
    assert glass.kind == _ffi.MATERIAL_SMOOTH_DIELECTRIC and T[glass.eta].value[0] == pytest.approx(1.33)
    assert sc.shapes[3].material == sc.shapes[1].material                       # TransformEnd restored the material too
    coat = M[sc.shapes[4].material]
    assert coat.kind == _ffi.MATERIAL_COATED_DIFFUSE and not coat.remap_roughness
    assert T[coat.roughness].value[:2] == pytest.approx((0.1, 0.3)) and T[coat.thickness].value[0] == pytest.approx(0.01)
    assert T[coat.coat_albedo].value == (1, 1, 1, 1)
    last = M[sc.shapes[6].material]   # no current material outside the attribute blocks: a fresh grey diffuse
    assert last.kind == _ffi.MATERIAL_DIFFUSE and T[last.albedo].value == (0.5, 0.5, 0.5, 1.0)
    # the "scale" texture is a constant in the reference (pbrt.rs:1031-1036)
    half = [t for t in T if t.kind == _ffi.TEXTURE_CONSTANT and t.value == (0.5, 0.5, 0.5, 1.0)]
    assert half


def test_errors_and_include(tmp_path):
    with pytest.raises(rc.PbrtParseError, match="no camera"):
        rc.scene_from_pbrt_string('WorldBegin LightSource "point" Shape "sphere"')
    with pytest.raises(rc.PbrtParseError, match="no lights"):
        rc.scene_from_pbrt_string('Camera "perspective" WorldBegin Shape "sphere"')
    with pytest.raises(rc.PbrtParseError, match="missing parameter: P"):
        rc.scene_from_pbrt_string('Camera "perspective" WorldBegin LightSource "point" Shape "trianglemesh"')
    with pytest.raises(rc.PbrtParseError):
        rc.scene_from_pbrt_file(str(tmp_path / "absent.pbrt"))
    with pytest.raises(rc.PbrtParseError, match="bad float"):
        rc.scene_from_pbrt_string("Translate 1 2 x")
    # roughness rules of extract_roughness: u without v, or both descriptions => smooth
    sc = rc.scene_from_pbrt_string('Camera "orthographic" WorldBegin LightSource "point" '
                                   'Material "dielectric" "float uroughness" 0.1 Shape "sphere" '
                                   'Material "conductor" "float roughness" 0.1 "float uroughness" 0.1 "float vroughness" 0.1 Shape "sphere"')
    assert [m.kind for m in sc.materials] == [_ffi.MATERIAL_SMOOTH_DIELECTRIC, _ffi.MATERIAL_SMOOTH_CONDUCTOR]
    assert sc.camera.kind == _ffi.CAMERA_ORTHOGRAPHIC and (sc.camera.raster_width, sc.camera.raster_height) == (640, 480)
    # Include is resolved against the including file's directory; a missing image becomes the magenta error colour
    (tmp_path / "geo").mkdir()
    (tmp_path / "geo" / "inner.pbrt").write_text('Texture "img" "spectrum" "imagemap" "string filename" "nope.png"\n'
                                                 'Material "diffuse" "texture reflectance" "img"\nShape "sphere" "float radius" 2\n')
    (tmp_path / "main.pbrt").write_text('Camera "perspective"\nWorldBegin\nLightSource "point"\nInclude "geo/inner.pbrt"\n')
    sc = rc.scene_from_pbrt_file(str(tmp_path / "main.pbrt"))
    assert len(sc.shapes) == 1 and sc.shapes[0].shape.radius == 2
    assert sc.textures[sc.materials[sc.shapes[0].material].albedo].value == (1.0, 0.0, 1.0, 1.0)


def _hand_built():
    """the scene of SCENE_SMALL written against the SceneBuilder"""
    b = rc.SceneBuilder()
    b.add_camera(rc.Camera.lookat_camera_perspective((0, 0, 4), (0, 0, 0), (0, 1, 0), False, float(np.float32(45) * np.float32(math.pi / 180)), 80, 60))
    b.add_light(rc.Light(_ffi.LIGHT_POINT, a=(2.0, 2.0, 2.0), b=(100.0, 100.0, 100.0)))
    red = b.add_material(rc.Material(_ffi.MATERIAL_DIFFUSE, albedo=b.add_constant_texture((0.8, 0.2, 0.2, 1.0))))
    b.add_shape_with_transform(rc.Sphere((0.0, 0.0, 0.0), 1.0), red, Transform.identity().compose(Transform.translate((0.5, 0, 0))), None)
    grey = b.add_material(rc.Material(_ffi.MATERIAL_DIFFUSE, albedo=b.add_constant_texture((0.5, 0.5, 0.5, 1.0))))
    floor = rc.Mesh(np.array([[-5, -1, -5], [5, -1, -5], [5, -1, 5], [-5, -1, 5]], dtype=np.float32), np.array([[0, 1, 2], [0, 2, 3]], dtype=np.uint32),
                    np.array([[0, 1, 0]] * 4, dtype=np.float32), None)
    b.add_shape_with_transform(floor, grey, Transform.identity(), None)
    return b.build()


SCENE_SMALL = '''
Film "rgb" "integer xresolution" [ 80 ] "integer yresolution" [ 60 ]
LookAt 0 0 4  0 0 0  0 1 0
Camera "perspective" "float fov" [ 45 ]
WorldBegin
LightSource "point" "rgb I" [ 100 100 100 ] "point3 from" [ 2 2 2 ]
AttributeBegin
    Translate 0.5 0 0
    Material "diffuse" "rgb reflectance" [ 0.8 0.2 0.2 ]
    Shape "sphere" "float radius" [ 1 ]
AttributeEnd
AttributeBegin
    Material "diffuse" "rgb reflectance" [ 0.5 0.5 0.5 ]
    Shape "trianglemesh" "point3 P" [ -5 -1 -5  5 -1 -5  5 -1 5  -5 -1 5 ] "normal N" [ 0 1 0  0 1 0  0 1 0  0 1 0 ] "integer indices" [ 0 1 2  0 2 3 ]
AttributeEnd
'''


def test_imported_scene_renders_like_the_hand_built_one(oracle):
    A = rc.AovFlags
    st = rc.RaytracerSettings(outputs=A.BEAUTY | A.NORMALS | A.DEBUG_IDS | A.DEBUG_DEPTH, samples_per_pixel=4)
    a, _ = oracle.render(rc.scene_from_pbrt_string(SCENE_SMALL), st, num_threads=4)
    b, _ = oracle.render(_hand_built(), st, num_threads=4)
    # the look-at camera is rebuilt from the inverse CTM (pbrt.rs:730-744): same camera up to rounding of that round trip
    assert (a.debug_ids != b.debug_ids).any(axis=-1).mean() < 2e-3
    assert np.abs(a.debug_depth - b.debug_depth).max() < 1e-3 or (np.abs(a.debug_depth - b.debug_depth) > 1e-3).mean() < 2e-3
    assert abs(a.beauty.mean() - b.beauty.mean()) < 2e-3 * b.beauty.mean()
    # orientation: +y up (floor at the bottom), +x to the right with the handedness swap, sphere centre 3.0 away
    ids, depth = a.debug_ids[..., 0], a.debug_depth
    ys, xs = np.nonzero((ids == 0) & (depth > 0))
    assert xs.mean() > 44 and abs(depth[(ids == 0) & (depth > 0)].min() - (math.sqrt(16.25) - 1.0)) < 0.01
    ys, _ = np.nonzero((ids == 1) & (depth > 0))
    assert ys.min() > 30


def test_reference_test_scene_if_present():
    path = "/root/reference/scenes/test.pbrt"
    if not os.path.exists(path):
        pytest.skip("reference tree not present on this box")
    sc = rc.scene_from_pbrt_file(path)
    assert (sc.camera.raster_width, sc.camera.raster_height) == (400, 400)
    assert [type(s.shape).__name__ for s in sc.shapes] == ["Sphere", "Sphere", "Mesh"]
    assert len(sc.lights) == 1 and sc.lights[0].b == (100, 100, 100) and len(sc.materials) == 3


@pytest.mark.gpu
def test_pbrt_scene_on_the_cuda_backend(oracle):
    from parity import assert_first_hit_parity, assert_beauty_parity, beauty_close, SPECULAR_GATES
    A = rc.AovFlags
    sc = rc.scene_from_pbrt_string(SCENE.replace('Shape "cone"', ""))
    st = rc.RaytracerSettings(outputs=A.BEAUTY | A.NORMALS | A.UV_COORDS | A.ALBEDO | A.MIP_LEVEL | A.DEBUG_IDS | A.DEBUG_DEPTH, samples_per_pixel=4,
                              max_ray_depth=4, light_sample_count=2)
    with rc.CudaRenderer(sc) as r:
        out = r.render(st)
    ref, _ = oracle.render(sc, st, num_threads=8)
    assert_first_hit_parity(out, ref)
    assert_beauty_parity(out.beauty, ref.beauty, **SPECULAR_GATES)


def test_plymesh_ascii_and_winding(tmp_path):
    """Shape "plymesh" (pbrt.rs:1119-1133): the PLY is read with the clockwise-winding swap; ascii files, quads split into fans"""
    (tmp_path / "quad.ply").write_text("ply\nformat ascii 1.0\nelement vertex 4\nproperty float x\nproperty float y\nproperty float z\n"
                                       "element face 1\nproperty list uchar int vertex_indices\nend_header\n"
                                       "0 0 0\n1 0 0\n1 1 0\n0 1 0\n4 0 1 2 3\n")
    (tmp_path / "s.pbrt").write_text('Camera "perspective"\nWorldBegin\nLightSource "point"\nShape "plymesh" "string filename" "quad.ply"\n'
                                     'Shape "plymesh" "string filename" "missing.ply"\n')
    sc = rc.scene_from_pbrt_file(str(tmp_path / "s.pbrt"))
    assert len(sc.shapes) == 1                                    # the missing file is skipped with a warning
    m = sc.shapes[0].shape
    assert m.vertices.shape == (4, 3) and m.tris.tolist() == [[0, 2, 1], [0, 3, 2]] and m.normals is None
