"""PBRT-v4 scene import — the subset `scene_from_pbrt_file` understands (crates/raytracing/src/scene/pbrt.rs:1271-1411),
the second file format the CLI accepts (`--scene-path x.pbrt`, crates/cli/src/main.rs:151). Caller side of the backend
boundary: it produces the same `Scene` the glTF importer and the builtin scenes produce, which `to_desc()` flattens
into the C ABI.

What the reference's importer does, and this one therefore does too:
  * tokens: whitespace / `#` comments, quoted strings, `[` `]`, bare words (pbrt.rs:301-376);
  * parameter lists `"type name" value | [values]` for integer, float, point2, point3/point, vector3/vector,
    normal3/normal, rgb/color, spectrum (read as rgb), bool, string, texture (pbrt.rs:410-607);
  * transforms: Identity, LookAt (left-handed: handedness swap, composed as the INVERSE look-at), Translate, Scale,
    Rotate (degrees), Transform / ConcatTransform (column-major), each composed "current first, then new"
    (pbrt.rs:609-706); Attribute/TransformBegin/End both push / pop (transform, material, pending area light);
  * Film x/yresolution (640x480 default), Camera perspective (fov default 90) / orthographic, rebuilt as a look-at
    camera from the inverse CTM (pbrt.rs:708-784);
  * materials diffuse / conductor / dielectric / coateddiffuse with the roughness rules of extract_roughness
    (pbrt.rs:834-952), named materials, textures constant / imagemap (bilinear, repeat) / scale (a constant!) /
    checkerboard (pbrt.rs:985-1057);
  * shapes sphere / trianglemesh / plymesh (clockwise winding) / disk (placeholder sphere) (pbrt.rs:1059-1161);
  * lights point / distant / spot (as point) and AreaLightSource diffuse attached to the NEXT shape (pbrt.rs:1163-1255);
  * Include; Sampler / Integrator / PixelFilter / Accelerator / ColorSpace skipped; instancing and media ignored.
A scene without a camera or without lights is an error (the reference hits `todo!()` there, pbrt.rs:1286-1294).
"""
from __future__ import annotations

import math
import os
import sys
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _ffi
from .geometry import Transform, f32, unit
from .scene import Camera, Light, Material, Mesh, Scene, SceneBuilder, Sphere, Texture, _decode_image, mesh_from_ply_bytes


class PbrtParseError(ValueError):
    """ParseError (pbrt.rs:65-97)."""


def _warn(msg: str) -> None:
    print(f"warning: {msg}", file=sys.stderr)


# ---------------------------------------------------------------------------------------------
# tokens (pbrt.rs:301-376)
# ---------------------------------------------------------------------------------------------
class _Tokens:
    _WS = b" \t\n\r\x0b\x0c"

    def __init__(self, text: str):
        self.b = text.encode("utf-8")
        self.pos = 0

    def _skip(self) -> None:
        b, n = self.b, len(self.b)
        while self.pos < n:
            c = b[self.pos]
            if c == 0x23:  # '#'
                while self.pos < n and b[self.pos] != 0x0A:
                    self.pos += 1
            elif c in self._WS:
                self.pos += 1
            else:
                break

    def peek(self) -> Optional[str]:
        self._skip()
        b, n, start = self.b, len(self.b), self.pos
        if start >= n:
            return None
        if b[start] == 0x22:  # '"'
            end = start + 1
            while end < n and b[end] != 0x22:
                end += 1
            if end < n:
                end += 1
        elif b[start] in b"[]":
            end = start + 1
        else:
            end = start
            while end < n and b[end] not in self._WS and b[end] not in b'[]"#':
                end += 1
        return b[start:end].decode("utf-8")

    def next(self) -> Optional[str]:
        tok = self.peek()
        if tok is not None:
            self.pos += len(tok.encode("utf-8"))
        return tok

    def expect(self, what: str) -> None:
        tok = self.next()
        if tok is None:
            raise PbrtParseError("unexpected end of file")
        if tok != what:
            raise PbrtParseError(f"expected '{what}', got '{tok}'")


def _quoted(tok: Optional[str]) -> str:
    if tok is None:
        raise PbrtParseError("unexpected end of file")
    if len(tok) >= 2 and tok[0] == '"' and tok[-1] == '"':
        return tok[1:-1]
    raise PbrtParseError("bad string")


def _float(tok: Optional[str]) -> np.float32:
    if tok is None:
        raise PbrtParseError("unexpected end of file")
    try:
        return f32(float(tok))
    except ValueError:
        raise PbrtParseError(f"bad float: {tok}")


def _int(tok: Optional[str]) -> int:
    try:
        return int(tok)
    except (TypeError, ValueError):
        raise PbrtParseError(f"bad integer: {tok}")


# ---------------------------------------------------------------------------------------------
# parameter lists (pbrt.rs:100-235, 410-607). A value is (kind, payload); a one-element list collapses to a scalar
# of its kind exactly as the reference's ParameterValue does, and the getters accept the same kinds.
# ---------------------------------------------------------------------------------------------
_TUPLE = {"point2": ("point2", 2), "point3": ("point3", 3), "point": ("point3", 3), "vector3": ("vector3", 3), "vector": ("vector3", 3),
          "normal3": ("normal3", 3), "normal": ("normal3", 3)}


class _Params:
    def __init__(self):
        self.items: List[tuple] = []   # (name, kind, value); kind ends with 's' for lists

    def get(self, name: str):
        for n, k, v in self.items:
            if n == name:
                return k, v
        return None

    def get_float(self, name: str):
        kv = self.get(name)
        if kv is None:
            return None
        k, v = kv
        if k == "float":
            return v
        if k in ("floats", "integers") and len(v):
            return f32(v[0])
        if k == "integer":
            return f32(v)
        return None

    def get_float_or(self, name, default):
        v = self.get_float(name)
        return f32(default) if v is None else v

    def get_integer_or(self, name, default):
        kv = self.get(name)
        if kv is not None:
            k, v = kv
            if k == "integer":
                return v
            if k == "integers" and len(v):
                return v[0]
        return default

    def get_integers(self, name):
        kv = self.get(name)
        return kv[1] if kv is not None and kv[0] == "integers" else None

    def _tuples(self, name, kind):
        kv = self.get(name)
        if kv is None:
            return None
        k, v = kv
        if k == kind + "s":
            return v
        if k == kind:
            return [v]
        return None

    def get_point3(self, name):
        v = self._tuples(name, "point3")
        return v[0] if v else None

    def get_point3s(self, name):
        return self._tuples(name, "point3")

    def get_normal3s(self, name):
        return self._tuples(name, "normal3")

    def get_point2s(self, name):
        return self._tuples(name, "point2")

    def get_floats(self, name):
        kv = self.get(name)
        if kv is None:
            return None
        k, v = kv
        return v if k == "floats" else ([v] if k == "float" else None)

    def get_rgb(self, name):
        kv = self.get(name)
        if kv is None:
            return None
        k, v = kv
        if k == "rgb" or (k == "floats" and len(v) >= 3):
            return (f32(v[0]), f32(v[1]), f32(v[2]))
        return None

    def get_rgb_or(self, name, default):
        v = self.get_rgb(name)
        return tuple(f32(x) for x in default) if v is None else v

    def get_string(self, name):
        kv = self.get(name)
        return kv[1] if kv is not None and kv[0] == "string" else None

    def get_texture(self, name):
        kv = self.get(name)
        return kv[1] if kv is not None and kv[0] == "texture" else None

    def get_bool(self, name):
        kv = self.get(name)
        return kv[1] if kv is not None and kv[0] == "bool" else None


def _param_value(toks: _Tokens, ptype: str):
    brackets = toks.peek() == "["
    if brackets:
        toks.next()

    def more() -> bool:
        tok = toks.peek()
        return tok is not None and tok != "]" and not tok.startswith('"')

    def collect(read):
        vals = []
        while more():
            vals.append(read())
            if not brackets:
                break
        return vals

    if ptype == "integer":
        vals = collect(lambda: _int(toks.next()))
        value = ("integer", vals[0]) if len(vals) == 1 else ("integers", vals)
    elif ptype == "float":
        vals = collect(lambda: _float(toks.next()))
        value = ("float", vals[0]) if len(vals) == 1 else ("floats", vals)
    elif ptype in _TUPLE:
        kind, n = _TUPLE[ptype]
        vals = collect(lambda: tuple(_float(toks.next()) for _ in range(n)))
        value = (kind, vals[0]) if len(vals) == 1 else (kind + "s", vals)
    elif ptype in ("rgb", "color", "spectrum"):
        if ptype == "spectrum":
            _warn("spectrum parameters not fully supported, treating as RGB")
        value = ("rgb", tuple(_float(toks.next()) for _ in range(3)))
    elif ptype == "bool":
        tok = toks.next()
        if tok is None:
            raise PbrtParseError("unexpected end of file")
        clean = tok.strip('"')
        if clean not in ("true", "false"):
            raise PbrtParseError(f"bad bool: {tok}")
        value = ("bool", clean == "true")
    elif ptype in ("string", "texture"):
        value = (ptype, _quoted(toks.next()))
    else:
        _warn(f"unknown parameter type '{ptype}', skipping")
        collect(lambda: toks.next())
        value = ("float", f32(0.0))
    if brackets:
        toks.expect("]")
    return value


def _param_list(toks: _Tokens) -> _Params:
    params = _Params()
    while True:
        tok = toks.peek()
        if tok is None or not tok.startswith('"'):
            break
        decl = _quoted(toks.next())
        parts = decl.split()
        if len(parts) != 2:
            raise PbrtParseError(f"bad parameter: {decl}")
        kind, value = _param_value(toks, parts[0])
        params.items.append((parts[1], kind, value))
    return params


# ---------------------------------------------------------------------------------------------
# parser state (pbrt.rs:237-299)
# ---------------------------------------------------------------------------------------------
@dataclass
class _State:
    ctm: Transform = field(default_factory=Transform.identity)
    stack: list = field(default_factory=list)
    film: tuple = (640, 480)
    named_materials: list = field(default_factory=list)
    named_textures: list = field(default_factory=list)
    material: Optional[int] = None
    area_light: Optional[tuple] = None
    has_camera: bool = False
    has_lights: bool = False

    def push(self):
        self.stack.append((self.ctm, self.material, self.area_light))

    def pop(self):
        if self.stack:
            self.ctm, self.material, self.area_light = self.stack.pop()
        else:
            _warn("AttributeEnd without matching AttributeBegin")

    def named_texture(self, name):
        return next((t for n, t in self.named_textures if n == name), None)

    def named_material(self, name):
        return next((m for n, m in self.named_materials if n == name), None)


def _apply_vector(t: Transform, v) -> np.ndarray:
    """Transform::apply_vector: the linear part only."""
    return (t.forward[:3, :3] @ np.asarray(v, dtype=f32)).astype(f32)


def _to_radians(deg) -> np.float32:
    return f32(f32(deg) * f32(math.pi / 180.0))   # f32::to_radians


def _floats(toks: _Tokens, n: int):
    return [_float(toks.next()) for _ in range(n)]


def _matrix(toks: _Tokens) -> Transform:
    toks.expect("[")
    m = np.array(_floats(toks, 16), dtype=f32).reshape(4, 4).T.copy()   # pbrt matrices are column-major
    toks.expect("]")
    return Transform.from_matrix(m)


def _camera(toks: _Tokens, st: _State, b: SceneBuilder) -> None:
    kind = _quoted(toks.next())
    params = _param_list(toks)
    c2w = st.ctm.invert()
    pos, target, up = c2w.apply_point((0, 0, 0)), c2w.apply_point((0, 0, 1)), _apply_vector(c2w, (0, 1, 0))
    w, h = st.film
    if kind == "orthographic":
        cam = Camera.lookat_camera_orthographic(pos, target, up, False, w, h, f32(1.0) / f32(min(w, h)))
    else:
        if kind != "perspective":
            _warn(f"unsupported camera type '{kind}', defaulting to perspective")
            fov = f32(90.0)
        else:
            fov = params.get_float_or("fov", 90.0)
        cam = Camera.lookat_camera_perspective(pos, target, up, False, float(_to_radians(fov)), w, h)
    b.add_camera(cam)
    st.has_camera = True


def _rgb_texture(st: _State, b: SceneBuilder, params: _Params, name: str, default) -> int:
    tex_name = params.get_texture(name)
    if tex_name is not None:
        t = st.named_texture(tex_name)
        if t is not None:
            return t
    c = params.get_rgb_or(name, default)
    return b.add_constant_texture((c[0], c[1], c[2], 1.0))


def _float_texture(st: _State, b: SceneBuilder, params: _Params, name: str, default: float) -> int:
    tex_name = params.get_texture(name)
    if tex_name is not None:
        t = st.named_texture(tex_name)
        if t is not None:
            return t
    v = params.get_float_or(name, default)
    return b.add_constant_texture((v, v, v, 1.0))


def _roughness(params: _Params, b: SceneBuilder, st: _State) -> Optional[int]:
    """extract_roughness (pbrt.rs:834-871): `roughness` xor (`uroughness` and `vroughness`), else smooth."""
    iso = params.get("roughness") is not None
    has_u, has_v = params.get("uroughness") is not None, params.get("vroughness") is not None
    if has_u != has_v:
        _warn("bad anisotropic roughness description; both u and v components are required. falling back to smooth")
        return None
    if iso and has_u:
        _warn("bad roughness description; both `roughness` and `uroughness/vroughness` descriptions provided. falling back to smooth")
        return None
    if iso:
        return _float_texture(st, b, params, "roughness", 0.0)
    if has_u:
        ax, ay = params.get_float("uroughness"), params.get_float("vroughness")
        if ax is None or ay is None:
            raise NotImplementedError("texture values for uroughness / vroughness (todo!() in the reference, pbrt.rs:864)")
        return b.add_constant_texture((ax, ay, 0.0, 0.0))
    return None


def _material(kind: str, params: _Params, st: _State, b: SceneBuilder) -> Material:
    """create_material (pbrt.rs:873-952)."""
    if kind == "diffuse":
        return Material(_ffi.MATERIAL_DIFFUSE, albedo=_rgb_texture(st, b, params, "reflectance", (0.5, 0.5, 0.5)))
    if kind == "conductor":
        eta = _rgb_texture(st, b, params, "eta", (0.2, 0.2, 0.2))
        k = _rgb_texture(st, b, params, "k", (3.0, 3.0, 3.0))
        rough = _roughness(params, b, st)
        if rough is None:
            return Material(_ffi.MATERIAL_SMOOTH_CONDUCTOR, eta=eta, kappa=k)
        remap = params.get_bool("remaproughness")
        return Material(_ffi.MATERIAL_ROUGH_CONDUCTOR, eta=eta, kappa=k, roughness=rough, remap_roughness=True if remap is None else remap)
    if kind == "dielectric":
        eta = b.add_constant_texture((params.get_float_or("eta", 1.5), 0.0, 0.0, 0.0))
        rough = _roughness(params, b, st)
        if rough is None:
            return Material(_ffi.MATERIAL_SMOOTH_DIELECTRIC, eta=eta)
        remap = params.get_bool("remaproughness")
        return Material(_ffi.MATERIAL_ROUGH_DIELECTRIC, eta=eta, roughness=rough, remap_roughness=True if remap is None else remap)
    if kind == "coateddiffuse":
        albedo = _rgb_texture(st, b, params, "reflectance", (0.5, 0.5, 0.5))
        eta = b.add_constant_texture((params.get_float_or("eta", 1.5), 0.0, 0.0, 0.0))
        rough = _roughness(params, b, st)
        remap = params.get_bool("remaproughness")
        thick = b.add_constant_texture((params.get_float_or("thickness", 0.01), 0.0, 0.0, 0.0))
        ca = params.get_rgb_or("albedo", (1.0, 1.0, 1.0))
        coat = b.add_constant_texture((ca[0], ca[1], ca[2], 1.0))
        return Material(_ffi.MATERIAL_COATED_DIFFUSE, albedo=albedo, eta=eta, roughness=_ffi.NONE if rough is None else rough,
                        thickness=thick, coat_albedo=coat, remap_roughness=True if remap is None else remap)
    _warn(f"unsupported material type '{kind}', defaulting to diffuse gray")
    return Material(_ffi.MATERIAL_DIFFUSE, albedo=b.add_constant_texture((0.5, 0.5, 0.5, 1.0)))


def _texture(toks: _Tokens, st: _State, b: SceneBuilder, base: str) -> None:
    """parse_texture_directive (pbrt.rs:985-1057)."""
    name = _quoted(toks.next())
    _quoted(toks.next())   # "spectrum" / "float"
    kind = _quoted(toks.next())
    params = _param_list(toks)
    magenta = Texture(_ffi.TEXTURE_CONSTANT, value=(1.0, 0.0, 1.0, 1.0))
    if kind == "constant":
        v = params.get_rgb_or("value", (1.0, 1.0, 1.0))
        tex = Texture(_ffi.TEXTURE_CONSTANT, value=(float(v[0]), float(v[1]), float(v[2]), 1.0))
    elif kind == "imagemap":
        fn = params.get_string("filename")
        tex = magenta
        if fn is None:
            _warn("imagemap texture missing filename")
        else:
            try:
                with open(os.path.join(base, fn), "rb") as f:
                    image = _decode_image(f.read())
                tex = Texture(_ffi.TEXTURE_IMAGE, image=b.add_image(image), filter=_ffi.FILTER_BILINEAR, wrap=_ffi.WRAP_REPEAT)
            except Exception as e:   # the reference warns and substitutes the error colour
                _warn(f"failed to load texture '{fn}': {e}")
    elif kind == "scale":
        s = float(params.get_float_or("scale", 1.0))
        tex = Texture(_ffi.TEXTURE_CONSTANT, value=(s, s, s, 1.0))
    elif kind == "checkerboard":
        t1, t2 = params.get_rgb_or("tex1", (0.0, 0.0, 0.0)), params.get_rgb_or("tex2", (1.0, 1.0, 1.0))
        tex = Texture(_ffi.TEXTURE_CHECKER, value=(float(t1[0]), float(t1[1]), float(t1[2]), 1.0),
                      value2=(float(t2[0]), float(t2[1]), float(t2[2]), 1.0))
    else:
        _warn(f"unsupported texture type '{kind}', using constant white")
        tex = Texture(_ffi.TEXTURE_CONSTANT, value=(1.0, 1.0, 1.0, 1.0))
    st.named_textures.append((name, b.add_texture(tex)))


def _shape(toks: _Tokens, st: _State, b: SceneBuilder, base: str) -> None:
    """parse_shape_directive (pbrt.rs:1059-1161)."""
    kind = _quoted(toks.next())
    params = _param_list(toks)
    material = st.material
    if material is None:   # a fresh grey diffuse per shape, like the reference
        material = b.add_material(Material(_ffi.MATERIAL_DIFFUSE, albedo=b.add_constant_texture((0.5, 0.5, 0.5, 1.0))))
    if kind in ("sphere", "disk"):
        if kind == "disk":
            _warn("disk shape not supported, creating placeholder sphere")
        shape = Sphere((0.0, 0.0, 0.0), float(params.get_float_or("radius", 1.0)))
    elif kind == "trianglemesh":
        P = params.get_point3s("P")
        if P is None:
            raise PbrtParseError("missing parameter: P")
        vertices = np.array(P, dtype=f32).reshape(-1, 3)
        idx = params.get_integers("indices")
        if idx is not None:
            n = len(idx) // 3
            if len(idx) % 3:
                raise PbrtParseError("trianglemesh indices not a multiple of 3")   # the reference panics on the short chunk
            tris = np.array(idx[:3 * n], dtype=np.int64).astype(np.uint32).reshape(-1, 3)
        else:
            tris = np.arange(3 * (len(vertices) // 3), dtype=np.uint32).reshape(-1, 3)
        N = params.get_normal3s("N")
        normals = np.array(N, dtype=f32).reshape(-1, 3) if N is not None else None
        uv2 = params.get_point2s("uv")
        if uv2 is not None:
            uvs = np.array(uv2, dtype=f32).reshape(-1, 2)
        else:
            flat = params.get_floats("uv")
            if flat is not None:
                flat = list(flat) + [f32(0.0)] * (len(flat) % 2)
                uvs = np.array(flat, dtype=f32).reshape(-1, 2)
            else:
                uvs = None
        shape = Mesh(vertices, tris, normals, uvs)
    elif kind == "plymesh":
        fn = params.get_string("filename")
        if fn is None:
            raise PbrtParseError("missing parameter: filename")
        try:
            with open(os.path.join(base, fn), "rb") as f:
                shape = mesh_from_ply_bytes(f.read(), True)   # pbrt meshes are wound clockwise
        except Exception as e:
            _warn(f"failed to load PLY file '{fn}': {e}")
            return
    else:
        _warn(f"unsupported shape type '{kind}', skipping")
        return
    if st.area_light is not None:
        st.has_lights = True
    b.add_shape_with_transform(shape, material, st.ctm, st.area_light)
    st.area_light = None


def _light(toks: _Tokens, st: _State, b: SceneBuilder) -> None:
    """parse_light_source_directive (pbrt.rs:1163-1237)."""
    kind = _quoted(toks.next())
    params = _param_list(toks)
    if kind in ("point", "spot"):
        I = params.get_rgb_or("I", (1.0, 1.0, 1.0))
        if kind == "point":
            s = params.get_float_or("scale", 1.0)
            I = tuple(f32(c * s) for c in I)
        else:
            _warn("spot light converted to point light")
        frm = params.get_point3("from") or (0.0, 0.0, 0.0)
        pos = st.ctm.apply_point(frm)
        b.add_light(Light(_ffi.LIGHT_POINT, a=tuple(float(x) for x in pos), b=tuple(float(x) for x in I)))
        st.has_lights = True
    elif kind == "distant":
        L = params.get_rgb_or("L", (1.0, 1.0, 1.0))
        s = params.get_float_or("scale", 1.0)
        frm = np.array(params.get_point3("from") or (0.0, 0.0, 1.0), dtype=f32)
        to = np.array(params.get_point3("to") or (0.0, 0.0, 0.0), dtype=f32)
        d = _apply_vector(st.ctm, unit((to - frm).astype(f32)))
        b.add_light(Light(_ffi.LIGHT_DIRECTION, a=tuple(float(x) for x in d), b=tuple(float(f32(c * s)) for c in L)))
        st.has_lights = True
    elif kind in ("infinite", "environment"):
        _warn("infinite/environment lights not supported")
    else:
        _warn(f"unsupported light type '{kind}', skipping")


def _skip(toks: _Tokens) -> None:
    tok = toks.peek()
    if tok is not None and tok.startswith('"'):
        toks.next()
    _param_list(toks)


def _parse(text: str, base: str, st: _State, b: SceneBuilder) -> None:
    """parse_pbrt_content (pbrt.rs:1299-1411)."""
    toks = _Tokens(text)
    while True:
        d = toks.next()
        if d is None or d == "WorldEnd":
            break
        if d == "Identity":
            st.ctm = Transform.identity()
        elif d == "LookAt":
            v = _floats(toks, 9)
            # pbrt is left-handed: the handedness swap lives in the camera-to-world look-at; the CTM takes its inverse
            st.ctm = st.ctm.compose(Transform.look_at(v[0:3], v[3:6], v[6:9], True).invert())
        elif d == "Translate":
            st.ctm = st.ctm.compose(Transform.translate(_floats(toks, 3)))
        elif d == "Scale":
            st.ctm = st.ctm.compose(Transform.scale(_floats(toks, 3)))
        elif d == "Rotate":
            v = _floats(toks, 4)
            st.ctm = st.ctm.compose(Transform.rotate(float(_to_radians(v[0])), v[1:4]))
        elif d == "Transform":
            st.ctm = _matrix(toks)
        elif d == "ConcatTransform":
            st.ctm = st.ctm.compose(_matrix(toks))
        elif d == "Film":
            _quoted(toks.next())
            p = _param_list(toks)
            st.film = (p.get_integer_or("xresolution", 640), p.get_integer_or("yresolution", 480))
        elif d == "Camera":
            _camera(toks, st, b)
        elif d == "Material":
            kind = _quoted(toks.next())
            st.material = b.add_material(_material(kind, _param_list(toks), st, b))
        elif d == "MakeNamedMaterial":
            name = _quoted(toks.next())
            p = _param_list(toks)
            st.named_materials.append((name, b.add_material(_material(p.get_string("type") or "diffuse", p, st, b))))
        elif d == "NamedMaterial":
            name = _quoted(toks.next())
            m = st.named_material(name)
            if m is None:
                _warn(f"unknown named material '{name}', using current material")
            else:
                st.material = m
        elif d == "Texture":
            _texture(toks, st, b, base)
        elif d == "Shape":
            _shape(toks, st, b, base)
        elif d == "LightSource":
            _light(toks, st, b)
        elif d == "AreaLightSource":
            kind = _quoted(toks.next())
            p = _param_list(toks)
            if kind == "diffuse":
                L, s = p.get_rgb_or("L", (1.0, 1.0, 1.0)), p.get_float_or("scale", 1.0)
                st.area_light = tuple(float(f32(c * s)) for c in L)
            else:
                _warn(f"unsupported area light type '{kind}', ignoring")
        elif d == "WorldBegin":
            st.ctm = Transform.identity()
        elif d in ("AttributeBegin", "TransformBegin"):
            st.push()
        elif d in ("AttributeEnd", "TransformEnd"):
            st.pop()
        elif d == "Include":
            path = os.path.join(base, _quoted(toks.next()))
            try:
                with open(path, "r") as f:
                    included = f.read()
            except OSError as e:
                raise PbrtParseError(f"{path}: {e}")
            _parse(included, os.path.dirname(path) or base, st, b)
        elif d in ("Sampler", "Integrator", "PixelFilter", "Accelerator", "ColorSpace"):
            _skip(toks)
        elif d == "ReverseOrientation":
            pass
        elif d in ("ObjectBegin", "ObjectEnd", "ObjectInstance"):
            if d != "ObjectEnd":
                _skip(toks)
            _warn("instancing (ObjectBegin/End/Instance) not supported")
        elif d in ("MediumInterface", "MakeNamedMedium"):
            _skip(toks)
            _warn("media/volumes not supported")
        elif d.startswith('"'):
            continue
        else:
            _warn(f"unknown directive '{d}', ignoring")


def scene_from_pbrt_string(text: str, base_path: str = ".") -> Scene:
    st, b = _State(), SceneBuilder()
    _parse(text, base_path, st, b)
    if not st.has_camera:
        raise PbrtParseError("no camera in scene")
    if not st.has_lights:
        raise PbrtParseError("no lights found in scene")
    return b.build()


def scene_from_pbrt_file(path: str) -> Scene:
    """pub fn scene_from_pbrt_file(filepath) -> Result<Scene, ParseError> (pbrt.rs:1271-1278)."""
    try:
        with open(path, "r") as f:
            text = f.read()
    except OSError as e:
        raise PbrtParseError(f"{path}: {e}")
    return scene_from_pbrt_string(text, os.path.dirname(os.path.abspath(path)))
