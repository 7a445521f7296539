"""Minimal OpenEXR scanline reader / writer for the render planes (f32, uncompressed, arbitrary channel names).

The reference CLI writes one EXR per render with channels R,G,B / Normal.X,Y,Z / Albedo.X,Y,Z / U / V / "Mip Level"
(crates/cli/src/main.rs:398-467) sorted alphabetically (`AnyChannels::sort`, crates/raytracing-cpu/src/utils.rs:119-121),
and `rttest` reads them back channel by channel (visual-testing/src/rttest/diff.py:20-62). cv2 can only address R/G/B/A
and the `OpenEXR` python module is not in this image, hence this small self-contained codec:
single-part scanline file, one scan line per chunk, NO_COMPRESSION, FLOAT (or UINT) pixels, increasing-y line order —
a subset every EXR reader accepts. The reader handles exactly what the writer produces (plus HALF channels).
"""
from __future__ import annotations

import struct
from typing import Dict, List, Tuple

import numpy as np

_MAGIC = 20000630
_PIXEL_DTYPES = {0: np.uint32, 1: np.float16, 2: np.float32}


def _attr(name: str, typ: str, payload: bytes) -> bytes:
    return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(payload)) + payload


def write_exr(path: str, channels: Dict[str, np.ndarray]) -> None:
    """channels: name -> [H, W] float32 (or uint32) plane. Channels are stored sorted by name, as the spec requires."""
    names = sorted(channels)
    assert names, "no channels"
    h, w = channels[names[0]].shape
    planes = []
    chlist = b""
    for n in names:
        a = np.ascontiguousarray(channels[n])
        assert a.shape == (h, w), f"channel {n}: shape {a.shape} != {(h, w)}"
        ptype = 0 if a.dtype == np.uint32 else 2
        if ptype == 2:
            a = a.astype(np.float32, copy=False)
        planes.append(a)
        chlist += n.encode() + b"\0" + struct.pack("<iBxxxii", ptype, 0, 1, 1)
    chlist += b"\0"
    box = struct.pack("<iiii", 0, 0, w - 1, h - 1)
    header = (_attr("channels", "chlist", chlist) + _attr("compression", "compression", b"\0") +
              _attr("dataWindow", "box2i", box) + _attr("displayWindow", "box2i", box) +
              _attr("lineOrder", "lineOrder", b"\0") + _attr("pixelAspectRatio", "float", struct.pack("<f", 1.0)) +
              _attr("screenWindowCenter", "v2f", struct.pack("<ff", 0.0, 0.0)) +
              _attr("screenWindowWidth", "float", struct.pack("<f", 1.0)) + b"\0")
    row_bytes = sum(4 * w for _ in planes)
    head = struct.pack("<ii", _MAGIC, 2) + header
    table_off = len(head)
    first = table_off + 8 * h
    offsets = np.arange(h, dtype=np.uint64) * np.uint64(8 + row_bytes) + np.uint64(first)
    # every chunk: y, byte count, then the channels' rows in name order
    body = np.empty((h, 8 + row_bytes), dtype=np.uint8)
    hdr = np.empty((h, 2), dtype="<i4")
    hdr[:, 0] = np.arange(h)
    hdr[:, 1] = row_bytes
    body[:, :8] = hdr.view(np.uint8).reshape(h, 8)
    off = 8
    for a in planes:
        body[:, off:off + 4 * w] = a.astype(a.dtype.newbyteorder("<"), copy=False).view(np.uint8).reshape(h, 4 * w)
        off += 4 * w
    with open(path, "wb") as f:
        f.write(head)
        f.write(offsets.astype("<u8").tobytes())
        f.write(body.tobytes())


def read_exr(path: str) -> Tuple[Dict[str, np.ndarray], int, int]:
    """-> ({channel name: [H, W] array}, width, height) for uncompressed scanline files."""
    data = open(path, "rb").read()
    magic, version = struct.unpack_from("<ii", data, 0)
    if magic != _MAGIC:
        raise ValueError("not an OpenEXR file")
    if version & 0x200 or version & 0x1000:
        raise ValueError("tiled / multi-part EXR files are not supported")
    pos = 8
    attrs = {}
    while data[pos] != 0:
        e = data.index(b"\0", pos)
        name = data[pos:e].decode()
        pos = e + 1
        e = data.index(b"\0", pos)
        typ = data[pos:e].decode()
        pos = e + 1
        (size,) = struct.unpack_from("<i", data, pos)
        pos += 4
        attrs[name] = (typ, data[pos:pos + size])
        pos += size
    pos += 1
    if attrs["compression"][1][0] != 0:
        raise ValueError("only uncompressed EXR files are supported")
    x0, y0, x1, y1 = struct.unpack("<iiii", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    chans: List[Tuple[str, int]] = []
    cl = attrs["channels"][1]
    p = 0
    while cl[p] != 0:
        e = cl.index(b"\0", p)
        name = cl[p:e].decode()
        (ptype,) = struct.unpack_from("<i", cl, e + 1)
        chans.append((name, ptype))
        p = e + 1 + 16
    offsets = np.frombuffer(data, dtype="<u8", count=h, offset=pos)
    out = {n: np.empty((h, w), dtype=_PIXEL_DTYPES[t]) for n, t in chans}
    for off in offsets:
        off = int(off)
        y, _size = struct.unpack_from("<ii", data, off)
        q = off + 8
        for n, t in chans:
            dt = np.dtype(_PIXEL_DTYPES[t]).newbyteorder("<")
            out[n][y - y0] = np.frombuffer(data, dtype=dt, count=w, offset=q)
            q += dt.itemsize * w
    return out, w, h


def channels_of_render_output(out, outputs) -> Dict[str, np.ndarray]:
    """The channel set of save_to_exr (crates/cli/src/main.rs:398-467) for a RenderOutput."""
    from .renderer import AovFlags
    ch: Dict[str, np.ndarray] = {}
    if outputs & AovFlags.BEAUTY and out.beauty is not None:
        for i, n in enumerate(("R", "G", "B")):
            ch[n] = out.beauty[..., i]
    if outputs & AovFlags.NORMALS and out.normals is not None:
        for i, n in enumerate(("Normal.X", "Normal.Y", "Normal.Z")):
            ch[n] = out.normals[..., i]
    if outputs & AovFlags.ALBEDO and out.albedo is not None:
        for i, n in enumerate(("Albedo.X", "Albedo.Y", "Albedo.Z")):
            ch[n] = out.albedo[..., i]
    if outputs & AovFlags.UV_COORDS and out.uv is not None:
        ch["U"], ch["V"] = out.uv[..., 0], out.uv[..., 1]
    if outputs & AovFlags.MIP_LEVEL and out.mip_level is not None:
        ch["Mip Level"] = out.mip_level
    return ch
