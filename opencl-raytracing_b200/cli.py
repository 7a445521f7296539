"""`cli --backend cuda` — the caller side of the boundary, mirroring crates/cli/src/main.rs.

Same flag surface and file contract as the reference driver so that `rttest` (visual-testing/) can drive this backend
the way it drives `dist/cli` (runner.py:102-124): `--scene-path | --scene-name`, `-o`, `--output-format`, `-d`, `-s`,
`-l`, `--sampler`, sub-commands `full [--aov n,u,a,m] [--no-beauty]`, `pixel X Y [count] [offset]`, `list-scenes`;
outputs land under `scenes/output/` unless `--output-dir` says otherwise; EXR channel names / PNG conversions as in
main.rs:345-467 and raytracing-cpu/src/utils.rs. Additions: `--backend cuda` is the only backend (there is no CPU
path), `--width/--height` override the camera raster (the BASELINE configs need 512x512 / 1080p / 4K, which the
reference CLI cannot express), `--gpu` selects the device, `--num-threads` is rejected like it is for OptiX.

    python -m raytracing_cuda.cli --scene-path scenes/cb.glb -s 64 -l 1 --width 512 --height 512 -o cb.exr full --aov n,u
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from typing import List, Optional

import numpy as np

from .renderer import AovFlags, RaytracerSettings, Sampler
from .backend import CudaBackendSettings, CudaRenderer
from .scene import scene_from_gltf_file
from .pbrt import scene_from_pbrt_file
from . import exr, test_scenes


def backend_settings_from_args(args) -> CudaBackendSettings:
    """--gpu N: one device; --devices a,b,c: the multi-device context of ABI v3 (the CLI arm of INTEGRATION.md)"""
    if getattr(args, "devices", None):
        ids = [int(x) for x in args.devices.split(",") if x.strip() != ""]
        if len(ids) > 1:
            return CudaBackendSettings(num_devices=len(ids), device_ids=ids, tile_size=args.tile_size)
        return CudaBackendSettings(device_id=ids[0])
    return CudaBackendSettings(device_id=args.gpu)


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(prog="cli", description="B200 render backend driver (mirror of crates/cli)")
    g = p.add_mutually_exclusive_group()
    g.add_argument("--scene-path", help="Load a scene from disk (GLTF / GLB)")
    g.add_argument("--scene-name", help="Load a builtin test scene by name")
    p.add_argument("-o", "--output", help="Output filename (written under scenes/output/)")
    p.add_argument("--output-format", choices=["png", "exr"], help="Force output format (otherwise inferred from extension)")
    p.add_argument("--output-dir", default=os.path.join("scenes", "output"))
    p.add_argument("--backend", choices=["cuda"], default="cuda", help="Rendering backend")
    p.add_argument("-t", "--num-threads", type=int, help="CPU worker threads (rejected: not a CPU backend)")
    p.add_argument("-d", "--ray-depth", type=int, help="Maximum ray depth (bounces)")
    p.add_argument("-s", "--spp", type=int, help="Samples per pixel")
    p.add_argument("-l", "--light-samples", type=int, help="Light sample count")
    p.add_argument("--sampler", choices=["independent", "stratified"], help="Sampler type")
    p.add_argument("--width", type=int, help="Override the camera raster width")
    p.add_argument("--height", type=int, help="Override the camera raster height")
    p.add_argument("--gpu", type=int, default=0, help="CUDA device")
    p.add_argument("--devices", help="Comma-separated CUDA devices: one call renders the frame on all of them (tiles dealt inside the "
                                     "library, rtcuda_backend_settings.num_devices); the planes live on / return from the first one")
    p.add_argument("--tile-size", type=int, default=0, help="Tile edge of the multi-device deal (0 = 64, the reference's RenderTile)")
    sub = p.add_subparsers(dest="command")
    full = sub.add_parser("full", help="Full frame render with AOV control")
    full.add_argument("--aov", help="Comma-separated AOV list (e.g. normal,uv or n,u)")
    full.add_argument("--no-beauty", action="store_true", help="Disable beauty output")
    pix = sub.add_parser("pixel", help="Render a single pixel and print diagnostics")
    pix.add_argument("x", type=int)
    pix.add_argument("y", type=int)
    pix.add_argument("sample_count", type=int, nargs="?")
    pix.add_argument("sample_offset", type=int, nargs="?")
    sub.add_parser("list-scenes", help="List all builtin test scenes as JSON")
    return p


def _vec(v) -> str:
    return "(" + ", ".join(repr(float(x)) for x in v) + ")"


def save_png(path: str, rgb: np.ndarray, exposure: float) -> None:
    """utils.rs:7-29: value / exposure * 255, clamped, truncated; linear data, gamma 1.0 recorded in the file."""
    from PIL import Image, PngImagePlugin
    data = np.clip(rgb.astype(np.float32) / np.float32(exposure) * np.float32(255.0), 0.0, 255.0).astype(np.uint8)
    info = PngImagePlugin.PngInfo()
    Image.fromarray(data, "RGB").save(path, pnginfo=info, gamma=1.0)


def save_render_output(out, outputs: AovFlags, fmt: Optional[str], path: str) -> List[str]:
    """main.rs:296-467: one EXR with every plane, or one PNG per plane (beauty un-suffixed, exposure 1000)."""
    ext = os.path.splitext(path)[1].lower().lstrip(".")
    if fmt is None:
        fmt = ext if ext in ("png", "exr") else "exr"
        if ext not in ("png", "exr"):
            print(f"warning: extension not recognized, defaulting to exr", file=sys.stderr)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    written = []
    if fmt == "exr":
        exr.write_exr(path, exr.channels_of_render_output(out, outputs))
        return [path]
    stem = os.path.splitext(path)[0]
    if outputs & AovFlags.BEAUTY and out.beauty is not None:
        save_png(path, out.beauty, 1000.0)
        written.append(path)
    if outputs & AovFlags.NORMALS and out.normals is not None:
        save_png(stem + "_NORMALS.png", (out.normals + 1.0) / 2.0, 1.0)
        written.append(stem + "_NORMALS.png")
    if outputs & AovFlags.ALBEDO and out.albedo is not None:
        save_png(stem + "_ALBEDO.png", out.albedo, 1.0)
        written.append(stem + "_ALBEDO.png")
    if outputs & AovFlags.UV_COORDS and out.uv is not None:
        uv = np.concatenate([out.uv, np.zeros_like(out.uv[..., :1])], axis=-1)
        save_png(stem + "_UV_COORDS.png", uv, 1.0)
        written.append(stem + "_UV_COORDS.png")
    if outputs & AovFlags.MIP_LEVEL:
        print("warning: MIP_LEVEL png output not supported (yet)", file=sys.stderr)
    return written


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    if args.command == "list-scenes":
        print(json.dumps([t.name for t in test_scenes.all_test_scenes()]))
        return 0
    if not args.scene_path and not args.scene_name:
        print("error: either --scene-path or --scene-name is required", file=sys.stderr)
        return 1
    if args.num_threads is not None:
        print("error: --threads is not supported with the cuda backend", file=sys.stderr)
        return 1
    settings = RaytracerSettings()
    if args.scene_path:
        ext = os.path.splitext(args.scene_path)[1].lower()
        if ext == ".pbrt":   # crates/cli/src/main.rs:146-157: dispatch on the extension
            scene = scene_from_pbrt_file(args.scene_path)
        else:
            if ext not in (".gltf", ".glb"):
                print(f"warning: unrecognized file extension {ext!r}, trying to import as gltf", file=sys.stderr)
            scene = scene_from_gltf_file(args.scene_path)
    else:
        found = [t for t in test_scenes.all_test_scenes() if t.name == args.scene_name]
        if not found:
            print("error: failed to find scene", file=sys.stderr)
            return 1
        settings = found[0].settings_func()
        scene = found[0].scene_func()
    if args.width or args.height:
        scene.camera = scene.camera.with_raster_size(args.width or scene.camera.raster_width, args.height or scene.camera.raster_height)
    if args.ray_depth is not None:
        settings.max_ray_depth = args.ray_depth
    if args.light_samples is not None:
        settings.light_sample_count = args.light_samples
    if args.spp is not None:
        settings.samples_per_pixel = args.spp
    settings.accumulate_bounces = True
    if args.sampler == "independent":
        settings.sampler = Sampler.independent()
    elif args.sampler == "stratified":
        strata = int(np.ceil(np.sqrt(np.float32(settings.samples_per_pixel))))
        settings.sampler = Sampler.stratified(True, strata, strata)

    backend = backend_settings_from_args(args)
    if args.command == "pixel":
        low = args.sample_offset or 0
        high = low + (args.sample_count if args.sample_count is not None else 1)
        with CudaRenderer(scene, backend) as r:
            outs = r.render_pixel(settings, args.x, args.y, low, high)
        for o in outs:
            print(f"sample {o.sample_index}")
            print(f"hit: {str(o.hit).lower()}")
            print(f"uv: {_vec(o.uv)}")
            print(f"normal: {_vec(o.normal)}")
            print(f"radiance: {_vec(o.radiance)}")
        return 0

    outputs = AovFlags(settings.outputs)
    if args.command == "full":
        for a in (args.aov.split(",") if args.aov else []):
            a = a.strip()
            if a in ("n", "normal"):
                outputs |= AovFlags.NORMALS
            elif a in ("a", "albedo"):
                outputs |= AovFlags.ALBEDO
            elif a in ("u", "uv"):
                outputs |= AovFlags.UV_COORDS
            elif a in ("m", "mip"):
                outputs |= AovFlags.MIP_LEVEL
            elif a in ("b", "beauty"):
                print("warning: beauty is implicit", file=sys.stderr)
            else:
                print(f"warning: unknown AOV specified: {a}", file=sys.stderr)
        if args.no_beauty:
            outputs &= ~AovFlags.BEAUTY
    settings.outputs = outputs
    if not int(outputs):
        print("warning: no outputs specified (--no-beauty, and no AOVs), quitting...", file=sys.stderr)
        return 0
    t0 = time.time()
    with CudaRenderer(scene, backend) as r:
        out = r.render(settings)
        stats = r.stats()
    if outputs & AovFlags.BEAUTY:
        print(f"finished rendering beauty in {stats['render_ms'] / 1e3:.3f} seconds", file=sys.stderr)   # lib.rs:810-811
    bad = 0 if out.beauty is None else int((~np.isfinite(out.beauty)).sum())
    if bad:
        print(f"warning: {bad} non-finite beauty channel values (lib.rs:813-854)", file=sys.stderr)
    path = os.path.join(args.output_dir, args.output or "output.exr")
    for w in save_render_output(out, outputs, args.output_format, path):
        print(f"wrote {w} ({time.time() - t0:.2f} s)", file=sys.stderr)
    return 0


if __name__ == "__main__":
    sys.exit(main())
