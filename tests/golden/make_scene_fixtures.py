"""Generates tests/golden/scenes/*.npz from the reference's scene assets (run in the build container,
where /root/reference exists; the GPU box only sees the committed .npz files).

Each fixture is the flattened scene (`Scene.save_npz`) produced by this repo's own glTF / PLY importers
from /root/reference/scenes/*.glb and crates/raytracing/src/scene/test_scenes/assets/bunny.ply, at the
importer's default raster (HEIGHT = 600, scene.rs:247). Tests re-derive other raster sizes with
`Camera.with_raster_size`.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import raytracing_cuda as rc  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(HERE, "scenes")


def main():
    os.makedirs(OUT, exist_ok=True)
    for name in ("cb", "cb_texture", "cbbunny", "cbbunny_area_light", "cbbunny_area_light_transforms", "checker", "test"):
        sc = rc.scene_from_gltf_file(os.path.join(REF, "scenes", name + ".glb"))
        sc.save_npz(os.path.join(OUT, name + ".npz"))
        print(name, sc.camera.raster_width, sc.camera.raster_height, sc.triangle_count(), "tris", len(sc.lights), "lights")
    with open(os.path.join(REF, "crates/raytracing/src/scene/test_scenes/assets/bunny.ply"), "rb") as fh:
        m = rc.mesh_from_ply_bytes(fh.read(), False)
    np.savez_compressed(os.path.join(OUT, "bunny_mesh.npz"), vertices=m.vertices, tris=m.tris, normals=m.normals)
    print("bunny", m.vertices.shape, m.tris.shape)


if __name__ == "__main__":
    main()
