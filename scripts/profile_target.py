"""ncu target: one BASELINE config at a reduced sample count — two renders of one wavefront batch each; per batch the launch
order is raygen, then per depth extend, shade, shadow, shadow_gather (35 traversal / shade launches at depth 8).

  python scripts/profile_target.py [fixture | C5] [spp]     default: C3's scene (cbbunny_area_light_transforms) at 16 spp
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracing_cuda as rc
name = sys.argv[1] if len(sys.argv) > 1 else "cbbunny_area_light_transforms"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ls = 4
if name == "C5":   # the 16.8 M-triangle procedural mesh inside the Cornell box (bench.py --workload C5), 1080p
    sc = rc.Scene.load_npz(os.path.join(ROOT, "tests/golden/scenes/cb.npz"))
    sc = rc.test_scenes.synthetic_mesh_scene(sc, 4096, 2048)
    ls = 1
else:
    sc = rc.Scene.load_npz(os.path.join(ROOT, "tests/golden/scenes", name + ".npz"))
sc.camera = sc.camera.with_raster_size(1920, 1080)
st = rc.RaytracerSettings(samples_per_pixel=spp, light_sample_count=ls)
with rc.CudaRenderer(sc) as r:
    for _ in range(2):
        out = r.render(st)
        print(r.stats()["render_ms"], float(out.beauty.mean()))
