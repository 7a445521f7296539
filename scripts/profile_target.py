"""ncu target: BASELINE config C3 (cbbunny_area_light_transforms, 1920x1080, depth 8, light samples 4) at 16 spp —
two renders of one 33 Mi-path batch each; per batch the launch order is raygen, then per depth
extend, shade, shadow (26 traversal/shade launches)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracing_cuda as rc
name = sys.argv[1] if len(sys.argv) > 1 else "cbbunny_area_light_transforms"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
sc = rc.Scene.load_npz(os.path.join(ROOT, "tests/golden/scenes", name + ".npz"))
sc.camera = sc.camera.with_raster_size(1920, 1080)
st = rc.RaytracerSettings(samples_per_pixel=spp)
with rc.CudaRenderer(sc) as r:
    for _ in range(2):
        out = r.render(st)
        print(r.stats()["render_ms"], float(out.beauty.mean()))
