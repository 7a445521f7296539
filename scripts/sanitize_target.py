"""compute-sanitizer target: small renders that touch every kernel (PLOC build, mip build, raygen / extend / shade (both
instantiations) / shadow / gather / resolve, AOV and pixel kernels, multi-batch path)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracing_cuda as rc
A = rc.AovFlags
for name, spp, cap in (("cbbunny_area_light_transforms", 2, 4000), ("cb_texture", 2, 0)):
    sc = rc.Scene.load_npz(os.path.join(ROOT, "tests/golden/scenes", name + ".npz"))
    sc.camera = sc.camera.with_raster_size(96, 54)
    st = rc.RaytracerSettings(samples_per_pixel=spp, outputs=A.BEAUTY | A.NORMALS | A.UV_COORDS | A.ALBEDO | A.MIP_LEVEL | A.DEBUG_IDS | A.DEBUG_DEPTH)
    with rc.CudaRenderer(sc, rc.CudaBackendSettings(max_paths_in_flight=cap, collect_stats=3)) as r:
        out = r.render(st)
        r.render_pixel(st, 40, 30, 0, 4)
        print(name, float(out.beauty.mean()), r.stats()["kernel_launches"])
t = [t for t in rc.test_scenes.all_test_scenes() if t.name == "rough_dielectric"][0]
sc = t.scene_func()
st = rc.RaytracerSettings(samples_per_pixel=1, light_sample_count=12, max_ray_depth=3)   # general shade kernel, two-pass NEE
print("rough_dielectric", float(rc.render(sc, st).beauty.mean()))
