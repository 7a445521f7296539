"""First-light check on a B200: AOV + beauty parity of a few small scenes against the CPU oracle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import raytracing_cuda as rc
import oracle_py as orc
from raytracing_cuda import test_scenes as ts

def lum(img): return (0.2126 * img[..., 0] + 0.7152 * img[..., 1] + 0.0722 * img[..., 2])

def check(name, scene, settings, threads=8):
    t0 = time.time()
    r = rc.CudaRenderer(scene, rc.CudaBackendSettings(collect_stats=True))
    t1 = time.time()
    out = r.render(settings)
    t2 = time.time()
    st = r.stats()
    ref, ost = orc.render(scene, settings, num_threads=threads)
    t3 = time.time()
    print(f"== {name}: upload {t1-t0:.3f}s render {t2-t1:.3f}s (device {st['render_ms']:.2f} ms, build {st['bvh_build_ms']:.2f} ms) oracle {t3-t2:.3f}s")
    print("   gpu stats", {k: st[k] for k in ('samples','primary_rays','bounce_rays','shadow_rays','aov_rays','nodes_fetched','prims_fetched','kernel_launches','bvh_node_count','bvh_prim_count')})
    print("   oracle  ", ost)
    if out.debug_ids is not None:
        agree = (out.debug_ids == ref.debug_ids).all(axis=2).mean()
        print(f"   id agreement {agree*100:.4f}%")
    for plane in ("normals", "uv", "debug_depth", "albedo", "mip_level"):
        a, b = getattr(out, plane), getattr(ref, plane)
        if a is not None:
            d = np.abs(a - b)
            print(f"   {plane}: max abs diff {d.max():.3e}, frac > 1e-4: {(d > 1e-4).mean():.2e}")
    if out.beauty is not None:
        a, b = out.beauty, ref.beauty
        print(f"   beauty mean gpu {a.mean(axis=(0,1))} oracle {b.mean(axis=(0,1))} nan {np.isnan(a).sum()}/{np.isnan(b).sum()}")
        la, lb = lum(a), lum(b)
        spp = settings.samples_per_pixel
        sigma = np.sqrt((la.var() + lb.var()) / la.size)  # crude upper bound on the std of the mean difference
        print(f"   mean lum diff {la.mean()-lb.mean():.3e} (crude sigma {sigma:.3e}), mse {((a-b)**2).mean():.3e}")
    r.close()
    return out, ref

A = rc.AovFlags
dbg = A.NORMALS | A.UV_COORDS | A.DEBUG_IDS | A.DEBUG_DEPTH
check("sphere C1", ts.sphere_scene(), rc.RaytracerSettings(outputs=dbg, samples_per_pixel=4, max_ray_depth=5))
check("cube", ts.cube_scene(), rc.RaytracerSettings(outputs=dbg))
g = os.path.join(ROOT, "tests/golden/scenes")
cb = rc.Scene.load_npz(os.path.join(g, "cb.npz")); cb.camera = cb.camera.with_raster_size(256, 256)
check("cb 256", cb, rc.RaytracerSettings(outputs=dbg | A.BEAUTY, samples_per_pixel=16, light_sample_count=1))
bun = rc.Scene.load_npz(os.path.join(g, "cbbunny_area_light_transforms.npz")); bun.camera = bun.camera.with_raster_size(320, 180)
check("bunny 320x180", bun, rc.RaytracerSettings(outputs=dbg | A.BEAUTY, samples_per_pixel=8))
tex = rc.Scene.load_npz(os.path.join(g, "cb_texture.npz")); tex.camera = tex.camera.with_raster_size(320, 180)
check("cb_texture 320x180", tex, rc.RaytracerSettings(outputs=dbg | A.BEAUTY | A.ALBEDO | A.MIP_LEVEL, samples_per_pixel=8))
for t in ts.all_test_scenes()[3:]:
    sc = t.scene_func(); st = t.settings_func(); st.samples_per_pixel = min(st.samples_per_pixel, 8) if st.sampler.kind != "stratified" else st.samples_per_pixel
    st.outputs = dbg | A.BEAUTY
    check(t.name, sc, st)
