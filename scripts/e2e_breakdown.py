"""Where the end-to-end time of raytracing_cuda.render(scene, settings) goes (C3): context, upload + BVH build, render, download."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracing_cuda as rc
sc = rc.Scene.load_npz(os.path.join(ROOT, "tests/golden/scenes/cbbunny_area_light_transforms.npz"))
sc.camera = sc.camera.with_raster_size(1920, 1080)
st = rc.RaytracerSettings(samples_per_pixel=int(sys.argv[1]) if len(sys.argv) > 1 else 256)
for it in range(3):
    t0 = time.time()
    r = rc.CudaRenderer(sc)
    t1 = time.time()
    out = r.render(st)
    t2 = time.time()
    stats = r.stats()
    r.close()
    t3 = time.time()
    print(f"iter {it}: ctx+upload+build {1e3 * (t1 - t0):.1f} ms, render call {1e3 * (t2 - t1):.1f} ms (device render {stats['render_ms']:.1f} ms, "
          f"bvh build {stats.get('bvh_build_ms', -1):.1f} ms), close {1e3 * (t3 - t2):.1f} ms")
