"""Reproduce bench.py's e2e sequence with per-call timing of init / to_desc / upload / render / release (diagnostic)."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import raytracing_cuda as rc
from raytracing_cuda import _ffi
torch.cuda.set_device(0)
sc = rc.Scene.load_npz(os.path.join(ROOT, "tests/golden/scenes/cbbunny_area_light_transforms.npz"))
sc.camera = sc.camera.with_raster_size(1920, 1080)
st = rc.RaytracerSettings(samples_per_pixel=256)
use_dr = len(sys.argv) > 1 and sys.argv[1] == "dr"
if use_dr:
    dr = rc.multi_gpu.DistributedRenderer(sc, 0, 1, device_id=0, collect_stats=_ffi.STATS_KERNEL_TIMES)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda:0")
    for _ in range(2):
        flush.zero_(); torch.cuda.synchronize()
        dr.render_local(st)
    print("dr render ms", dr.renderer.stats()["render_ms"])
lib = _ffi.load_library()
for it in range(4):
    t = [time.time()]
    ctx, scn = C.c_void_p(), C.c_void_p()
    bs = rc.CudaBackendSettings().to_c()
    lib.rtcuda_init(C.byref(bs), C.byref(ctx)); t.append(time.time())
    holder = sc.to_desc(); t.append(time.time())
    lib.rtcuda_scene_upload(ctx, C.byref(holder.desc), C.byref(scn)); t.append(time.time())
    out = rc.RenderOutput.allocate(1920, 1080, rc.AovFlags.BEAUTY); t.append(time.time())
    s, o = st.to_c(), out.to_c()
    lib.rtcuda_render(scn, C.byref(s), C.byref(o)); t.append(time.time())
    stt = _ffi.Stats(); lib.rtcuda_get_stats(scn, C.byref(stt))
    lib.rtcuda_scene_release(scn); t.append(time.time())
    lib.rtcuda_shutdown(ctx); t.append(time.time())
    names = ["init", "to_desc", "upload", "alloc_out", "render", "release", "shutdown"]
    print(it, {n: round(1e3 * (t[i + 1] - t[i]), 1) for i, n in enumerate(names)}, "device", round(stt.render_ms, 1), "upload_ms", round(stt.upload_ms, 1), "bvh", round(stt.bvh_build_ms, 1))
