"""One process, N GPUs: rc.render(scene, settings, CudaBackendSettings(num_devices=N)) with wall-clock per phase (diagnostic).
usage: e2e_probe_multi.py WORKLOAD N [calls]   (RTCUDA_TRACE=1 adds the library's own per-thread marks on stderr)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import raytracing_cuda as rc
wl, n = sys.argv[1], int(sys.argv[2])
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 3
sc, st = bench.load_workload(wl)
tile = 64 if wl == "C5" else 16
bs = rc.CudaBackendSettings(num_devices=n, device_ids=list(range(n)), tile_size=tile) if n > 1 else rc.CudaBackendSettings()
for it in range(calls):
    t0 = time.perf_counter()
    r = rc.CudaRenderer(sc, bs)
    t1 = time.perf_counter()
    out = r.render(st)
    t2 = time.perf_counter()
    stats = r.stats()
    r.close()
    t3 = time.perf_counter()
    print(f"call {it}: ctx+upload+build {1e3 * (t1 - t0):.1f} ms (setup {r.setup_ms}), render+frame {1e3 * (t2 - t1):.1f} ms (device {stats['render_ms']:.1f}), close {1e3 * (t3 - t2):.1f} ms, "
          f"total {1e3 * (t3 - t0):.1f} ms", flush=True)
