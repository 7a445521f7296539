"""ncu target: one bench workload (bench.py WORKLOADS) at a reduced sample count — two renders of one wavefront batch each.
  python scripts/profile_workload.py CM [spp]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import raytracing_cuda as rc
wl = sys.argv[1] if len(sys.argv) > 1 else "CM"
sc, st = bench.load_workload(wl)
if len(sys.argv) > 2:
    st.samples_per_pixel = int(sys.argv[2])
with rc.CudaRenderer(sc) as r:
    for _ in range(2):
        out = r.render(st)
        print(r.stats()["render_ms"], float(out.beauty.mean()))
