"""ctypes binding of tests/hostsim/libhostsim.so — TEST INFRASTRUCTURE ONLY (see hostsim.cpp): the
per-thread kernel bodies of the CUDA library compiled for the CPU and run sequentially."""
import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import raytracing_cuda as rc  # noqa: E402
from raytracing_cuda import _ffi  # noqa: E402

LIB_PATH = os.path.join(HERE, "libhostsim.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(HERE, "hostsim.cpp")] + [os.path.join(ROOT, "opencl-raytracing_b200", "csrc", f)
                                                   for f in os.listdir(os.path.join(ROOT, "opencl-raytracing_b200", "csrc")) if f.endswith(".h")]
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", LIB_PATH, srcs[0]])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(LIB_PATH)
        l.hostsim_render.argtypes = [C.POINTER(_ffi.SceneDesc), C.POINTER(_ffi.Settings), C.POINTER(_ffi.Outputs), C.c_uint32,
                                     C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]
        l.hostsim_render.restype = C.c_int
        l.hostsim_render_samples.argtypes = l.hostsim_render.argtypes + [C.c_uint32, C.c_uint32]
        l.hostsim_render_samples.restype = C.c_int
        fp = C.POINTER(C.c_float)
        l.hostsim_ray_triangle.argtypes = [C.c_int, fp, fp, fp, fp, fp, C.c_float, C.c_float, fp]
        l.hostsim_ray_triangle.restype = C.c_int
        l.hostsim_raster_rect.argtypes = [C.POINTER(_ffi.SceneDesc), C.POINTER(C.c_int)]
        l.hostsim_raster_rect.restype = C.c_int
        _lib = l
    return _lib


def render(scene, settings, tile_rank=0, tile_world=1, capacity=0, sample_range=None):
    """sample_range=(lo, hi): only those samples, the beauty plane holds their un-normalised sum (rtcuda_render_samples_device)"""
    holder = scene.to_desc()
    out = rc.RenderOutput.allocate(scene.camera.raster_width, scene.camera.raster_height, rc.AovFlags(settings.outputs))
    s, o = settings.to_c(), out.to_c()
    stats = (C.c_uint64 * 8)()
    lo, hi = sample_range if sample_range is not None else (0, 0)
    if lib().hostsim_render_samples(C.byref(holder.desc), C.byref(s), C.byref(o), tile_rank, tile_world, capacity, stats, lo, hi) != 0:
        raise RuntimeError("hostsim_render failed")
    names = ["primary_rays", "bounce_rays", "shadow_rays", "aov_rays", "nodes_fetched", "prims_fetched", "bvh_node_count", "collapse_levels"]
    return out, dict(zip(names, list(stats)))


def ray_triangle(mode, p0, p1, p2, o, d, t_min=0.0, t_max=float("inf")):
    """(hit, t, u, v) of the kernel-side triangle tests: mode 0 = Moller-Trumbore (reference), 1 = watertight"""
    fa = lambda v: (C.c_float * 3)(*[float(x) for x in v])
    out = (C.c_float * 3)()
    h = lib().hostsim_ray_triangle(mode, fa(p0), fa(p1), fa(p2), fa(o), fa(d), t_min, t_max, out)
    return bool(h), out[0], out[1], out[2]


def raster_rect(scene):
    """rt_cull.h scene_raster_rect: (x0, y0, x1, y1) inclusive, or None when every pixel is kept"""
    holder = scene.to_desc()
    rect = (C.c_int * 4)()
    return tuple(rect) if lib().hostsim_raster_rect(C.byref(holder.desc), rect) else None
