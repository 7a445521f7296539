"""`-m gpu`: parity of libraytracing_cuda.so (through the C ABI, include/rtcuda.h) against the CPU oracle on
a B200 — BASELINE.json configs at oracle-sized resolutions, every material / camera / sampler the reference
tests (tests/tests.toml = the builtin scenes), plus size-independent properties at the full BASELINE sizes."""
import numpy as np
import pytest

from conftest import load_scene, bunny_mesh
from parity import assert_first_hit_parity, assert_beauty_parity, SPECULAR_GATES, STEEP_TEXTURE_GATES, luminance, beauty_close, mean_luminance_z

pytestmark = pytest.mark.gpu
A = None
import os
NT = os.cpu_count() or 8   # oracle threads


@pytest.fixture(autouse=True)
def _flags(rc):
    global A
    A = rc.AovFlags


def dbg():
    return A.NORMALS | A.UV_COORDS | A.ALBEDO | A.MIP_LEVEL | A.DEBUG_IDS | A.DEBUG_DEPTH


def gpu_render(rc, scene, settings, **backend):
    with rc.CudaRenderer(scene, rc.CudaBackendSettings(**backend)) as r:
        out = r.render(settings)
        stats = r.stats()
        if out.beauty is not None:   # the device side of the reference's NaN / Inf scan (lib.rs:813-854)
            assert stats["nonfinite_values"] == int((~np.isfinite(out.beauty)).sum())
        return out, stats


def test_extension_is_loaded(rc):
    import ctypes
    lib = rc._ffi.load_library()
    assert isinstance(lib, ctypes.CDLL) and lib.rtcuda_abi_version() == rc._ffi.ABI_VERSION


def test_c1_sphere(rc, oracle):
    """BASELINE config C1: builtin `sphere`, normals only, 4 spp, depth 5"""
    sc = rc.test_scenes.sphere_scene()
    st = rc.test_scenes.all_test_scenes()[0].settings_func()
    st.samples_per_pixel, st.max_ray_depth = 4, 5
    st.outputs = dbg()
    out, stats = gpu_render(rc, sc, st)
    ref, _ = oracle.render(sc, st)
    assert stats["aov_rays"] == 160000 and stats["kernel_launches"] == 1
    assert_first_hit_parity(out, ref)
    st.outputs = A.NORMALS
    only, _ = gpu_render(rc, sc, st)
    assert only.beauty is None and np.array_equal(only.normals, out.normals)


@pytest.mark.parametrize("name", ["cube", "cube_orthographic"])
def test_builtin_normals(rc, oracle, name):
    t = [t for t in rc.test_scenes.all_test_scenes() if t.name == name][0]
    sc, st = t.scene_func(), t.settings_func()
    st.outputs = dbg()
    out, _ = gpu_render(rc, sc, st)
    ref, _ = oracle.render(sc, st)
    assert_first_hit_parity(out, ref)


def test_c2_cornell_box_512(rc, oracle):
    """BASELINE config C2: scenes/cb.glb 512x512, NEE light-samples 1 (spp reduced to what the oracle renders in
    seconds; the full 64 spp is exercised by the different-seed test below)"""
    sc = load_scene("cb", 512, 512)
    st = rc.RaytracerSettings(outputs=dbg() | A.BEAUTY, samples_per_pixel=8, light_sample_count=1)
    out, stats = gpu_render(rc, sc, st, collect_stats=True)
    ref, ostats = oracle.render(sc, st, num_threads=8)
    assert_first_hit_parity(out, ref)
    assert stats["primary_rays"] == ostats["primary_rays"] == 512 * 512 * 8
    # the oracle (like the reference) also traces the last-depth rays that cannot add radiance; the backend reports them apart
    assert stats["final_rays_skipped"] > 0
    assert abs(stats["bounce_rays"] + stats["final_rays_skipped"] - ostats["bounce_rays"]) <= ostats["bounce_rays"] // 2000
    assert_beauty_parity(out.beauty, ref.beauty)
    assert abs(luminance(out.beauty).mean() - luminance(ref.beauty).mean()) <= 1e-3 * luminance(ref.beauty).mean()
    assert stats["nodes_fetched"] > 0 and stats["prims_fetched"] > 0


def test_c2_independent_seeds_are_statistically_indistinguishable(rc, oracle):
    """north_star: at equal spp the beauty mean luminance is within 3 sigma of the reference — checked with
    DIFFERENT seeds so the two estimates are independent"""
    sc = load_scene("cb", 256, 256)
    gpu, _ = gpu_render(rc, sc, rc.RaytracerSettings(samples_per_pixel=64, light_sample_count=1, seed=1234))
    ref, _ = oracle.render(sc, rc.RaytracerSettings(samples_per_pixel=64, light_sample_count=1, seed=99), num_threads=8)
    z = mean_luminance_z(gpu.beauty, ref.beauty, 64, 64)
    assert abs(z) < 3.0, z


@pytest.mark.parametrize("name,w,h,spp", [("cbbunny_area_light_transforms", 480, 270, 4), ("cbbunny_area_light", 480, 270, 4),
                                          ("cb_texture", 480, 270, 4), ("checker", 320, 180, 2), ("cbbunny", 320, 180, 2),
                                          ("test", 320, 240, 2)])
def test_gltf_scenes(rc, oracle, name, w, h, spp):
    """BASELINE configs C3 / C4 (instanced transforms + area light; texture fetch + AOVs) at reduced raster"""
    sc = load_scene(name, w, h)
    st = rc.RaytracerSettings(outputs=dbg() | A.BEAUTY, samples_per_pixel=spp)
    out, stats = gpu_render(rc, sc, st, max_paths_in_flight=100000)   # several pixel / sample batches
    ref, ostats = oracle.render(sc, st, num_threads=8)
    assert_first_hit_parity(out, ref)
    assert stats["primary_rays"] == ostats["primary_rays"]
    assert abs(stats["bounce_rays"] + stats["final_rays_skipped"] - ostats["bounce_rays"]) <= max(8, ostats["bounce_rays"] // 1000)
    assert_beauty_parity(out.beauty, ref.beauty, what=name, **(STEEP_TEXTURE_GATES if name == "checker" else {}))
    la, lb = luminance(out.beauty), luminance(ref.beauty)
    assert abs(la.mean() - lb.mean()) <= 2e-3 * lb.mean()


def test_c3_primary_hits_full_raster(rc, oracle):
    """BASELINE config C3 at the full 1920x1080 raster: primary-hit ids >= 99.99 %, normal / uv / depth 1e-4"""
    sc = load_scene("cbbunny_area_light_transforms", 1920, 1080)
    st = rc.RaytracerSettings(outputs=dbg())
    out, _ = gpu_render(rc, sc, st)
    ref, _ = oracle.render(sc, st)
    assert_first_hit_parity(out, ref)


def test_c4_texture_aovs_full_raster(rc, oracle):
    """BASELINE config C4: cb_texture.glb 1080p with normal,uv AOVs (+ albedo / mip level through the trilinear path)"""
    sc = load_scene("cb_texture", 1920, 1080)
    st = rc.RaytracerSettings(outputs=dbg(), samples_per_pixel=128)
    out, _ = gpu_render(rc, sc, st)
    ref, _ = oracle.render(sc, st)
    assert_first_hit_parity(out, ref)


@pytest.mark.parametrize("name", ["checkered_plane", "dielectric", "metal", "rough_metal", "rough_dielectric", "out_of_focus_sphere"])
def test_builtin_materials_and_cameras(rc, oracle, name):
    t = [t for t in rc.test_scenes.all_test_scenes() if t.name == name][0]
    sc, st = t.scene_func(), t.settings_func()
    if st.sampler.kind != "stratified":
        st.samples_per_pixel = min(st.samples_per_pixel, 16)   # the tail of the specular scenes is gated statistically: keep the noise floor low
    st.outputs = dbg() | A.BEAUTY
    out, _ = gpu_render(rc, sc, st)
    ref, _ = oracle.render(sc, st, num_threads=NT)
    assert_first_hit_parity(out, ref)
    la, lb = luminance(out.beauty), luminance(ref.beauty)
    assert np.isnan(la).sum() == np.isnan(lb).sum()
    assert abs(np.nanmean(la) - np.nanmean(lb)) <= 2e-3 * abs(np.nanmean(lb)) + 1e-7
    assert_beauty_parity(out.beauty, ref.beauty, **SPECULAR_GATES)


def test_coated_diffuse_bunny(rc, oracle):
    sc = rc.test_scenes.coated_diffuse_bunny_scene(bunny=bunny_mesh())
    import math
    sc.camera = rc.Camera.lookat_camera_perspective((0.0, 4.4, 0.4), (0, 0, 0.75), (0, 0, 1), False,
                                                    float(np.float32(37.8) * np.float32(math.pi / 180)), 128, 128)
    st = rc.RaytracerSettings(outputs=dbg() | A.BEAUTY, samples_per_pixel=2, max_ray_depth=3, light_sample_count=1)
    out, _ = gpu_render(rc, sc, st)
    ref, _ = oracle.render(sc, st, num_threads=8)
    assert_first_hit_parity(out, ref)
    la, lb = luminance(out.beauty), luminance(ref.beauty)
    assert abs(np.nanmean(la) - np.nanmean(lb)) <= 1e-2 * abs(np.nanmean(lb))


def test_environment_light(rc, oracle):
    sc = rc.test_scenes.environment_lighting_scene(rc.test_scenes.synthetic_environment_map())
    sc.camera = rc.Camera.lookat_camera_perspective((0.013, 0, 0.007), (0.1, 1, 0.05), (0, 0, 1), False, 0.66, 200, 200)
    st = rc.RaytracerSettings(outputs=dbg() | A.BEAUTY, samples_per_pixel=4)
    out, _ = gpu_render(rc, sc, st)
    ref, _ = oracle.render(sc, st, num_threads=8)
    assert_first_hit_parity(out, ref)
    assert np.abs(out.beauty - ref.beauty).max() <= 1e-3


def test_stratified_sampler(rc, oracle):
    sc = load_scene("cb", 128, 128)
    for jitter in (True, False):
        st = rc.RaytracerSettings(outputs=A.BEAUTY, samples_per_pixel=16, light_sample_count=2, sampler=rc.Sampler.stratified(jitter, 4, 4))
        out, _ = gpu_render(rc, sc, st)
        ref, _ = oracle.render(sc, st, num_threads=8)
        assert_beauty_parity(out.beauty, ref.beauty, what=str(jitter))


def test_settings_variants(rc, oracle):
    sc = load_scene("cb", 128, 128)
    for kw in (dict(max_ray_depth=0), dict(max_ray_depth=1), dict(max_ray_depth=3, accumulate_bounces=False), dict(seed=7),
               dict(antialias_primary_rays=False), dict(light_sample_count=3)):
        st = rc.RaytracerSettings(outputs=A.BEAUTY, samples_per_pixel=4, **kw)
        out, _ = gpu_render(rc, sc, st)
        ref, _ = oracle.render(sc, st, num_threads=8)
        assert_beauty_parity(out.beauty, ref.beauty, what=str(kw))
        assert abs(out.beauty.mean() - ref.beauty.mean()) <= 2e-3 * ref.beauty.mean() + 1e-8, kw


def test_pixel_diagnostics(rc, oracle):
    """`pixel X Y count offset` (cli main.rs:202-232) through rtcuda_render_pixel vs the oracle's render_single_pixel"""
    sc = load_scene("cbbunny_area_light_transforms", 320, 180)
    st = rc.RaytracerSettings(samples_per_pixel=16)
    with rc.CudaRenderer(sc) as r:
        for (x, y) in ((160, 90), (10, 170), (400, 500)):   # the last one is clamped (lib.rs:867-876)
            got = r.render_pixel(st, x, y, 3, 11)
            want = oracle.render_pixel(sc, st, x, y, 3, 11)
            assert [g.sample_index for g in got] == list(range(3, 11))
            for g, w in zip(got, want):
                assert g.hit == w.hit
                assert np.allclose(g.uv, w.uv, atol=1e-4) and np.allclose(g.normal, w.normal, atol=1e-4)
            gr, wr = np.array([g.radiance for g in got]), np.array([w.radiance for w in want])
            assert np.allclose(gr.mean(axis=0), wr.mean(axis=0), rtol=2e-2, atol=1e-5)
        assert r.render_pixel(st, 5, 5, 4, 4) == []
    single = rc.render_single_pixel(sc, st, 160, 90, 5)
    assert single.sample_index == 5 and single.hit


def test_deterministic(rc):
    """the reference renderer is deterministic (visual-testing/README.md:101-103): so is the wavefront"""
    sc = load_scene("cbbunny_area_light_transforms", 320, 180)
    st = rc.RaytracerSettings(outputs=A.BEAUTY | A.NORMALS, samples_per_pixel=8)
    with rc.CudaRenderer(sc) as r:
        a = r.render(st)
        b = r.render(st)
    c, _ = gpu_render(rc, sc, st, max_paths_in_flight=50000)
    assert np.array_equal(a.beauty, b.beauty) and np.array_equal(a.normals, b.normals)
    assert np.array_equal(a.beauty, c.beauty)   # batch shape does not change any pixel


def test_tile_partition_is_exact(rc):
    """multi-GPU partition: tile-disjoint frames of 3 'ranks' sum bit-exactly to the full frame"""
    sc = load_scene("cb", 200, 136)
    st = rc.RaytracerSettings(outputs=A.BEAUTY | A.NORMALS | A.UV_COORDS, samples_per_pixel=4, light_sample_count=1)
    full, _ = gpu_render(rc, sc, st)
    parts = [gpu_render(rc, sc, st, tile_rank=r, tile_world=3)[0] for r in range(3)]
    owner = rc.multi_gpu.tile_owner_map(200, 136, 3)
    for r, p in enumerate(parts):
        assert (p.beauty[owner != r] == 0).all()
    for plane in ("beauty", "normals", "uv"):
        assert np.array_equal(sum(getattr(p, plane) for p in parts), getattr(full, plane))


def test_tile_size_partition_is_exact(rc):
    """finer deal of the frame (backend_settings.tile_size = 16): same pixels, still a bit-exact gather"""
    sc = load_scene("cb", 200, 136)
    st = rc.RaytracerSettings(outputs=A.BEAUTY | A.NORMALS, samples_per_pixel=4, light_sample_count=1)
    full, _ = gpu_render(rc, sc, st)
    whole16, _ = gpu_render(rc, sc, st, tile_size=16)
    assert np.array_equal(whole16.beauty, full.beauty)          # the tile grid only orders the work
    parts = [gpu_render(rc, sc, st, tile_rank=r, tile_world=4, tile_size=16)[0] for r in range(4)]
    owner = rc.multi_gpu.tile_owner_map(200, 136, 4, tile=16)
    for r, p in enumerate(parts):
        assert (p.beauty[owner != r] == 0).all() and (p.beauty[owner == r] != 0).any()
    for plane in ("beauty", "normals"):
        assert np.array_equal(sum(getattr(p, plane) for p in parts), getattr(full, plane))
    with pytest.raises(rc._ffi.RtCudaError):
        gpu_render(rc, sc, st, tile_size=24)


def test_primary_rays_that_miss_the_scene_bounds_are_not_queued(rc, oracle):
    """raygen drops camera rays that miss the (grown) scene bounds when there is no environment light — the reference's
    root-AABB reject (accel.rs:95): same pixels as the oracle, every sample still counted as a primary ray"""
    sc = load_scene("cbbunny_area_light_transforms", 320, 180)      # the box covers about a quarter of the 16:9 raster
    st = rc.RaytracerSettings(outputs=A.BEAUTY | A.DEBUG_IDS, samples_per_pixel=4)
    out, stats = gpu_render(rc, sc, st)
    ref, ostats = oracle.render(sc, st, num_threads=8)
    assert stats["primary_rays"] == ostats["primary_rays"] == 320 * 180 * 4
    assert 0.5 * stats["primary_rays"] < stats["primary_rays_culled"] < 0.8 * stats["primary_rays"]
    assert_beauty_parity(out.beauty, ref.beauty)
    assert (out.beauty[(ref.beauty == 0).all(axis=-1)] == 0).all()          # culled pixels are exactly black, like the oracle's misses
    env = rc.test_scenes.environment_lighting_scene(rc.test_scenes.synthetic_environment_map())
    env.camera = rc.Camera.lookat_camera_perspective((0.013, 0, 0.007), (0.1, 1, 0.05), (0, 0, 1), False, 0.66, 64, 64)
    _, estats = gpu_render(rc, env, rc.RaytracerSettings(samples_per_pixel=2))
    assert estats["primary_rays_culled"] == 0                                # misses light the pixel through the environment map


def test_material_split_matches_the_single_launch(rc, monkeypatch):
    """mixed materials: Diffuse kernel + general kernel over the deferred vertices (launch_shade) against everything through the
    general kernel (RTCUDA_NO_MATERIAL_SPLIT): same rays, the same frame up to the rounding of two instantiations"""
    t = [t for t in rc.test_scenes.all_test_scenes() if t.name == "rough_metal"][0]
    sc, st = t.scene_func(), t.settings_func()
    st.samples_per_pixel, st.outputs = 16, A.BEAUTY
    split, s1 = gpu_render(rc, sc, st)
    monkeypatch.setenv("RTCUDA_NO_MATERIAL_SPLIT", "1")
    single, s2 = gpu_render(rc, sc, st)
    assert s1["kernel_launches"] > s2["kernel_launches"]
    assert abs(s1["bounce_rays"] - s2["bounce_rays"]) <= max(4, s2["bounce_rays"] // 2000)
    rep = assert_beauty_parity(split.beauty, single.beauty, what="material split", **SPECULAR_GATES)
    print(f"\n[parity] material split vs single launch: {rep}")


def test_reused_mesh_instances(rc, oracle):
    """one mesh under three Transform primitives (glTF mesh reuse, scene/scene.rs:430-443): the flattened world-space tree gives
    the reference's ids / normals / uv / depth and beauty"""
    from conftest import instanced_bunnies_scene
    sc = instanced_bunnies_scene(320, 240)
    st = rc.RaytracerSettings(outputs=dbg() | A.BEAUTY, samples_per_pixel=8, max_ray_depth=6, light_sample_count=2)
    out, stats = gpu_render(rc, sc, st)
    ref, _ = oracle.render(sc, st)
    assert_first_hit_parity(out, ref)
    rep = assert_beauty_parity(out.beauty, ref.beauty, what="instanced bunnies")
    print(f"\n[parity] reused mesh x3: {rep}")


def test_page_locked_frame_planes_equal_pageable_ones(rc):
    """rtcuda_host_alloc planes (the copy engine writes the frame itself) vs planes from anywhere else (pinned staging + host copy):
    same bits, one and several devices; the buffers go back to the library's cache when the arrays are collected"""
    import gc
    sc = load_scene("cb_texture", 200, 120)
    st = rc.RaytracerSettings(outputs=A.BEAUTY | A.NORMALS | A.UV_COORDS | A.DEBUG_IDS, samples_per_pixel=2)
    lib = rc._ffi.load_library()
    ids = _device_ids(2)
    for backend in ({}, {"num_devices": len(ids), "device_ids": ids, "tile_size": 16}):
        with rc.CudaRenderer(sc, rc.CudaBackendSettings(**backend)) as r:
            pinned = r.render(st)                                           # planes from rtcuda_host_alloc
            plain = rc.RenderOutput.allocate(200, 120, rc.AovFlags(st.outputs))   # numpy's own memory
            r.render_into(st, plain)
            for plane in ("beauty", "normals", "uv", "debug_ids"):
                assert np.array_equal(getattr(pinned, plane), getattr(plain, plane)), plane
    ptrs = {getattr(pinned, plane).ctypes.data for plane in ("beauty", "normals", "uv", "debug_ids")}
    del pinned
    gc.collect()
    held = [rc._ffi.host_array(lib, (120, 200, 3), np.float32) for _ in range(64)]   # parked buffers are handed out again
    assert all(h is not None for h in held) and any(h.ctypes.data in ptrs for h in held)


def test_own_arrays_and_scene_wide_arrays_upload_the_same_scene(rc):
    """rtcuda_shape.vertices / tris / normals / uvs (zero-copy from the caller's meshes) vs the concatenated scene-wide arrays"""
    import ctypes as C
    sc = load_scene("cb_texture", 160, 90)
    st = rc.RaytracerSettings(outputs=A.BEAUTY | A.NORMALS | A.UV_COORDS | A.DEBUG_IDS, samples_per_pixel=2)
    a, _ = gpu_render(rc, sc, st)                       # CudaRenderer uploads with own_arrays=True
    lib = rc._ffi.load_library()
    ctx, scn = C.c_void_p(), C.c_void_p()
    bs = rc.CudaBackendSettings().to_c()
    rc._ffi.check(lib, lib.rtcuda_init(C.byref(bs), C.byref(ctx)), "init")
    holder = sc.to_desc(own_arrays=False)
    rc._ffi.check(lib, lib.rtcuda_scene_upload(ctx, C.byref(holder.desc), C.byref(scn)), "upload")
    b = rc.RenderOutput.allocate(160, 90, A(st.outputs))
    s, o = st.to_c(), b.to_c()
    rc._ffi.check(lib, lib.rtcuda_render(scn, C.byref(s), C.byref(o)), "render")
    lib.rtcuda_scene_release(scn)
    lib.rtcuda_shutdown(ctx)
    for plane in ("beauty", "normals", "uv", "debug_ids"):
        assert np.array_equal(getattr(a, plane), getattr(b, plane)), plane


def test_sample_range_render(rc):
    """rtcuda_render_samples_device: the sums of disjoint sample ranges add up to the frame (same streams per sample
    index; only the association of the float sum differs), and the full range reproduces it bit for bit"""
    import torch
    sc = load_scene("cb", 96, 64)
    st = rc.RaytracerSettings(samples_per_pixel=12, light_sample_count=2)
    with rc.CudaRenderer(sc) as r:
        full = r.render(st).beauty
        plane = torch.zeros((64, 96, 3), dtype=torch.float32, device="cuda:0")
        r.render_samples_device(st, 0, 12, plane.data_ptr())
        torch.cuda.synchronize()
        inv = torch.tensor(1.0 / 12, dtype=torch.float32)
        assert np.array_equal((plane.cpu() * inv).numpy(), full)
        acc = torch.zeros_like(plane)
        for lo, hi in ((0, 5), (5, 6), (6, 12)):
            r.render_samples_device(st, lo, hi, plane.data_ptr())
            torch.cuda.synchronize()
            acc += plane
        np.testing.assert_allclose((acc.cpu() * inv).numpy(), full, rtol=2e-6, atol=1e-7)
        assert r.stats()["samples"] == 96 * 64 * 6
        with pytest.raises(rc._ffi.RtCudaError):
            r.render_samples_device(st, 4, 13, plane.data_ptr())
    # frames large enough for the captured frame graph: the mean and the un-normalised sum of the same range into the SAME plane
    # must not share a graph (the finalize scale is baked into it)
    sc = load_scene("cb", 512, 512)
    st = rc.RaytracerSettings(samples_per_pixel=16, light_sample_count=1)
    with rc.CudaRenderer(sc) as r:
        plane = torch.zeros((512, 512, 3), dtype=torch.float32, device="cuda:0")
        r.render_device(st, {"beauty": plane.data_ptr()})
        torch.cuda.synchronize()
        mean = plane.clone()
        r.render_samples_device(st, 0, 16, plane.data_ptr())
        torch.cuda.synchronize()
        assert torch.equal(plane * torch.tensor(1.0 / 16, dtype=torch.float32), mean)
        r.render_device(st, {"beauty": plane.data_ptr()})
        torch.cuda.synchronize()
        assert torch.equal(plane, mean)


def test_full_size_c2_properties(rc):
    """BASELINE config C2 at full size (512x512, 64 spp): energy bound, symmetry-free sanity, convergence towards
    the low-spp estimate, no NaN"""
    sc = load_scene("cb", 512, 512)
    hi, stats = gpu_render(rc, sc, rc.RaytracerSettings(samples_per_pixel=64, light_sample_count=1), collect_stats=True)
    lo, _ = gpu_render(rc, sc, rc.RaytracerSettings(samples_per_pixel=16, light_sample_count=1, seed=5))
    assert stats["samples"] == 512 * 512 * 64 and not np.isnan(hi.beauty).any()
    assert hi.beauty.min() >= 0 and hi.beauty.max() <= 1.0 + 1e-4          # emitter radiance 1, albedo < 1
    assert abs(mean_luminance_z(hi.beauty, lo.beauty, 64, 16)) < 3.5
    rays = stats["primary_rays"] + stats["bounce_rays"] + stats["shadow_rays"]
    assert stats["primary_rays"] <= rays <= stats["samples"] * 17


def test_synthetic_mesh_large(rc, oracle):
    """BASELINE config C5 (scaled down to 2 x 256 x 128 triangles for the oracle; bench.py runs larger): LBVH
    over a dense displaced sphere, parity of primary hits"""
    base = load_scene("cb", 256, 256)
    sc = rc.test_scenes.synthetic_mesh_scene(base, 256, 128)
    st = rc.RaytracerSettings(outputs=dbg() | A.BEAUTY, samples_per_pixel=2, light_sample_count=1)
    out, stats = gpu_render(rc, sc, st)
    ref, _ = oracle.render(sc, st, num_threads=8)
    assert stats["bvh_prim_count"] == 12 + 2 * 256 * 128 - 2 * 256
    assert_first_hit_parity(out, ref)
    assert_beauty_parity(out.beauty, ref.beauty)


def test_error_behaviour(rc):
    import ctypes as C
    lib = rc._ffi.load_library()
    ctx = C.c_void_p()
    bs = rc.CudaBackendSettings(device_id=99).to_c()
    assert lib.rtcuda_init(C.byref(bs), C.byref(ctx)) == 1 and b"device_id" in lib.rtcuda_last_error()
    sc = rc.test_scenes.sphere_scene()
    with rc.CudaRenderer(sc) as r:
        out = rc.RenderOutput.allocate(100, 100, A.NORMALS)
        with pytest.raises(rc._ffi.RtCudaError, match="INVALID_ARGUMENT"):
            r.render_into(rc.RaytracerSettings(outputs=A.NORMALS), out)
        with pytest.raises(rc._ffi.RtCudaError, match="INVALID_ARGUMENT"):
            r.render(rc.RaytracerSettings(samples_per_pixel=0))
    bad = rc.test_scenes.sphere_scene()
    bad.shapes[0].material = 7
    with pytest.raises(rc._ffi.RtCudaError, match="out of range"):
        rc.CudaRenderer(bad)
    import copy
    broken = copy.deepcopy(load_scene("cb", 64, 64))
    broken.shapes[1].shape.tris[0, 2] = 1000           # checked on the device after the upload
    with pytest.raises(rc._ffi.RtCudaError, match="triangle index out of range"):
        rc.CudaRenderer(broken)
    empty = rc.SceneBuilder()
    empty.add_camera(rc.Camera.lookat_camera_perspective((0, 0, 0), (0, 0, -1), (0, 1, 0), False, 0.7, 32, 32))
    out, _ = gpu_render(rc, empty.build(), rc.RaytracerSettings(outputs=A.BEAUTY | A.NORMALS, samples_per_pixel=2))
    assert (out.beauty == 0).all() and (out.normals == 0).all()


def test_many_light_samples_two_pass_nee(rc, oracle):
    """K > NEE_STAGE light samples per vertex: the count pass + write pass of k_shade (rt_integrator.h nee_pass)"""
    sc = load_scene("cb", 96, 96)
    sc.lights.append(rc.Light(rc._ffi.LIGHT_POINT, a=(0.1, 0.4, 0.2), b=(0.05, 0.05, 0.05)))
    st = rc.RaytracerSettings(outputs=A.BEAUTY, samples_per_pixel=4, light_sample_count=12, max_ray_depth=3)
    out, stats = gpu_render(rc, sc, st)
    ref, ostats = oracle.render(sc, st)
    assert stats["shadow_rays"] > 0
    assert_beauty_parity(out.beauty, ref.beauty)


def test_builders_give_identical_frames(rc, monkeypatch):
    """PLOC (default) and LBVH trees: bit-identical frames (closest hit is independent of the tree), fewer node fetches"""
    sc = load_scene("cbbunny_area_light_transforms", 320, 180)
    st = rc.RaytracerSettings(outputs=A.BEAUTY | A.NORMALS | A.DEBUG_IDS | A.DEBUG_DEPTH, samples_per_pixel=4)
    a, sa = gpu_render(rc, sc, st, collect_stats=1)
    monkeypatch.setenv("RTCUDA_BUILDER", "lbvh")
    b, sb = gpu_render(rc, sc, st, collect_stats=1)
    for plane in ("beauty", "normals", "debug_ids", "debug_depth"):
        assert np.array_equal(getattr(a, plane), getattr(b, plane)), plane
    assert sa["nodes_fetched"] < 0.8 * sb["nodes_fetched"]
    assert sa["bvh_fallback_lbvh"] == 0 and sb["bvh_fallback_lbvh"] == 0
    # a PLOC tree deeper than the traversal stack covers is rebuilt as an LBVH instead of refusing the scene
    monkeypatch.delenv("RTCUDA_BUILDER")
    monkeypatch.setenv("RTCUDA_TEST_PLOC_TOO_DEEP", "1")
    c, sc_stats = gpu_render(rc, sc, st, collect_stats=1)
    assert sc_stats["bvh_fallback_lbvh"] == 1 and abs(sc_stats["nodes_fetched"] - sb["nodes_fetched"]) < 0.01 * sb["nodes_fetched"]
    assert np.array_equal(c.beauty, a.beauty) and np.array_equal(c.debug_ids, a.debug_ids)


def test_cached_memory_release(rc):
    """the path-state arena of a closed renderer is parked and reused; rtcuda_release_cached_memory gives it back"""
    sc = load_scene("cb", 64, 64)
    st = rc.RaytracerSettings(outputs=A.BEAUTY, samples_per_pixel=2, light_sample_count=1)
    a, _ = gpu_render(rc, sc, st)
    b, _ = gpu_render(rc, sc, st)      # second context takes the parked arena
    rc._ffi.load_library().rtcuda_release_cached_memory()
    c, _ = gpu_render(rc, sc, st)
    assert np.array_equal(a.beauty, b.beauty) and np.array_equal(a.beauty, c.beauty)


def test_watertight_mode(rc, oracle):
    """RTCUDA_BACKEND_WATERTIGHT: same first hits as the reference's test up to grazing edges, same image statistically"""
    sc = load_scene("cbbunny_area_light_transforms", 320, 180)
    st = rc.RaytracerSettings(outputs=dbg() | A.BEAUTY, samples_per_pixel=4)
    wt, _ = gpu_render(rc, sc, st, watertight=True)
    mt, _ = gpu_render(rc, sc, st)
    ref, _ = oracle.render(sc, st)
    assert_first_hit_parity(wt, ref)
    assert (wt.debug_ids == mt.debug_ids).all(axis=-1).mean() >= 0.9999
    assert_beauty_parity(wt.beauty, mt.beauty)


# ---------------------------------------------------------------------------------------------------------------------
# The bench configurations themselves against the oracle (same seed, per-pixel gates over the lit pixels)
# ---------------------------------------------------------------------------------------------------------------------
def _bench_config_parity(rc, oracle, sc, st, what, **backend):
    out, stats = gpu_render(rc, sc, st, **backend)
    ref, ostats = oracle.render(sc, st, num_threads=NT)
    fh = assert_first_hit_parity(out, ref)
    rep = assert_beauty_parity(out.beauty, ref.beauty, what=what)
    la, lb = luminance(out.beauty), luminance(ref.beauty)
    assert abs(la.mean() - lb.mean()) <= 1e-3 * lb.mean(), what
    assert stats["primary_rays"] == ostats["primary_rays"]
    print(f"\n[parity] {what}: first hit {fh}; beauty {rep}")
    return out, ref, stats


def test_c3_bench_config_beauty_1080p(rc, oracle):
    """BASELINE config C3 as bench.py runs it (1920x1080, depth 8, 4 light samples) at the spp the oracle affords
    (16 of 256: the per-sample streams are the same at any spp, sample.rs:69-87): per-pixel beauty over the lit pixels"""
    sc = load_scene("cbbunny_area_light_transforms", 1920, 1080)
    st = rc.RaytracerSettings(outputs=A.BEAUTY | A.DEBUG_IDS | A.NORMALS | A.DEBUG_DEPTH, samples_per_pixel=16, max_ray_depth=8, light_sample_count=4)
    _bench_config_parity(rc, oracle, sc, st, "C3 1080p 16 spp")


def test_c2_bench_config_full(rc, oracle):
    """BASELINE config C2 in full: cb.glb 512x512, 64 spp, 1 light sample, depth 8"""
    sc = load_scene("cb", 512, 512)
    st = rc.RaytracerSettings(outputs=A.BEAUTY | A.DEBUG_IDS | A.NORMALS | A.DEBUG_DEPTH, samples_per_pixel=64, max_ray_depth=8, light_sample_count=1)
    _bench_config_parity(rc, oracle, sc, st, "C2 512x512 64 spp")


def test_c4_bench_config_beauty_1080p(rc, oracle):
    """BASELINE config C4: cb_texture.glb 1080p with the trilinear JPEG texture, beauty at 8 of its 128 spp + AOVs"""
    sc = load_scene("cb_texture", 1920, 1080)
    st = rc.RaytracerSettings(outputs=dbg() | A.BEAUTY, samples_per_pixel=8, max_ray_depth=8, light_sample_count=4)
    _bench_config_parity(rc, oracle, sc, st, "C4 1080p 8 spp")


def test_c5_one_million_triangles(rc, oracle):
    """BASELINE config C5's mesh generator at 1024 x 512 quads (1.05 M triangles) through the oracle's BVH2: the device
    PLOC build + collapse on a tree that no longer fits L1, parity of first hits and of the beauty plane"""
    base = load_scene("cb", 512, 512)
    sc = rc.test_scenes.synthetic_mesh_scene(base, 1024, 512)
    st = rc.RaytracerSettings(outputs=A.BEAUTY | A.DEBUG_IDS | A.NORMALS | A.DEBUG_DEPTH, samples_per_pixel=2, light_sample_count=1)
    out, ref, stats = _bench_config_parity(rc, oracle, sc, st, "C5 at 1.05 M triangles")
    assert stats["bvh_prim_count"] == 12 + 2 * 1024 * 512 - 2 * 1024


def test_texture_variants(rc, oracle):
    """Mix / nested Scale, Mirror / Clamp wrap, nearest / bilinear / trilinear, u8 / u16 / f32 images with 2-4 channels
    (texture.rs:235-459, materials/texture.rs:45-68, image.rs:56-121), incl. the device Lanczos3 pyramid of a 40x24 image"""
    from conftest import texture_zoo_scene
    sc = texture_zoo_scene(384, 384)
    st = rc.RaytracerSettings(outputs=dbg() | A.BEAUTY, samples_per_pixel=8, max_ray_depth=4)
    out, ref, _ = _bench_config_parity(rc, oracle, sc, st, "texture zoo", max_paths_in_flight=300000)
    assert len(np.unique(ref.albedo.reshape(-1, 3), axis=0)) > 2000 and ref.mip_level.max() > 0.5


@pytest.mark.parametrize("name", ["rough_metal", "rough_dielectric"])
def test_general_surface_kernel_batched(rc, oracle, name):
    """k_shade<Surface> (every non-Diffuse material) over many batches: 500x500, 16 spp in 200 k-path batches"""
    t = [t for t in rc.test_scenes.all_test_scenes() if t.name == name][0]
    sc, st = t.scene_func(), t.settings_func()
    st.samples_per_pixel = 16
    st.outputs = A.BEAUTY | A.DEBUG_IDS
    out, stats = gpu_render(rc, sc, st, max_paths_in_flight=200000)
    whole, _ = gpu_render(rc, sc, st)
    assert np.array_equal(out.beauty, whole.beauty)                 # batch shape does not change any pixel
    ref, _ = oracle.render(sc, st, num_threads=NT)
    rep = assert_beauty_parity(out.beauty, ref.beauty, what=name, **SPECULAR_GATES)
    la, lb = luminance(out.beauty), luminance(ref.beauty)
    assert abs(np.nanmean(la) - np.nanmean(lb)) <= 2e-3 * abs(np.nanmean(lb))
    print(f"\n[parity] {name} 16 spp: {rep}")


# ---------------------------------------------------------------------------------------------------------------------
# Multi-GPU behind the C ABI (backend_settings.num_devices): one process, one call, the complete frame
# ---------------------------------------------------------------------------------------------------------------------
def _device_ids(n):
    """n ranks on a one-GPU box share the GPU (same code path, same pixels); a box with several GPUs runs one rank per
    GPU, at most as many ranks as it has GPUs (ranks that share one GPU of several are not a configuration the library serves:
    two host threads growing one peer-mapped memory pool at the same time fail in cudaMallocAsync on a 2-GPU box)"""
    import torch
    have = torch.cuda.device_count()
    return [0] * n if have == 1 else list(range(min(n, have)))


@pytest.mark.parametrize("n,tile", [(2, 0), (3, 16), (8, 16)])
def test_multi_device_render_is_bit_identical(rc, n, tile):
    """rc.render(scene, settings, CudaBackendSettings(num_devices=N)): tiles dealt to N sub-contexts inside the library, every
    GPU copies its owned pixels into the caller's host planes — every plane equals the single-GPU frame bit for bit"""
    sc = load_scene("cb_texture", 400, 225)
    st = rc.RaytracerSettings(outputs=dbg() | A.BEAUTY, samples_per_pixel=4)
    one, s1 = gpu_render(rc, sc, st)
    ids = _device_ids(n)
    many, sn = gpu_render(rc, sc, st, num_devices=len(ids), device_ids=ids, tile_size=tile, collect_stats=1)
    for plane in ("beauty", "normals", "albedo", "uv", "mip_level", "debug_ids", "debug_depth"):
        assert np.array_equal(getattr(one, plane), getattr(many, plane)), plane
    assert sn["samples"] == s1["samples"] and sn["primary_rays"] == s1["primary_rays"] and sn["shadow_rays"] == s1["shadow_rays"]
    assert sn["bounce_rays"] == s1["bounce_rays"] and sn["nodes_fetched"] > 0


def test_multi_device_device_planes_and_progressive_sums(rc):
    """rtcuda_render_device / rtcuda_render_samples_device / ..._accumulate_device on a multi-device scene: planes on
    device_ids[0], owned pixels of the other GPUs arrive by peer copy; progressive accumulation converges to the frame"""
    import torch
    sc = load_scene("cb", 256, 192)
    st = rc.RaytracerSettings(outputs=A.BEAUTY | A.NORMALS, samples_per_pixel=12, light_sample_count=2)
    full, _ = gpu_render(rc, sc, st)
    ids = _device_ids(4)
    with rc.CudaRenderer(sc, rc.CudaBackendSettings(num_devices=len(ids), device_ids=ids, tile_size=16)) as r:
        dev = f"cuda:{ids[0]}"
        beauty = torch.full((192, 256, 3), 7.0, dtype=torch.float32, device=dev)
        normals = torch.full((192, 256, 3), 7.0, dtype=torch.float32, device=dev)
        torch.cuda.synchronize()
        r.render_device(st, {"beauty": beauty.data_ptr(), "normals": normals.data_ptr()})
        assert np.array_equal(beauty.cpu().numpy(), full.beauty) and np.array_equal(normals.cpu().numpy(), full.normals)
        r.render_samples_device(st, 0, 12, beauty.data_ptr())
        inv = torch.tensor(1.0 / 12, dtype=torch.float32)
        assert np.array_equal((beauty.cpu() * inv).numpy(), full.beauty)
        acc = torch.zeros_like(beauty)
        torch.cuda.synchronize()
        for lo, hi in ((0, 4), (4, 5), (5, 12)):
            r.render_samples_accumulate_device(st, lo, hi, acc.data_ptr())
        np.testing.assert_allclose((acc.cpu() * inv).numpy(), full.beauty, rtol=2e-6, atol=1e-7)
        assert r.render_pixel(st, 100, 100, 0, 2)[0].sample_index == 0
    with pytest.raises(rc._ffi.RtCudaError):
        rc.CudaRenderer(sc, rc.CudaBackendSettings(num_devices=2, device_ids=[0, 99]))
    with pytest.raises(rc._ffi.RtCudaError):
        rc.CudaRenderer(sc, rc.CudaBackendSettings(num_devices=2, device_ids=[0, 0], tile_world=2, tile_rank=1))


def test_progressive_accumulation_single_device(rc):
    """the viewer hook: consecutive sample ranges added in place; plane / hi is the image so far (examples/progressive_viewer.py)"""
    import torch
    sc = load_scene("cbbunny_area_light_transforms", 320, 180)
    st = rc.RaytracerSettings(samples_per_pixel=16)
    with rc.CudaRenderer(sc) as r:
        full = r.render(st).beauty
        acc = torch.zeros((180, 320, 3), dtype=torch.float32, device="cuda:0")
        torch.cuda.synchronize()
        for lo, hi in ((0, 1), (1, 2), (2, 4), (4, 8), (8, 16)):
            r.render_samples_accumulate_device(st, lo, hi, acc.data_ptr())
            img = (acc / hi).cpu().numpy()
            assert np.isfinite(img).all()
        np.testing.assert_allclose(img, full, rtol=3e-6, atol=1e-7)


def test_multi_device_large_mesh_upload_is_shared(rc):
    """geometry above the sharing threshold travels once over PCIe (one slice per GPU) and is forwarded between the GPUs:
    the replicas are identical to a single-GPU upload (same frame)"""
    base = load_scene("cb", 256, 256)
    sc = rc.test_scenes.synthetic_mesh_scene(base, 1024, 512)       # 12 MB of indices, 6 MB of vertices and normals
    st = rc.RaytracerSettings(outputs=A.BEAUTY | A.DEBUG_IDS, samples_per_pixel=2, light_sample_count=1)
    one, _ = gpu_render(rc, sc, st)
    ids = _device_ids(3)
    many, _ = gpu_render(rc, sc, st, num_devices=len(ids), device_ids=ids)
    assert np.array_equal(one.beauty, many.beauty) and np.array_equal(one.debug_ids, many.debug_ids)
