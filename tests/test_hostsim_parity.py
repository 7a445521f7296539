"""CPU parity of the CUDA library's per-thread kernel bodies (compiled for the host, tests/hostsim) against
the oracle. This is how the builder / traversal / shading logic is exercised in the GPU-less build container;
the `-m gpu` tests repeat the same comparisons through libraytracing_cuda.so on a B200."""
import numpy as np
import pytest

from conftest import load_scene, bunny_mesh, instanced_bunnies_scene
from parity import assert_first_hit_parity, assert_beauty_parity, luminance, beauty_close

A = None


@pytest.fixture(autouse=True)
def _flags(rc):
    global A
    A = rc.AovFlags


def dbg():
    return A.NORMALS | A.UV_COORDS | A.ALBEDO | A.MIP_LEVEL | A.DEBUG_IDS | A.DEBUG_DEPTH


def test_sphere_c1(rc, oracle, hostsim):
    sc = rc.test_scenes.sphere_scene()
    st = rc.RaytracerSettings(outputs=dbg(), samples_per_pixel=4, max_ray_depth=5)
    out, stats = hostsim.render(sc, st)
    ref, _ = oracle.render(sc, st)
    assert stats["aov_rays"] == 160000
    assert_first_hit_parity(out, ref)


@pytest.mark.parametrize("name", ["cube", "cube_orthographic"])
def test_builtin_normals(rc, oracle, hostsim, name):
    t = [t for t in rc.test_scenes.all_test_scenes() if t.name == name][0]
    sc, st = t.scene_func(), t.settings_func()
    st.outputs = dbg()
    out, _ = hostsim.render(sc, st)
    ref, _ = oracle.render(sc, st)
    assert_first_hit_parity(out, ref)


@pytest.mark.parametrize("name,w,h,spp,ls", [("cb", 96, 96, 8, 1), ("cbbunny_area_light_transforms", 96, 54, 4, 4),
                                             ("cbbunny_area_light", 96, 54, 2, 2), ("cb_texture", 96, 54, 4, 4),
                                             ("checker", 96, 54, 2, 1), ("cbbunny", 64, 36, 2, 1)])
def test_gltf_scenes(rc, oracle, hostsim, name, w, h, spp, ls):
    sc = load_scene(name, w, h)
    st = rc.RaytracerSettings(outputs=dbg() | A.BEAUTY, samples_per_pixel=spp, light_sample_count=ls)
    out, stats = hostsim.render(sc, st, capacity=4096)   # several batches: exercises pixel / sample chunking
    ref, ostats = oracle.render(sc, st, num_threads=4)
    assert stats["bvh_node_count"] != 0xdeadbeef
    assert_first_hit_parity(out, ref)
    assert stats["primary_rays"] == ostats["primary_rays"]
    assert abs(stats["bounce_rays"] - ostats["bounce_rays"]) <= max(4, ostats["bounce_rays"] // 2000)
    # same sampler streams => the two renders follow the same paths up to float rounding
    la, lb = luminance(out.beauty), luminance(ref.beauty)
    assert abs(la.mean() - lb.mean()) <= 2e-3 * lb.mean() + 1e-7
    assert_beauty_parity(out.beauty, ref.beauty)


@pytest.mark.parametrize("name", ["checkered_plane", "dielectric", "metal", "rough_metal", "rough_dielectric", "out_of_focus_sphere"])
def test_builtin_materials_and_cameras(rc, oracle, hostsim, name):
    t = [t for t in rc.test_scenes.all_test_scenes() if t.name == name][0]
    sc, st = t.scene_func(), t.settings_func()
    if name != "checkered_plane":
        sc.camera = type(sc.camera).lookat_camera_thin_lens_perspective((0, 0, 0), (0, 0, -5), (0, 1, 0), False, 0.7853982, 64, 64, 0.1, 3.0) \
            if name == "out_of_focus_sphere" else _small_cornell_camera(rc)
    if st.sampler.kind != "stratified":
        st.samples_per_pixel = min(st.samples_per_pixel, 4)
    st.outputs = dbg() | A.BEAUTY
    out, stats = hostsim.render(sc, st)
    ref, ostats = oracle.render(sc, st, num_threads=4)
    assert_first_hit_parity(out, ref)
    la, lb = luminance(out.beauty), luminance(ref.beauty)
    assert np.isnan(la).sum() == np.isnan(lb).sum()
    assert abs(np.nanmean(la) - np.nanmean(lb)) <= 5e-3 * abs(np.nanmean(lb)) + 1e-7


def _small_cornell_camera(rc):
    import math
    return rc.Camera.lookat_camera_perspective((0.0, 1.0 + 3.4, 0.4), (0, 0, 0.75), (0, 0, 1), False,
                                               float(np.float32(37.8) * np.float32(math.pi / 180)), 64, 64)


def test_coated_diffuse_bunny(rc, oracle, hostsim):
    sc = rc.test_scenes.coated_diffuse_bunny_scene(bunny=bunny_mesh())
    sc.camera = _small_cornell_camera(rc)
    st = rc.RaytracerSettings(outputs=dbg() | A.BEAUTY, samples_per_pixel=1, max_ray_depth=3, light_sample_count=1)
    out, _ = hostsim.render(sc, st)
    ref, _ = oracle.render(sc, st, num_threads=8)
    assert_first_hit_parity(out, ref)
    la, lb = luminance(out.beauty), luminance(ref.beauty)
    assert abs(np.nanmean(la) - np.nanmean(lb)) <= 2e-2 * abs(np.nanmean(lb))


def test_environment_light(rc, oracle, hostsim):
    sc = rc.test_scenes.environment_lighting_scene(rc.test_scenes.synthetic_environment_map())
    # slightly off-axis: with the reference's axis-aligned view the cube face's diagonal runs exactly through
    # pixel centres, and which of the two triangles owns an edge hit is a BVH-order tie
    sc.camera = rc.Camera.lookat_camera_perspective((0.013, 0, 0.007), (0.1, 1, 0.05), (0, 0, 1), False, 0.66, 64, 64)
    st = rc.RaytracerSettings(outputs=dbg() | A.BEAUTY, samples_per_pixel=4)
    out, _ = hostsim.render(sc, st)
    ref, _ = oracle.render(sc, st, num_threads=4)
    assert_first_hit_parity(out, ref)
    assert np.abs(out.beauty - ref.beauty).max() <= 1e-3


def test_stratified_sampler(rc, oracle, hostsim):
    sc = load_scene("cb", 64, 64)
    st = rc.RaytracerSettings(outputs=A.BEAUTY, samples_per_pixel=9, light_sample_count=2, sampler=rc.Sampler.stratified(True, 3, 3))
    out, _ = hostsim.render(sc, st)
    ref, _ = oracle.render(sc, st, num_threads=4)
    assert_beauty_parity(out.beauty, ref.beauty)
    st2 = rc.RaytracerSettings(outputs=A.BEAUTY, samples_per_pixel=9, light_sample_count=2, sampler=rc.Sampler.stratified(False, 3, 3))
    out2, _ = hostsim.render(sc, st2)
    ref2, _ = oracle.render(sc, st2, num_threads=4)
    assert_beauty_parity(out2.beauty, ref2.beauty)


def test_settings_variants(rc, oracle, hostsim):
    sc = load_scene("cb", 48, 48)
    for kw in (dict(max_ray_depth=1), dict(max_ray_depth=3, accumulate_bounces=False), dict(seed=7), dict(antialias_primary_rays=False)):
        st = rc.RaytracerSettings(outputs=A.BEAUTY, samples_per_pixel=4, light_sample_count=1, **kw)
        out, _ = hostsim.render(sc, st)
        ref, _ = oracle.render(sc, st, num_threads=4)
        assert_beauty_parity(out.beauty, ref.beauty, what=str(kw))
        assert abs(out.beauty.mean() - ref.beauty.mean()) <= 2e-3 * ref.beauty.mean() + 1e-8, kw


def test_tile_partition_sums_to_full_frame(rc, hostsim):
    sc = load_scene("cb", 160, 130)
    st = rc.RaytracerSettings(outputs=A.BEAUTY | A.NORMALS, samples_per_pixel=2, light_sample_count=1)
    full, _ = hostsim.render(sc, st)
    parts = [hostsim.render(sc, st, tile_rank=r, tile_world=3)[0] for r in range(3)]
    owner = rc.multi_gpu.tile_owner_map(160, 130, 3)
    for r, p in enumerate(parts):
        assert (p.beauty[owner != r] == 0).all()
    assert np.array_equal(sum(p.beauty for p in parts), full.beauty)
    assert np.array_equal(sum(p.normals for p in parts), full.normals)


def test_empty_scene(rc, hostsim):
    b = rc.SceneBuilder()
    b.add_camera(rc.Camera.lookat_camera_perspective((0, 0, 0), (0, 0, -1), (0, 1, 0), False, 0.7, 32, 32))
    out, _ = hostsim.render(b.build(), rc.RaytracerSettings(outputs=A.BEAUTY | A.NORMALS, samples_per_pixel=2))
    assert (out.beauty == 0).all() and (out.normals == 0).all()


def test_many_light_samples_two_pass_nee(rc, oracle, hostsim):
    """more light samples per vertex than the thread-local staging holds (rt_integrator.h NEE_STAGE = 8): shade counts the
    shadow rays in one pass and writes them in a second one; plus a point light so that K mixes light kinds"""
    sc = load_scene("cb", 48, 48)
    sc.lights.append(rc.Light(rc._ffi.LIGHT_POINT, a=(0.1, 0.4, 0.2), b=(0.05, 0.05, 0.05)))
    st = rc.RaytracerSettings(outputs=A.BEAUTY, samples_per_pixel=2, light_sample_count=12, max_ray_depth=3)
    out, stats = hostsim.render(sc, st, capacity=2048)
    ref, ostats = oracle.render(sc, st, num_threads=4)
    assert stats["shadow_rays"] > 0
    assert_beauty_parity(out.beauty, ref.beauty)
    assert abs(out.beauty.mean() - ref.beauty.mean()) <= 2e-3 * ref.beauty.mean() + 1e-8


def test_builders_give_identical_frames(rc, hostsim, monkeypatch):
    """closest hits do not depend on the tree (conservative culling, id-ordered ties): the PLOC tree and the LBVH tree
    (RTCUDA_BUILDER=lbvh, the A/B switch of rt_build.h) render bit-identical frames, with fewer node visits for PLOC"""
    sc = load_scene("cbbunny_area_light_transforms", 64, 36)
    st = rc.RaytracerSettings(outputs=A.BEAUTY | A.NORMALS | A.DEBUG_IDS | A.DEBUG_DEPTH, samples_per_pixel=2)
    a, sa = hostsim.render(sc, st)
    monkeypatch.setenv("RTCUDA_BUILDER", "lbvh")
    b, sb = hostsim.render(sc, st)
    for plane in ("beauty", "normals", "debug_ids", "debug_depth"):
        assert np.array_equal(getattr(a, plane), getattr(b, plane)), plane
    assert sa["nodes_fetched"] < 0.8 * sb["nodes_fetched"]


def test_watertight_triangle_never_leaks_on_shared_edges(hostsim):
    """Woop's test (RTCUDA_BACKEND_WATERTIGHT): a ray through a point of the edge shared by two triangles hits at least one of
    them, for every such ray; agrees with Moller-Trumbore in the interior. (The reference's test is not watertight,
    geometry.rs:301-340; it stays the default for parity.)"""
    rng = np.random.default_rng(3)
    leaks_wt = leaks_mt = 0
    for _ in range(600):
        a, b, c, d = (rng.standard_normal(3).astype(np.float32) for _ in range(4))   # triangles (a, b, c) and (b, a, d) share edge ab
        s = np.float32(rng.random())
        on_edge = (a + s * (b - a)).astype(np.float32)
        o = (on_edge + rng.standard_normal(3) * 3).astype(np.float32)
        dirv = (on_edge - o).astype(np.float32)
        hit_wt = hostsim.ray_triangle(1, a, b, c, o, dirv)[0] or hostsim.ray_triangle(1, b, a, d, o, dirv)[0]
        hit_mt = hostsim.ray_triangle(0, a, b, c, o, dirv)[0] or hostsim.ray_triangle(0, b, a, d, o, dirv)[0]
        # the silhouette case (both triangles on the same side of the ray plane) may legitimately miss both: only count
        # rays for which c and d lie on opposite sides of the plane through the ray and the edge
        n = np.cross(dirv.astype(np.float64), (b - a).astype(np.float64))
        if np.dot(n, c - on_edge) * np.dot(n, d - on_edge) >= 0:
            continue
        leaks_wt += not hit_wt
        leaks_mt += not hit_mt
    assert leaks_wt == 0
    for _ in range(300):   # interior points: same hit, t / barycentrics equal to rounding
        a, b, c = (rng.standard_normal(3).astype(np.float32) for _ in range(3))
        w = rng.dirichlet((2, 2, 2))
        p = (w[0] * a + w[1] * b + w[2] * c).astype(np.float32)
        o = (p + rng.standard_normal(3) * 2).astype(np.float32)
        dirv = (p - o).astype(np.float32)
        h0, t0, u0, v0 = hostsim.ray_triangle(0, a, b, c, o, dirv)
        h1, t1, u1, v1 = hostsim.ray_triangle(1, a, b, c, o, dirv)
        if min(w) > 1e-3:
            assert h0 and h1
            assert abs(t0 - t1) < 1e-3 * max(1.0, abs(t0)) and abs(u0 - u1) < 1e-3 and abs(v0 - v1) < 1e-3


def test_watertight_mode_matches_reference_first_hits(rc, oracle, hostsim, monkeypatch):
    sc = load_scene("cbbunny_area_light_transforms", 96, 54)
    st = rc.RaytracerSettings(outputs=dbg() | A.BEAUTY, samples_per_pixel=2)
    monkeypatch.setenv("HOSTSIM_WATERTIGHT", "1")
    out, _ = hostsim.render(sc, st)
    ref, _ = oracle.render(sc, st, num_threads=4)
    assert_first_hit_parity(out, ref)
    assert_beauty_parity(out.beauty, ref.beauty)


@pytest.mark.parametrize("name", ["cb_texture", "cbbunny_area_light_transforms"])
def test_shading_records_are_copies(rc, hostsim, monkeypatch, name):
    """the per-primitive shading records (rt_scene.h ShadeRec) hold copies of the mesh attributes: frames are bit-identical
    with and without them (HOSTSIM_NO_SHADE_RECS takes the long way through instance -> indices -> vertex arrays)"""
    sc = load_scene(name, 64, 36)
    st = rc.RaytracerSettings(outputs=A.BEAUTY, samples_per_pixel=2, antialias_primary_rays=False)
    a, _ = hostsim.render(sc, st)
    monkeypatch.setenv("HOSTSIM_NO_SHADE_RECS", "1")
    b, _ = hostsim.render(sc, st)
    assert np.array_equal(a.beauty, b.beauty)


def test_texture_variants(rc, oracle, hostsim):
    """Mix / nested Scale textures, Mirror and Clamp wrap, nearest / bilinear / trilinear, u8 / u16 / f32 images with 2-4
    channels (texture.rs:235-459, materials/texture.rs:45-68, image.rs:56-121): kernel bodies vs the oracle"""
    from conftest import texture_zoo_scene
    sc = texture_zoo_scene(96, 96)
    st = rc.RaytracerSettings(outputs=dbg() | A.BEAUTY, samples_per_pixel=4, max_ray_depth=3)
    out, _ = hostsim.render(sc, st)
    ref, _ = oracle.render(sc, st, num_threads=4)
    assert_first_hit_parity(out, ref)
    assert len(np.unique(ref.albedo.reshape(-1, 3), axis=0)) > 500       # the walls really are textured
    assert ref.mip_level.max() > 0.5                                     # the trilinear wall reports a mip level
    assert_beauty_parity(out.beauty, ref.beauty)


def test_pixels_outside_the_scene_rectangle_are_dropped_exactly(rc, oracle, hostsim, monkeypatch):
    """rt_cull.h: pixels whose camera rays cannot reach the scene bounds take no path slots (api.cu build_pixel_list).
    Every pixel outside the rectangle is a miss in the oracle (ids NONE, beauty exactly 0), frames are bit-identical with
    and without the rule, and cameras / scenes that rule it out keep every pixel."""
    import math
    sc = load_scene("cbbunny_area_light_transforms", 192, 108)
    rect = hostsim.raster_rect(sc)
    assert rect is not None
    x0, y0, x1, y1 = rect
    inside = np.zeros((108, 192), dtype=bool)
    inside[y0:y1 + 1, x0:x1 + 1] = True
    assert 0.2 < inside.mean() < 0.45                       # the box covers about a quarter of the 16:9 raster
    st = rc.RaytracerSettings(outputs=A.BEAUTY | A.DEBUG_IDS, samples_per_pixel=4)
    ref, ostats = oracle.render(sc, st, num_threads=4)
    assert (ref.debug_ids[~inside] == 0xffffffff).all() and (ref.beauty[~inside] == 0).all()
    hit = ref.debug_ids[..., 0] != 0xffffffff
    ys, xs = np.nonzero(hit)
    assert x0 <= xs.min() and xs.max() <= x1 and y0 <= ys.min() and ys.max() <= y1
    assert xs.min() - x0 <= 5 and x1 - xs.max() <= 5        # ... and the rectangle is tight (2 px margin + the 1e-3 growth)
    a, sa = hostsim.render(sc, st)
    monkeypatch.setenv("HOSTSIM_NO_PIXEL_CULL", "1")
    b, sb = hostsim.render(sc, st)
    assert np.array_equal(a.beauty, b.beauty) and sa["primary_rays"] == sb["primary_rays"] == ostats["primary_rays"]
    monkeypatch.delenv("HOSTSIM_NO_PIXEL_CULL")
    # camera inside the box / next to the geometry: the bounds straddle the camera plane -> no rectangle
    inside_cam = load_scene("cb", 64, 64)
    inside_cam.camera = rc.Camera.lookat_camera_perspective((0.0, 0.3, 0.0), (0, 0.3, 1.0), (0, 1, 0), False, 0.8, 64, 64)
    assert hostsim.raster_rect(inside_cam) is None
    # thin lens and environment light: never
    assert hostsim.raster_rect(rc.test_scenes.out_of_focus_sphere_scene()) is None
    assert hostsim.raster_rect(rc.test_scenes.environment_lighting_scene(rc.test_scenes.synthetic_environment_map())) is None
    # orthographic camera
    ortho = rc.test_scenes.cube_orthographic_scene()
    r = hostsim.raster_rect(ortho)
    st2 = rc.RaytracerSettings(outputs=A.DEBUG_IDS)
    oref, _ = oracle.render(ortho, st2)
    if r is not None:
        ins = np.zeros(oref.debug_ids.shape[:2], dtype=bool)
        ins[r[1]:r[3] + 1, r[0]:r[2] + 1] = True
        assert (oref.debug_ids[~ins] == 0xffffffff).all()
        ys, xs = np.nonzero(oref.debug_ids[..., 0] != 0xffffffff)
        assert r[0] <= xs.min() and xs.max() <= r[2] and xs.min() - r[0] <= 6


def test_reused_mesh_instances(rc, oracle, hostsim):
    """one mesh under three Transform primitives (translation, rotation + scale, non-uniform scale): ids, normals (inverse
    transpose), uv, depth of the first hit and the beauty plane against the oracle"""
    sc = instanced_bunnies_scene(96, 72)
    assert len(sc.instances) == len(sc.shapes) + 2
    st = rc.RaytracerSettings(outputs=dbg() | A.BEAUTY, samples_per_pixel=2, max_ray_depth=4, light_sample_count=2)
    out, stats = hostsim.render(sc, st)
    ref, ostats = oracle.render(sc, st, num_threads=8)
    assert_first_hit_parity(out, ref)
    geoms = set(np.unique(out.debug_ids[..., 0]).tolist())
    assert {len(sc.instances) - 3, len(sc.instances) - 2, len(sc.instances) - 1} <= geoms   # all three instances are seen
    assert stats["primary_rays"] == ostats["primary_rays"]
    assert_beauty_parity(out.beauty, ref.beauty)


@pytest.mark.parametrize("name", ["rough_metal", "dielectric"])
def test_material_split_gives_the_same_frame(rc, hostsim, monkeypatch, name):
    """mixed materials: the Diffuse shade body first + the general body over the vertices it left (kernels.cu launch_shade) against
    everything through the general body — the same frame (the Diffuse branches of the two bodies are the same arithmetic)"""
    t = [t for t in rc.test_scenes.all_test_scenes() if t.name == name][0]
    sc, st = t.scene_func(), t.settings_func()
    sc.camera = _small_cornell_camera(rc)
    st.samples_per_pixel, st.outputs = 4, A.BEAUTY
    split, s1 = hostsim.render(sc, st)
    monkeypatch.setenv("HOSTSIM_NO_MATERIAL_SPLIT", "1")
    single, s2 = hostsim.render(sc, st)
    assert s1["bounce_rays"] == s2["bounce_rays"] and s1["shadow_rays"] == s2["shadow_rays"]
    assert np.array_equal(split.beauty, single.beauty, equal_nan=True)
