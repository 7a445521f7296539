// rt_scene.h — the scene as it lives in HBM (DESIGN.md "Data layout").
//
// Geometry is stored twice on purpose:
//   * object-space mesh arrays exactly as uploaded (vertices / tris / normals / uvs): used by `shade`
//     to reconstruct the reference's hit record (interpolated normal, uv, dpdu/dpdv, geometry.rs:229-298)
//     and by next-event estimation (lights.rs:61-112), both of which the reference evaluates in
//     object space;
//   * a packed world-space primitive array (48 B / primitive, three 128-bit loads) ordered by the
//     8-wide BVH: this is what the traversal kernels touch.
#pragma once
#include "rt_common.h"

#ifndef RT_FIXED_SLOTS
#define RT_FIXED_SLOTS 1   // 0: the dense primitive packing + per-child meta bytes of round 1 (A/B switch)
#endif

namespace rt {

// 48-byte traversal primitive. Triangle: world-space vertices, ids in the w lanes.
// Sphere: a = (center_obj.xyz, geom_id), b = (radius, -, -, prim_id=0), c.w = 1.
struct Prim {
    float4 a;  // v0.xyz | geom_id
    float4 b;  // v1.xyz | prim_id
    float4 c;  // v2.xyz | kind (0 triangle, 1 sphere)
};

// One child of the root aggregate (rtcuda_instance + the rtcuda_shape it refers to).
struct Instance {
    M4 o2w;       // Transform::forward
    M4 w2o;       // Transform::inverse
    uint32_t shape, kind, material, area_light;
    uint32_t vertex_offset, tri_offset, normal_offset, uv_offset;
    uint32_t tri_count, prim_base, _p0, _p1;   // prim_base: first build-primitive of this instance
    float center[3];
    float radius;
};

struct ShapeD {  // rtcuda_shape (lights address emitters by shape index)
    uint32_t kind, material, area_light, vertex_offset, vertex_count, tri_offset, tri_count, normal_offset, uv_offset;
    float center[3];
    float radius;
};

struct LightD {
    uint32_t kind, shape;
    float a[3];   // position / direction
    float b[3];   // intensity / radiance
    M4 light_to_world;
    uint32_t tri_table;   // diffuse area light: first record of its emitter in SceneD::light_tris
    uint32_t _pad[3];
};

// One emitter triangle as next-event estimation reads it (four 128-bit loads instead of three index loads and nine
// scalar vertex loads, and no per-sample cross product / square root for the values that depend on the triangle only):
// object-space vertices, Mesh::tri_area (mesh.rs:271-278) and, for meshes without vertex normals, the geometric normal
// unit(cross(p2 - p0, p1 - p0)) of lights.rs:93-95. Computed once per scene by light_tri_body with the reference's
// formulas, so sample_light returns what it would have computed from the mesh arrays.
// For emitters WITH vertex normals the three object-space normals of the triangle ride along (copies of the mesh array:
// lights.rs:96-105 interpolates them per sample), which spares three index loads and nine scalar normal loads per light sample.
struct LightTri { float4 p0_area, p1_nx, p2_ny, nz, n0, n1, n2; };

struct MaterialD { uint32_t kind, remap_roughness, albedo, eta, kappa, roughness, thickness, coat_albedo; };

struct TextureD {
    uint32_t kind, image, filter, wrap, a, b, c, mip_base;  // mip_base: index into mips[] or NONE
    float value[4];
    float value2[4];
};

struct ImageD {  // also used for generated mip levels
    uint32_t width, height, channels, format;
    uint64_t byte_offset;  // into image_bytes
};

struct MipChain {  // CpuMipmap (texture.rs:114-165): level 0 = resized mip0, then halvings down to 1x1
    uint32_t first_image;  // index into images[] of mip0
    uint32_t level_count;  // mips.len() + 1
};

struct CameraD {
    uint32_t kind, width, height;
    float near_clip, far_clip, aperture_radius, focal_distance;
    M4 raster_to_camera, camera_to_world;  // forward matrices only (lib.rs:145-195)
};

// What shade needs to rebuild the reference's hit record at a triangle, gathered once per scene in BVH primitive order
// (five 128-bit loads from one place instead of the chain packed primitive -> instance -> index triple -> three vertex
// normals / uvs, each a dependent round trip in a kernel that is bound by exactly that latency): the object-space vertex
// normals (or, for meshes without normals, the geometric normal unit(cross(p2 - p0, p1 - p0)), geometry.rs:244-247),
// the vertex uvs, and the ids shade would otherwise fetch through the instance. Values are copies, so the arithmetic of
// reconstruct_hit is unchanged. Spheres and the depth-0 anti-aliased path (needs object-space positions) keep the long way.
struct ShadeRec {
    float4 n0_geom;    // n0.xyz (or the flat normal) | geom_id
    float4 n1_prim;    // n1.xyz | prim_id
    float4 n2_flags;   // n2.xyz | REC_* bits
    float4 uv01;       // uv0.xy, uv1.xy
    float4 uv2_mat;    // uv2.xy | material | area light
};
enum { REC_FLAT = 1u, REC_UV = 2u, REC_SPHERE = 4u };
constexpr uint32_t PRIM_HOLE = 0xffffffffu;   // Prim::c.w of an unfilled primitive slot (the array is memset to 0xff before the build)

// 80-byte compressed 8-wide node (DESIGN.md "BVH8 node"): five 128-bit loads.
//   n0 = origin.xyz | ex | ey<<8 | ez<<16 | imask<<24
//   n1 = child_base | prim_base | valid | -
//   n2 = qlo_x[0..7] qlo_y[0..7]   n3 = qlo_z[0..7] qhi_x[0..7]   n4 = qhi_y[0..7] qhi_z[0..7]
// valid: bit 24+s = child s is an inner node (the imask again); bits 3s..3s+2 = unary primitive count (1, 3, 7) of leaf child s,
// whose primitives are prims[prim_base + 3s + k]. Leaf children occupy the lowest slots, so a node owns 3 * n_leaves consecutive
// primitive slots; slots a leaf does not fill are holes (PRIM_HOLE) that no ray ever reads.
// (RT_FIXED_SLOTS=0, the round-1 layout: n1.zw = 8 meta bytes, 0 = empty; inner: (1<<5) | (24+i); leaf: unary count << 5 | prim offset.)
struct Node8 { float4 n0, n1, n2, n3, n4; };

struct SceneD {
    CameraD camera;
    const Node8* nodes;
    const Prim* prims;
    const ShadeRec* shade_recs;   // by packed primitive index, or null
    uint32_t prim_count;
    uint32_t node_count;
    const Instance* instances;
    const ShapeD* shapes;
    const LightD* lights;
    const LightTri* light_tris;
    const MaterialD* materials;
    const TextureD* textures;
    const ImageD* images;
    const MipChain* mips;
    const uint8_t* image_bytes;
    const float* vertices;
    const uint32_t* tris;
    const float* normals;
    const float* uvs;
    uint32_t instance_count, light_count, material_count, texture_count;
    uint32_t env_texture;
    uint32_t all_diffuse;     // every material is Diffuse: shade with the DiffuseSurface instantiation
    uint32_t tex_uses_derivs; // some texture is an image or a checker: the only consumers of the uv derivatives (MatCtx); without
                              // one, the primary hit skips the camera-ray differentials (same values: nothing would read them)
    uint32_t watertight;      // RTCUDA_BACKEND_WATERTIGHT: Woop's watertight triangle test instead of the reference's Moller-Trumbore
    // Scene-wide constants of the shade kernel that live in the KERNEL PARAMETER block (constant bank: operands of the
    // arithmetic, no load instruction, no scoreboard wait) instead of behind pointers. ncu, round 2: a C3 vertex issued 207
    // global loads, 180 of them in the four light samples — the light record, its emitter's shape record and the emitter's
    // vertex normals, all at warp-uniform addresses (profiles/r4c_ncu_summary.md).
    //   light0           copy of lights[0] (valid when light_count >= 1); the other lights are read from memory
    //   light0_tri_count / light0_has_normals   what sample_light needs of the emitter's shape record
    //   mat_const        per material: w = 1: Diffuse with a CONSTANT albedo texture, xyz = that albedo (Diffuse surfaces skip
    //                    the material -> texture -> value chain of dependent loads); w = 2: Diffuse, albedo from its texture;
    //                    w = 0: any other material (the Diffuse shade kernel of a mixed scene leaves those vertices to the
    //                    general kernel)
    //   any_diffuse      some material is Diffuse (with all_diffuse == 0: the scene mixes materials)
    LightD light0;
    uint32_t light0_tri_count, light0_has_normals, use_light0, any_diffuse;
    const float4* mat_const;
    float scene_center[3];
    float scene_radius;       // +inf when the BVH root is a leaf (bvh2.rs:448-452 quirk, see rt_shade.h)
    // World bounds of all primitives, grown by 1e-3 of their largest extent: camera rays that miss them are not queued
    // (raygen_body) — the root-AABB reject of traverse_bvh (accel.rs:95) moved in front of the wavefront. Empty scene: lo > hi.
    float bounds_lo[3], bounds_hi[3];
};

RT_HD void set_scene_bounds(SceneD& sc, V3 mn, V3 mx, bool any) {
    const float pad = any ? 1.0e-3f * fmaxf(mx.x - mn.x, fmaxf(mx.y - mn.y, mx.z - mn.z)) + 1.0e-6f : 0.0f;
    const float big = 3.0e38f;
    sc.bounds_lo[0] = any ? mn.x - pad : big; sc.bounds_lo[1] = any ? mn.y - pad : big; sc.bounds_lo[2] = any ? mn.z - pad : big;
    sc.bounds_hi[0] = any ? mx.x + pad : -big; sc.bounds_hi[1] = any ? mx.y + pad : -big; sc.bounds_hi[2] = any ? mx.z + pad : -big;
}

RT_HD V3 load3(const float* p, uint32_t i) { return mk3(ldg(p + 3 * (size_t)i), ldg(p + 3 * (size_t)i + 1), ldg(p + 3 * (size_t)i + 2)); }
RT_HD void light_tri_body(uint32_t tri, const ShapeD& em, const float* vertices, const uint32_t* tris, const float* normals, LightTri* out) {
    const uint32_t* t = tris + 3 * (size_t)(em.tri_offset + tri);
    const V3 p0 = load3(vertices, em.vertex_offset + t[0]), p1 = load3(vertices, em.vertex_offset + t[1]), p2 = load3(vertices, em.vertex_offset + t[2]);
    const float area = length(cross(p1 - p0, p2 - p0)) / 2.0f;
    const V3 n = unit(cross(p2 - p0, p1 - p0));
    LightTri r;
    r.p0_area = make_float4(p0.x, p0.y, p0.z, area);
    r.p1_nx = make_float4(p1.x, p1.y, p1.z, n.x);
    r.p2_ny = make_float4(p2.x, p2.y, p2.z, n.y);
    r.nz = make_float4(n.z, 0.0f, 0.0f, 0.0f);
    r.n0 = r.n1 = r.n2 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (em.normal_offset != NONE) {
        const V3 a = load3(normals, em.normal_offset + t[0]), b = load3(normals, em.normal_offset + t[1]), c = load3(normals, em.normal_offset + t[2]);
        r.n0 = make_float4(a.x, a.y, a.z, 0.0f); r.n1 = make_float4(b.x, b.y, b.z, 0.0f); r.n2 = make_float4(c.x, c.y, c.z, 0.0f);
    }
    out[tri] = r;
}

RT_HD V2 load2(const float* p, uint32_t i) { return mk2(ldg(p + 2 * (size_t)i), ldg(p + 2 * (size_t)i + 1)); }

RT_HD void shade_rec_body(uint32_t i, const SceneD& sc, ShadeRec* out) {
    const Prim& pr = sc.prims[i];
    const uint32_t geom = f2u(pr.a.w), prim_id = f2u(pr.b.w), kind = f2u(pr.c.w);
    const Instance& inst = sc.instances[kind == 0xffffffffu ? 0u : geom];
    ShadeRec r;
    uint32_t flags = 0;
    if (kind == PRIM_HOLE) {   // unfilled slot: never referenced by a hit
        r.n0_geom = r.n1_prim = r.n2_flags = r.uv01 = r.uv2_mat = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        out[i] = r;
        return;
    }
    V3 n0 = mk3(0.0f), n1 = mk3(0.0f), n2 = mk3(0.0f);
    V2 uv0 = mk2(0, 0), uv1 = mk2(1, 0), uv2 = mk2(0, 1);
    if (kind != 0) flags = REC_SPHERE;
    else {
        const uint32_t* t = sc.tris + 3 * (size_t)(inst.tri_offset + prim_id);
        const uint32_t i0 = t[0], i1 = t[1], i2 = t[2];
        if (inst.normal_offset == NONE) {
            const V3 p0 = load3(sc.vertices, inst.vertex_offset + i0), p1 = load3(sc.vertices, inst.vertex_offset + i1),
                     p2 = load3(sc.vertices, inst.vertex_offset + i2);
            n0 = unit(cross(p2 - p0, p1 - p0));
            flags |= REC_FLAT;
        } else {
            n0 = load3(sc.normals, inst.normal_offset + i0);
            n1 = load3(sc.normals, inst.normal_offset + i1);
            n2 = load3(sc.normals, inst.normal_offset + i2);
        }
        if (inst.uv_offset != NONE) {
            uv0 = load2(sc.uvs, inst.uv_offset + i0);
            uv1 = load2(sc.uvs, inst.uv_offset + i1);
            uv2 = load2(sc.uvs, inst.uv_offset + i2);
            flags |= REC_UV;
        }
    }
    r.n0_geom = make_float4(n0.x, n0.y, n0.z, u2f(geom));
    r.n1_prim = make_float4(n1.x, n1.y, n1.z, u2f(prim_id));
    r.n2_flags = make_float4(n2.x, n2.y, n2.z, u2f(flags));
    r.uv01 = make_float4(uv0.x, uv0.y, uv1.x, uv1.y);
    r.uv2_mat = make_float4(uv2.x, uv2.y, u2f(inst.material), u2f(inst.area_light));
    out[i] = r;
}

}  // namespace rt
