// rt_traverse.h — closest-hit (`extend`) and any-hit (`shadow`) traversal of the compressed 8-wide BVH.
//
// Replaces traverse_bvh (crates/raytracing-cpu/src/accel.rs:65-258) + intersect_aabb /
// ray_triangle_intersect / ray_sphere_intersect (geometry.rs:51-78, 139-227, 301-340). Same answers
// (closest t in [t_min, t_max], inclusive bounds, no back-face culling, Möller–Trumbore), different
// structure: one ray per thread, an octant-ordered walk over 80-byte quantised nodes (five 128-bit
// loads) and 48-byte world-space triangles (three 128-bit loads), a node-group / primitive-group
// bit-mask stack instead of the reference's per-entry progress counters.
#pragma once
#include "rt_scene.h"

namespace rt {

struct Hit {
    float t;
    uint32_t prim;  // index into SceneD::prims, NONE on miss
    float u, v;     // triangle barycentrics (weights of v1, v2)
};

struct TraverseStats { uint32_t nodes, prims; };

constexpr int TRAVERSE_STACK = 64;  // 8 B entries; a ray holds at most two per level of the wide tree (api.cu refuses deeper trees)

// geometry.rs:139-227 — the root finding part only (the hit record is rebuilt in shade).
RT_HD bool sphere_t(V3 center, float radius, V3 o, V3 d, float t_min, float t_max, float& t_out) {
    V3 omc = o - center;
    float a = sqmag(d);
    float b = 2.0f * dot(d, omc);
    float c = sqmag(omc) - radius * radius;
    float disc = b * b - 4.0f * a * c;
    float t1, t2;
    if (disc < 0.0f) return false;
    if (disc == 0.0f) { t1 = t2 = -b / (2.0f * a); }
    else {
        float q = -0.5f * (b + rs_signum(b) * sqrtf(disc));
        t1 = q / a;
        t2 = c / q;
    }
    if (t1 > t2) { float s = t1; t1 = t2; t2 = s; }
    if (t1 >= t_min && t1 <= t_max) { t_out = t1; return true; }
    if (t2 >= t_min && t2 <= t_max) { t_out = t2; return true; }
    return false;
}

// geometry.rs:301-340 with world-space vertices (t is shared between spaces: geometry.rs:100-101).
RT_HD bool triangle_t(V3 p0, V3 p1, V3 p2, V3 o, V3 d, float t_min, float t_max, float& t, float& u, float& v) {
    V3 e1 = p1 - p0, e2 = p2 - p0;
    V3 P = cross(d, e2);
    float denom = dot(P, e1);
    if (denom == 0.0f) return false;
    float inv = 1.0f / denom;
    V3 T = o - p0;
    u = dot(P, T) * inv;
    if (u < 0.0f || u > 1.0f) return false;
    V3 Q = cross(T, e1);
    v = dot(Q, d) * inv;
    if (v < 0.0f || u + v > 1.0f) return false;
    t = dot(Q, e2) * inv;
    if (t < t_min || t > t_max) return false;
    return true;
}

// Watertight ray / triangle test (Woop, Benthin, Wald 2013), the north star's intersection for meshes whose silhouettes
// and shared edges must not leak; selected per context (RTCUDA_BACKEND_WATERTIGHT). The reference's Moller-Trumbore above
// stays the default because it is what the parity gates compare against. The scaled edge functions
// are evaluated with un-fused multiplies: a shared edge must produce exactly opposite values in its two triangles, which
// an FMA contraction (one product rounded, the other not) would break. The ray's shear constants are recomputed per test
// instead of being carried per ray: the mode costs registers only where it is used. No back-face culling, inclusive
// t range, barycentrics in the (v1, v2) convention of triangle_t.
#if defined(__CUDA_ARCH__)
RT_HD float mul_rn(float a, float b) { return __fmul_rn(a, b); }
RT_HD float sub_rn(float a, float b) { return __fsub_rn(a, b); }
#else
RT_HD float mul_rn(float a, float b) { volatile float r = a * b; return r; }
RT_HD float sub_rn(float a, float b) { volatile float r = a - b; return r; }
#endif
RT_HD float comp3(V3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }
RT_HD bool triangle_watertight(V3 p0, V3 p1, V3 p2, V3 o, V3 d, float t_min, float t_max, float& t, float& u, float& v) {
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    const int kz = ax > ay ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);
    int kx = kz == 2 ? 0 : kz + 1, ky = kx == 2 ? 0 : kx + 1;
    const float dz = comp3(d, kz);
    if (dz < 0.0f) { const int s = kx; kx = ky; ky = s; }
    const float sz = 1.0f / dz, sx = comp3(d, kx) * sz, sy = comp3(d, ky) * sz;
    const V3 A = p0 - o, B = p1 - o, C = p2 - o;
    const float Ax = sub_rn(comp3(A, kx), mul_rn(sx, comp3(A, kz))), Ay = sub_rn(comp3(A, ky), mul_rn(sy, comp3(A, kz)));
    const float Bx = sub_rn(comp3(B, kx), mul_rn(sx, comp3(B, kz))), By = sub_rn(comp3(B, ky), mul_rn(sy, comp3(B, kz)));
    const float Cx = sub_rn(comp3(C, kx), mul_rn(sx, comp3(C, kz))), Cy = sub_rn(comp3(C, ky), mul_rn(sy, comp3(C, kz)));
    float U = sub_rn(mul_rn(Cx, By), mul_rn(Cy, Bx)), V = sub_rn(mul_rn(Ax, Cy), mul_rn(Ay, Cx)), W = sub_rn(mul_rn(Bx, Ay), mul_rn(By, Ax));
    if (U == 0.0f || V == 0.0f || W == 0.0f) {   // on an edge in single precision: decide in double
        U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
        V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
        W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = U + V + W;
    if (det == 0.0f) return false;
    const float T = U * (sz * comp3(A, kz)) + V * (sz * comp3(B, kz)) + W * (sz * comp3(C, kz));
    const float inv = 1.0f / det;
    t = T * inv;
    if (!(t >= t_min && t <= t_max)) return false;
    u = V * inv;
    v = W * inv;
    return true;
}

#ifndef RT_FAST_RCP
#define RT_FAST_RCP 0   // 1: bare MUFU.RCP for 1 / direction (A/B switch; only the conservative box culling reads it)
#endif
RT_HD float safe_rcp_dir(float d) {
    float a = fabsf(d) > 1.0e-20f ? d : copysignf(1.0e-20f, d);
#if defined(__CUDA_ARCH__) && RT_FAST_RCP
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
#else
    return 1.0f / a;
#endif
}

// Quantised plane byte K of word w as the float 1 + q * 2^-15 (q in mantissa bits 8..15): one PRMT on the
// device, no int->float conversion (I2F issues on the quarter-rate XU pipe and was the limiter of the first
// version, profiles/r1_notes.md) and no bias subtraction: the node constants absorb the "1 +" (see node_step).
template <int K>
RT_HD float qfloat(uint32_t w) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(__byte_perm(0x3f800000u, w, 0x3240u + (K << 4)));  // selector as the immediate, the constant in a register
#else
    return u2f(0x3f800000u | (((w >> (8 * K)) & 0xffu) << 8));
#endif
}
template <int K>
RT_HD uint32_t byte_of(uint32_t w) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(w, 0u, 0x4440u + K);
#else
    return (w >> (8 * K)) & 0xffu;
#endif
}

// Traversal as a resumable per-ray state machine with two kinds of work unit, so that a warp can run all its
// lanes through the same kind at once (kernels.cu: the persistent kernels vote per iteration):
//   node_step — pop ONE wide node of the current node group, test its 8 quantised child boxes;
//   tri_step  — intersect ONE primitive of the current primitive group;
//   next      — refill the current groups from the stack; false when the walk is complete.
// Pending primitive groups can be postponed: node_step stashes a non-empty one on the stack (entries with no
// bits in the top byte are primitive groups). The closest hit is independent of the order in which primitives
// are tested: equal-t ties go to the larger (geom_id, prim_id) (the reference resolves them by its own BVH2 leaf
// order, "later hit overwrites", geometry.rs:335 / accel.rs:159-171, which no other tree can reproduce), so
// frames stay bit-reproducible although lanes pick up rays dynamically and the builder packs primitives with atomics.
#ifndef RT_ANYHIT_ORDERED
#define RT_ANYHIT_ORDERED 0   // 1: any-hit walks keep the octant order too (A/B switch)
#endif
template <bool ANY_HIT, bool STATS>
struct Traversal {
    static constexpr bool UNORDERED = ANY_HIT && !RT_ANYHIT_ORDERED;
    V3 o, d, idir;
    float t_min, closest;
    uint32_t octinv;
    uint2 ngroup, tgroup;
    uint2* stack;  // TRAVERSE_STACK entries of thread-local memory owned by the caller (keeps the scalar state in registers)
    int sp;
    Hit hit;
    uint32_t hit_geom, hit_pid;  // ids of the current closest hit (tie-break)
    uint32_t found;  // 32-bit flag (a bool would be packed into a half register)

    RT_HD bool init(const SceneD& sc, V3 o_, V3 d_, float t_min_, float t_max_) {
        o = o_; d = d_; t_min = t_min_; closest = t_max_;
        hit.prim = NONE; hit.t = t_max_; hit.u = hit.v = 0.0f;
        found = 0u;
        hit_geom = hit_pid = 0u;
        sp = 0;
        idir = mk3(safe_rcp_dir(d.x), safe_rcp_dir(d.y), safe_rcp_dir(d.z));
        octinv = 7u - ((d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u));
        ngroup = make_uint2(0u, 0x80000000u);  // root: "child 7^octinv of a group with base 0, imask 0"
        tgroup = make_uint2(0u, 0u);
        return sc.prim_count != 0;
    }

    RT_HD bool has_tris() const { return tgroup.y != 0u; }
    RT_HD bool has_nodes() const { return (ngroup.y & 0xff000000u) != 0u; }
    RT_HD bool can_stash() const { return sp < TRAVERSE_STACK - 2; }

#if RT_FIXED_SLOTS
    // child slot S (plane bytes I = S & 3 of the half's words): slab test; a hit sets the child's FIXED bits of the hit mask —
    // its inner-node bit 24 + S and its three primitive bits 3S..3S+2 — with one immediate; node_step then clears what the
    // node's `valid` word does not back (empty slots, leaves with fewer than three primitives, the wrong kind of child).
    template <int S>
    RT_HD void child_fixed(uint32_t nx_w, uint32_t fx_w, uint32_t ny_w, uint32_t fy_w, uint32_t nz_w, uint32_t fz_w,
                           V3 an, V3 cn, V3 af, V3 cf, float tf_cap, uint32_t& hitmask) const {
        constexpr int I = S & 3;
        const float tn = fmaxf(fmaxf(qfloat<I>(nx_w) * an.x + cn.x, qfloat<I>(ny_w) * an.y + cn.y), fmaxf(qfloat<I>(nz_w) * an.z + cn.z, t_min));
        const float tf = fminf(fminf(qfloat<I>(fx_w) * af.x + cf.x, qfloat<I>(fy_w) * af.y + cf.y), fminf(qfloat<I>(fz_w) * af.z + cf.z, tf_cap));
        if (tn <= tf) hitmask |= (7u << (3 * S)) | (1u << (24 + S));
    }
#endif

    // child I of a 4-child half: slab test against the near / far plane words of the three axes
    template <int I>
    RT_HD void child(uint32_t bits4, uint32_t index4, uint32_t nx_w, uint32_t fx_w, uint32_t ny_w, uint32_t fy_w, uint32_t nz_w, uint32_t fz_w,
                     V3 an, V3 cn, V3 af, V3 cf, float tf_cap, uint32_t& hitmask) const {
        const float tn = fmaxf(fmaxf(qfloat<I>(nx_w) * an.x + cn.x, qfloat<I>(ny_w) * an.y + cn.y), fmaxf(qfloat<I>(nz_w) * an.z + cn.z, t_min));
        const float tf = fminf(fminf(qfloat<I>(fx_w) * af.x + cf.x, qfloat<I>(fy_w) * af.y + cf.y), fminf(qfloat<I>(fz_w) * af.z + cf.z, tf_cap));
        if (tn <= tf) hitmask |= byte_of<I>(bits4) << byte_of<I>(index4);
    }

    // precondition: has_nodes()
    RT_HD void node_step(const SceneD& sc, TraverseStats* stats) {
        RT_CHECK(sp + 2 <= TRAVERSE_STACK);
        if (tgroup.y) stack[sp++] = tgroup;  // postponed primitives
        const uint32_t hits = ngroup.y;
        const int bit = bfind32(hits);
        ngroup.y &= ~(1u << bit);
        if (ngroup.y & 0xff000000u) stack[sp++] = ngroup;
        // closest-hit walks pop the children of a node front to back (hit-mask position = slot ^ octinv); an any-hit walk
        // has no use for the order: positions are the slots themselves, and the per-node octant permutation is not computed
        const uint32_t slot = UNORDERED ? (uint32_t)bit - 24u : ((uint32_t)bit - 24u) ^ octinv;
        const uint32_t rel = (uint32_t)popc32(hits & ~(0xffffffffu << slot) & 0xffu);
        RT_CHECK(ngroup.x + rel < sc.node_count);
        const Node8* node = sc.nodes + (ngroup.x + rel);
        const float4 n0 = ldg(&node->n0), n1 = ldg(&node->n1), n2 = ldg(&node->n2), n3 = ldg(&node->n3), n4 = ldg(&node->n4);
        if (STATS) stats->nodes++;

        // plane coordinate = origin + q * 2^e; along the ray t = q * a + b with a = 2^e / d, b = (origin - o) / d.
        // With qfloat = 1 + q * 2^-15: t = qfloat * A + C, A = 2^15 * a, C = b - A — one FFMA per plane. The rounding
        // of C costs at most 2^-9 quantisation steps; both slab sides are pushed outwards by 2^-7 steps (per axis, so an
        // axis-parallel ray keeps its exact in/out classification on the other axes) and the far side keeps the
        // 4e-7 relative slack of the box test, so culling stays conservative.
        const uint32_t e = f2u(n0.w);
        const V3 a15 = mk3(u2f((byte_of<0>(e) + 15u) << 23) * idir.x, u2f((byte_of<1>(e) + 15u) << 23) * idir.y,
                           u2f((byte_of<2>(e) + 15u) << 23) * idir.z);
        const uint32_t imask = e >> 24;
        const V3 c = mk3((n0.x - o.x) * idir.x - a15.x, (n0.y - o.y) * idir.y - a15.y, (n0.z - o.z) * idir.z - a15.z);
        const V3 pad = mk3(fabsf(a15.x), fabsf(a15.y), fabsf(a15.z)) * 2.3841858e-7f;  // 2^-22 * |A| = 2^-7 steps
        const float k_far = 1.0000004f;
        const V3 cn = c - pad, af = a15 * k_far, cf = (c + pad) * k_far;
        const float tf_cap = closest * k_far;
        const uint32_t meta_lo = f2u(n1.z), meta_hi = f2u(n1.w);
        // near / far plane words per axis, selected by the ray octant
        const uint32_t lox0 = f2u(n2.x), lox1 = f2u(n2.y), loy0 = f2u(n2.z), loy1 = f2u(n2.w);
        const uint32_t loz0 = f2u(n3.x), loz1 = f2u(n3.y), hix0 = f2u(n3.z), hix1 = f2u(n3.w);
        const uint32_t hiy0 = f2u(n4.x), hiy1 = f2u(n4.y), hiz0 = f2u(n4.z), hiz1 = f2u(n4.w);
        const bool nx = (octinv & 1u) == 0u, ny = (octinv & 2u) == 0u, nz = (octinv & 4u) == 0u;  // direction negative on the axis
        const uint32_t nearx0 = nx ? hix0 : lox0, nearx1 = nx ? hix1 : lox1, farx0 = nx ? lox0 : hix0, farx1 = nx ? lox1 : hix1;
        const uint32_t neary0 = ny ? hiy0 : loy0, neary1 = ny ? hiy1 : loy1, fary0 = ny ? loy0 : hiy0, fary1 = ny ? loy1 : hiy1;
        const uint32_t nearz0 = nz ? hiz0 : loz0, nearz1 = nz ? hiz1 : loz1, farz0 = nz ? loz0 : hiz0, farz1 = nz ? loz1 : hiz1;
#if RT_FIXED_SLOTS
        uint32_t hitmask = 0;
        child_fixed<0>(nearx0, farx0, neary0, fary0, nearz0, farz0, a15, cn, af, cf, tf_cap, hitmask);
        child_fixed<1>(nearx0, farx0, neary0, fary0, nearz0, farz0, a15, cn, af, cf, tf_cap, hitmask);
        child_fixed<2>(nearx0, farx0, neary0, fary0, nearz0, farz0, a15, cn, af, cf, tf_cap, hitmask);
        child_fixed<3>(nearx0, farx0, neary0, fary0, nearz0, farz0, a15, cn, af, cf, tf_cap, hitmask);
        child_fixed<4>(nearx1, farx1, neary1, fary1, nearz1, farz1, a15, cn, af, cf, tf_cap, hitmask);
        child_fixed<5>(nearx1, farx1, neary1, fary1, nearz1, farz1, a15, cn, af, cf, tf_cap, hitmask);
        child_fixed<6>(nearx1, farx1, neary1, fary1, nearz1, farz1, a15, cn, af, cf, tf_cap, hitmask);
        child_fixed<7>(nearx1, farx1, neary1, fary1, nearz1, farz1, a15, cn, af, cf, tf_cap, hitmask);
        hitmask &= meta_lo;   // the node's `valid` word
        if (!UNORDERED) {
            // closest-hit walks pop inner children front to back: move the bit of slot s to position s ^ octinv (three
            // conditional swaps of the top byte: neighbours, pairs, nibbles)
            uint32_t top = hitmask >> 24;
            const uint32_t t1 = ((top & 0x55u) << 1) | ((top >> 1) & 0x55u);
            top = (octinv & 1u) ? t1 : top;
            const uint32_t t2 = ((top & 0x33u) << 2) | ((top >> 2) & 0x33u);
            top = (octinv & 2u) ? t2 : top;
            const uint32_t t4 = ((top & 0x0fu) << 4) | (top >> 4);
            top = (octinv & 4u) ? t4 : top;
            hitmask = (hitmask & 0x00ffffffu) | (top << 24);
        }
        (void)meta_hi;
#else
        // four children at a time: hit-mask bit position (24 + slot ^ octinv for inner children, the primitive offset for
        // leaves) and the bits to set there (1 for inner, unary primitive count for leaves; 0 for an empty child)
        const uint32_t octinv4 = octinv * 0x01010101u;
        uint32_t hitmask = 0;
        {
            const uint32_t inner4 = (meta_lo & (meta_lo << 1)) & 0x10101010u;   // bit 4 set <=> (meta & 0x18) == 0x18
            const uint32_t index4 = UNORDERED ? (meta_lo & 0x1f1f1f1fu) : (meta_lo ^ (octinv4 & ((inner4 >> 4) * 0x07u))) & 0x1f1f1f1fu;
            const uint32_t bits4 = (meta_lo >> 5) & 0x07070707u;
            child<0>(bits4, index4, nearx0, farx0, neary0, fary0, nearz0, farz0, a15, cn, af, cf, tf_cap, hitmask);
            child<1>(bits4, index4, nearx0, farx0, neary0, fary0, nearz0, farz0, a15, cn, af, cf, tf_cap, hitmask);
            child<2>(bits4, index4, nearx0, farx0, neary0, fary0, nearz0, farz0, a15, cn, af, cf, tf_cap, hitmask);
            child<3>(bits4, index4, nearx0, farx0, neary0, fary0, nearz0, farz0, a15, cn, af, cf, tf_cap, hitmask);
        }
        {
            const uint32_t inner4 = (meta_hi & (meta_hi << 1)) & 0x10101010u;
            const uint32_t index4 = UNORDERED ? (meta_hi & 0x1f1f1f1fu) : (meta_hi ^ (octinv4 & ((inner4 >> 4) * 0x07u))) & 0x1f1f1f1fu;
            const uint32_t bits4 = (meta_hi >> 5) & 0x07070707u;
            child<0>(bits4, index4, nearx1, farx1, neary1, fary1, nearz1, farz1, a15, cn, af, cf, tf_cap, hitmask);
            child<1>(bits4, index4, nearx1, farx1, neary1, fary1, nearz1, farz1, a15, cn, af, cf, tf_cap, hitmask);
            child<2>(bits4, index4, nearx1, farx1, neary1, fary1, nearz1, farz1, a15, cn, af, cf, tf_cap, hitmask);
            child<3>(bits4, index4, nearx1, farx1, neary1, fary1, nearz1, farz1, a15, cn, af, cf, tf_cap, hitmask);
        }
#endif
        ngroup.x = f2u(n1.x);
        ngroup.y = (hitmask & 0xff000000u) | imask;
        tgroup.x = f2u(n1.y);
        tgroup.y = hitmask & 0x00ffffffu;
    }

    // precondition: has_tris(). Intersects the lowest-numbered pending primitive of the current group.
    RT_HD void tri_step(const SceneD& sc, TraverseStats* stats) {
        const int b = 31 - clz32(tgroup.y & (0u - tgroup.y));  // lowest set bit
        tgroup.y &= tgroup.y - 1u;
        const uint32_t pi = tgroup.x + (uint32_t)b;
        RT_CHECK(pi < sc.prim_count);
#ifdef RT_TRACE_PRIM_HOOK   // CPU harness only (tests/hostsim): which primitives a ray tests
        RT_TRACE_PRIM_HOOK(pi);
#endif
        const Prim* pr = sc.prims + pi;
        const float4 pa = ldg(&pr->a), pb = ldg(&pr->b), pc = ldg(&pr->c);
        if (STATS) stats->prims++;
        float t, u = 0.0f, v = 0.0f;
        bool h;
        if (f2u(pc.w) == 0u) {
            h = sc.watertight ? triangle_watertight(xyz(pa), xyz(pb), xyz(pc), o, d, t_min, closest, t, u, v)
                              : triangle_t(xyz(pa), xyz(pb), xyz(pc), o, d, t_min, closest, t, u, v);
        } else {
            const Instance* inst = sc.instances + f2u(pa.w);
            V3 oo = apply_point(inst->w2o, o), od = apply_vector(inst->w2o, d);
            h = sphere_t(xyz(pa), pb.x, oo, od, t_min, closest, t);
        }
        if (ANY_HIT) {   // occlusion only: any hit in range ends the walk, no ids / barycentrics / tie-break needed
            if (h) { found = 1u; hit.prim = pi; ngroup.y = 0u; tgroup.y = 0u; sp = 0; }
            return;
        }
        const uint32_t geom = f2u(pa.w), pid = f2u(pb.w);
        if (h && (t < closest || hit.prim == NONE || geom > hit_geom || (geom == hit_geom && pid > hit_pid))) {
            hit_geom = geom;
            hit_pid = pid;
            closest = t;
            hit.t = t;
            hit.prim = pi;
            hit.u = u;
            hit.v = v;
            found = 1u;
            if (ANY_HIT) { ngroup.y = 0u; tgroup.y = 0u; sp = 0; }
        }
    }

    // make the current groups non-empty from the stack; false when nothing is left
    RT_HD bool next() {
        if (tgroup.y | (ngroup.y & 0xff000000u)) return true;
        if (sp == 0) return false;
        const uint2 g = stack[--sp];
        if (g.y & 0xff000000u) ngroup = g;
        else tgroup = g;
        return true;
    }
};

// One ray start to finish: each node's primitives right after its box test.
template <bool ANY_HIT, bool STATS>
RT_HD bool traverse(const SceneD& sc, V3 o, V3 d, float t_min, float t_max, Hit& hit, TraverseStats* stats) {
    uint2 stack_mem[TRAVERSE_STACK];
    Traversal<ANY_HIT, STATS> tr;
    tr.stack = stack_mem;
    if (tr.init(sc, o, d, t_min, t_max)) {
        do {
            if (tr.has_tris()) tr.tri_step(sc, stats);
            else tr.node_step(sc, stats);
        } while (tr.next());
    }
    hit = tr.hit;
    return tr.found != 0u;
}

}  // namespace rt
