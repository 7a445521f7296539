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

constexpr int TRAVERSE_STACK = 32;

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

RT_HD float safe_rcp_dir(float d) {
    float a = fabsf(d) > 1.0e-20f ? d : copysignf(1.0e-20f, d);
    return 1.0f / a;
}

// quantised plane byte k of word w as a float. On the device this is one PRMT building 2^23 + q in the
// mantissa plus one FADD (exact); `(float)((w >> 8k) & 0xff)` compiles to I2F.U8, which issues on the
// quarter-rate XU pipe and made the 48 conversions per node the limiter of both traversal kernels
// (profiles/r1_notes.md: XU pipe 60-105 % busy).
template <int K>
RT_HD float qbyte(uint32_t w) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(__byte_perm(w, 0x4b000000u, 0x7440u + K)) - 8388608.0f;
#else
    return (float)((w >> (8 * K)) & 0xffu);
#endif
}

// Traversal as a resumable per-ray state machine: `init` once, then `step` until it returns false. One step
// visits ONE wide node (8 quantised slab tests) and intersects the primitives of the leaf children it hit.
// The one-ray-per-thread kernels just loop (`traverse` below); the persistent kernels interleave steps with
// warp-level ray refill so lanes whose ray ended do not idle while their neighbours walk a deep subtree.
template <bool ANY_HIT, bool STATS>
struct Traversal {
    V3 o, d, idir;
    float t_min, closest;
    uint32_t oct, octinv;
    uint2 ngroup;
    uint2* stack;  // TRAVERSE_STACK entries of thread-local memory owned by the caller (keeps the scalar state in registers)
    int sp;
    Hit hit;
    bool found;

    RT_HD bool init(const SceneD& sc, V3 o_, V3 d_, float t_min_, float t_max_) {
        o = o_; d = d_; t_min = t_min_; closest = t_max_;
        hit.prim = NONE; hit.t = t_max_; hit.u = hit.v = 0.0f;
        found = false;
        sp = 0;
        idir = mk3(safe_rcp_dir(d.x), safe_rcp_dir(d.y), safe_rcp_dir(d.z));
        oct = (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u);
        octinv = 7u - oct;
        ngroup = make_uint2(0u, 0x80000000u);  // root: "child 7^octinv of a group with base 0, imask 0"
        return sc.prim_count != 0;
    }

    template <int I>
    RT_HD void child(uint32_t meta_w, uint32_t nx_w, uint32_t fx_w, uint32_t ny_w, uint32_t fy_w, uint32_t nz_w, uint32_t fz_w,
                     float ax, float ay, float az, float bx, float by, float bz, uint32_t& hitmask) const {
        const uint32_t meta = (meta_w >> (8 * I)) & 0xffu;
        if (meta == 0u) return;
        const float tnx = qbyte<I>(nx_w) * ax + bx, tfx = qbyte<I>(fx_w) * ax + bx;
        const float tny = qbyte<I>(ny_w) * ay + by, tfy = qbyte<I>(fy_w) * ay + by;
        const float tnz = qbyte<I>(nz_w) * az + bz, tfz = qbyte<I>(fz_w) * az + bz;
        const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, t_min));
        const float tf = fminf(fminf(tfx, tfy), fminf(tfz, closest)) * 1.0000004f;
        if (tn <= tf) {
            const bool inner = (meta & 0x18u) == 0x18u;
            hitmask |= (meta >> 5) << ((meta ^ (inner ? octinv : 0u)) & 0x1fu);
        }
    }

    // precondition: ngroup has node hits. Returns true while more nodes remain.
    RT_HD bool step(const SceneD& sc, TraverseStats* stats) {
        uint2 tgroup;
        {
            const uint32_t hits = ngroup.y;
            const int bit = bfind32(hits);
            ngroup.y &= ~(1u << bit);
            if (ngroup.y & 0xff000000u) stack[sp++] = ngroup;
            const uint32_t slot = ((uint32_t)bit - 24u) ^ octinv;
            const uint32_t rel = (uint32_t)popc32(hits & ~(0xffffffffu << slot) & 0xffu);
            const Node8* node = sc.nodes + (ngroup.x + rel);
            const float4 n0 = ldg(&node->n0), n1 = ldg(&node->n1), n2 = ldg(&node->n2), n3 = ldg(&node->n3), n4 = ldg(&node->n4);
            if (STATS) stats->nodes++;

            const uint32_t e = f2u(n0.w);
            const float sx = u2f((e & 0xffu) << 23), sy = u2f(((e >> 8) & 0xffu) << 23), sz = u2f(((e >> 16) & 0xffu) << 23);
            const uint32_t imask = e >> 24;
            const float ax = sx * idir.x, ay = sy * idir.y, az = sz * idir.z;
            const float bx = (n0.x - o.x) * idir.x, by = (n0.y - o.y) * idir.y, bz = (n0.z - o.z) * idir.z;
            const uint32_t meta_lo = f2u(n1.z), meta_hi = f2u(n1.w);
            // near / far plane words per axis, selected by the ray octant
            const uint32_t lox0 = f2u(n2.x), lox1 = f2u(n2.y), loy0 = f2u(n2.z), loy1 = f2u(n2.w);
            const uint32_t loz0 = f2u(n3.x), loz1 = f2u(n3.y), hix0 = f2u(n3.z), hix1 = f2u(n3.w);
            const uint32_t hiy0 = f2u(n4.x), hiy1 = f2u(n4.y), hiz0 = f2u(n4.z), hiz1 = f2u(n4.w);
            const bool nx = (oct & 1u) != 0, ny = (oct & 2u) != 0, nz = (oct & 4u) != 0;
            const uint32_t nearx0 = nx ? hix0 : lox0, nearx1 = nx ? hix1 : lox1, farx0 = nx ? lox0 : hix0, farx1 = nx ? lox1 : hix1;
            const uint32_t neary0 = ny ? hiy0 : loy0, neary1 = ny ? hiy1 : loy1, fary0 = ny ? loy0 : hiy0, fary1 = ny ? loy1 : hiy1;
            const uint32_t nearz0 = nz ? hiz0 : loz0, nearz1 = nz ? hiz1 : loz1, farz0 = nz ? loz0 : hiz0, farz1 = nz ? loz1 : hiz1;

            uint32_t hitmask = 0;
            child<0>(meta_lo, nearx0, farx0, neary0, fary0, nearz0, farz0, ax, ay, az, bx, by, bz, hitmask);
            child<1>(meta_lo, nearx0, farx0, neary0, fary0, nearz0, farz0, ax, ay, az, bx, by, bz, hitmask);
            child<2>(meta_lo, nearx0, farx0, neary0, fary0, nearz0, farz0, ax, ay, az, bx, by, bz, hitmask);
            child<3>(meta_lo, nearx0, farx0, neary0, fary0, nearz0, farz0, ax, ay, az, bx, by, bz, hitmask);
            child<0>(meta_hi, nearx1, farx1, neary1, fary1, nearz1, farz1, ax, ay, az, bx, by, bz, hitmask);
            child<1>(meta_hi, nearx1, farx1, neary1, fary1, nearz1, farz1, ax, ay, az, bx, by, bz, hitmask);
            child<2>(meta_hi, nearx1, farx1, neary1, fary1, nearz1, farz1, ax, ay, az, bx, by, bz, hitmask);
            child<3>(meta_hi, nearx1, farx1, neary1, fary1, nearz1, farz1, ax, ay, az, bx, by, bz, hitmask);
            ngroup.x = f2u(n1.x);
            ngroup.y = (hitmask & 0xff000000u) | imask;
            tgroup.x = f2u(n1.y);
            tgroup.y = hitmask & 0x00ffffffu;
        }

        while (tgroup.y) {
            const int b = 31 - clz32(tgroup.y & (0u - tgroup.y));  // lowest set bit
            tgroup.y &= tgroup.y - 1u;
            const uint32_t pi = tgroup.x + (uint32_t)b;
            const Prim* pr = sc.prims + pi;
            const float4 pa = ldg(&pr->a), pb = ldg(&pr->b), pc = ldg(&pr->c);
            if (STATS) stats->prims++;
            float t, u = 0.0f, v = 0.0f;
            bool h;
            if (f2u(pc.w) == 0u) {
                h = triangle_t(xyz(pa), xyz(pb), xyz(pc), o, d, t_min, closest, t, u, v);
            } else {
                const Instance* inst = sc.instances + f2u(pa.w);
                V3 oo = apply_point(inst->w2o, o), od = apply_vector(inst->w2o, d);
                h = sphere_t(xyz(pa), pb.x, oo, od, t_min, closest, t);
            }
            if (h) {
                closest = t;
                hit.t = t;
                hit.prim = pi;
                hit.u = u;
                hit.v = v;
                found = true;
                if (ANY_HIT) return false;
            }
        }

        if (!(ngroup.y & 0xff000000u)) {
            if (sp == 0) return false;
            ngroup = stack[--sp];
        }
        return true;
    }
};

template <bool ANY_HIT, bool STATS>
RT_HD bool traverse(const SceneD& sc, V3 o, V3 d, float t_min, float t_max, Hit& hit, TraverseStats* stats) {
    uint2 stack_mem[TRAVERSE_STACK];
    Traversal<ANY_HIT, STATS> tr;
    tr.stack = stack_mem;
    if (tr.init(sc, o, d, t_min, t_max))
        while (tr.step(sc, stats)) {}
    hit = tr.hit;
    return tr.found;
}

}  // namespace rt
