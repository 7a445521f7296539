// ORACLE — TEST INFRASTRUCTURE ONLY. NOT PART OF THE PRODUCT PATH.
//
// A scalar CPU restatement of the reference CPU renderer (crates/raytracing-cpu) of
// buggy213/opencl-raytracing, written from the reference's behaviour, used ONLY as the checker in
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
// Nothing under opencl-raytracing_b200/ may include, link or call this file.
//
// PARITY STATUS: **parity unpinned** beyond the reference's own known-answer tests. The reference
// cannot be compiled here (no Rust toolchain, no Embree) and ships no golden images or sampler
// vectors. What IS pinned (tests/test_oracle_kats.py):
//   - test_sphere_uv_off_center        crates/raytracing-cpu/src/geometry.rs:342-373
//   - test_make_orthonormal_basis      crates/raytracing-cpu/src/geometry.rs:22-47
//   - test_permute                     crates/raytracing-cpu/src/sample.rs:256-275
//   - PCG32 XSH-RR published demo vectors (pcg32-demo, seed 42 / stream 54)
//   - self-consistency: BVH2 closest hit == brute-force closest hit
//   - rustc-hash 2.x FxHasher multiplier and the rotate_left(26) of finish(): found as `imm64 K ... rol r64, 26` in Rust
//     extension modules of this image that link the crate (test_fxhasher_constants_against_compiled_rustc_hash)
// Third-party arithmetic restated from the published algorithms (crate sources are not vendored
// in /root/reference): rustc-hash 2.1.1 FxHasher, rand_pcg 0.9.0 Lcg64Xsh32, rand 0.9.2
// StandardUniform<f32> / random_range(u32), image 0.25.8 imageops::resize(Lanczos3).
// Embree's rtcBuildBVH (external, absent) is replaced by a binned-SAH BVH2 builder with the same
// build arguments (maxLeafSize 8, branching 2): only equal-t tie-breaks depend on the topology.
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off: Rust never contracts a*b+c into fma).

#include "../include/rtcuda.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace {

constexpr float PI = 3.14159265358979323846f;      // f32::consts::PI
constexpr float FRAC_1_PI = 0.318309886183790671537767526745028724f;
constexpr float FRAC_PI_2 = 1.57079632679489661923132169163975144f;
constexpr float FRAC_PI_4 = 0.785398163397448309615660845819875721f;
constexpr float INF = std::numeric_limits<float>::infinity();

// ---------------------------------------------------------------------------------------------
// Vocabulary math. crates/raytracing/src/geometry/{vec2,vec3,vec4}.rs.
// NOTE `v / s` is `v * (1.0 / s)` in the reference (vec3.rs:122-126,179-184; vec2.rs:139-144).
// ---------------------------------------------------------------------------------------------
struct Vec2 { float x = 0, y = 0; };
inline Vec2 operator+(Vec2 a, Vec2 b) { return {a.x + b.x, a.y + b.y}; }
inline Vec2 operator-(Vec2 a, Vec2 b) { return {a.x - b.x, a.y - b.y}; }
inline Vec2 operator*(Vec2 a, float s) { return {a.x * s, a.y * s}; }
inline Vec2 operator*(float s, Vec2 a) { return a * s; }
inline float sqmag(Vec2 a) { return a.x * a.x + a.y * a.y; }

struct Vec3 { float x = 0, y = 0, z = 0; };
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator-(Vec3 a) { return {-a.x, -a.y, -a.z}; }
inline Vec3 operator*(Vec3 a, Vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline Vec3 operator*(Vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3 operator*(float s, Vec3 a) { return a * s; }
inline Vec3 operator/(Vec3 a, float s) { return a * (1.0f / s); }
inline Vec3& operator+=(Vec3& a, Vec3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
inline Vec3& operator*=(Vec3& a, Vec3 b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; return a; }
inline Vec3& operator*=(Vec3& a, float s) { a.x *= s; a.y *= s; a.z *= s; return a; }
inline Vec3& operator/=(Vec3& a, float s) { a *= (1.0f / s); return a; }
inline bool operator==(Vec3 a, Vec3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline bool operator!=(Vec3 a, Vec3 b) { return !(a == b); }
inline float dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(Vec3 u, Vec3 v) {
    return {u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x};
}
inline float sqmag(Vec3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline float length(Vec3 a) { return std::sqrt(sqmag(a)); }
inline Vec3 unit(Vec3 a) { return a / length(a); }
inline float max_component(Vec3 a) { return std::fmax(a.x, std::fmax(a.y, a.z)); }
inline Vec3 vmin(Vec3 a, Vec3 b) { return {std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)}; }
inline Vec3 vmax(Vec3 a, Vec3 b) { return {std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)}; }
// vec3.rs:93-95
inline Vec3 reflect(Vec3 v, Vec3 n) { return -v + 2.0f * dot(v, n) * n; }

struct Vec4 { float x = 0, y = 0, z = 0, w = 0; };
inline Vec4 operator+(Vec4 a, Vec4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline Vec4 operator-(Vec4 a, Vec4 b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
inline Vec4 operator*(Vec4 a, Vec4 b) { return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
inline Vec4 operator*(Vec4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline Vec4 operator*(float s, Vec4 a) { return a * s; }
inline bool operator==(Vec4 a, Vec4 b) { return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w; }

// Rust f32 helpers with their exact semantics.
inline float rs_clamp(float v, float lo, float hi) {  // f32::clamp: NaN stays NaN
    if (v < lo) return lo;
    if (v > hi) return hi;
    return v;
}
inline float rs_fract(float v) { return v - std::trunc(v); }
inline float rs_signum(float v) { return std::isnan(v) ? v : std::copysign(1.0f, v); }
inline uint32_t rs_as_u32(float v) {  // saturating `as u32`
    if (!(v > 0.0f)) return 0;  // NaN, negatives
    if (v >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)v;
}
inline int32_t rs_as_i32(float v) {
    if (std::isnan(v)) return 0;
    if (v <= -2147483648.0f) return INT32_MIN;
    if (v >= 2147483648.0f) return INT32_MAX;
    return (int32_t)v;
}

// crates/raytracing/src/geometry/matrix4x4.rs:326-359
struct Mat4 {
    float m[16];
    Vec3 apply_point(Vec3 p) const {
        float a = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3] * 1.0f;
        float b = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7] * 1.0f;
        float c = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11] * 1.0f;
        float d = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15] * 1.0f;
        return {a / d, b / d, c / d};
    }
    Vec3 apply_vector(Vec3 v) const {
        return {m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z,
                m[8] * v.x + m[9] * v.y + m[10] * v.z};
    }
    Vec3 apply_vector_transposed(Vec3 v) const {
        return {m[0] * v.x + m[4] * v.y + m[8] * v.z, m[1] * v.x + m[5] * v.y + m[9] * v.z,
                m[2] * v.x + m[6] * v.y + m[10] * v.z};
    }
};
inline Mat4 mat_from(const rtcuda_mat4& s) { Mat4 r; std::memcpy(r.m, s.m, sizeof r.m); return r; }
inline Mat4 mat_identity() { Mat4 r{}; r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f; return r; }

// crates/raytracing/src/geometry/transform.rs
struct Transform {
    Mat4 forward, inverse;
    Vec3 apply_point(Vec3 p) const { return forward.apply_point(p); }
    Vec3 apply_inverse_point(Vec3 p) const { return inverse.apply_point(p); }
    Vec3 apply_vector(Vec3 v) const { return forward.apply_vector(v); }
    Vec3 apply_normal(Vec3 n) const { return inverse.apply_vector_transposed(n); }  // transform.rs:68-73
    Transform invert() const { return {inverse, forward}; }
};
inline Transform tf_from(const rtcuda_transform& t) { return {mat_from(t.forward), mat_from(t.inverse)}; }
inline Transform tf_identity() { return {mat_identity(), mat_identity()}; }

struct AABB {
    Vec3 mn{INF, INF, INF}, mx{-INF, -INF, -INF};
    Vec3 center() const { return (mx + mn) / 2.0f; }               // aabb.rs:27-29
    float radius() const { return length(mx - center()); }        // aabb.rs:31-33
    void grow(const AABB& o) { mn = vmin(mn, o.mn); mx = vmax(mx, o.mx); }
    void grow(Vec3 p) { mn = vmin(mn, p); mx = vmax(mx, p); }
    float half_area() const { Vec3 d = mx - mn; return d.x * d.y + d.y * d.z + d.z * d.x; }
};
// aabb.rs:81-95
inline AABB transform_aabb(const AABB& a, const Transform& t) {
    AABB r;
    for (int i = 0; i < 8; i++) {
        Vec3 p{(i & 4) ? a.mx.x : a.mn.x, (i & 2) ? a.mx.y : a.mn.y, (i & 1) ? a.mx.z : a.mn.z};
        r.grow(t.apply_point(p));
    }
    return r;
}

// crates/raytracing-cpu/src/ray.rs
struct Ray {
    Vec3 origin, direction;
    Vec3 at(float t) const { return origin + direction * t; }
};
inline Ray ray_transform(const Ray& r, const Transform& t) { return {t.apply_point(r.origin), t.apply_vector(r.direction)}; }
struct RayDifferentials { Vec3 x_origin, y_origin, x_direction, y_direction; };

// ---------------------------------------------------------------------------------------------
// Hashing + RNG. rustc-hash 2.1.1 FxHasher (64-bit), rand_pcg 0.9.0 Lcg64Xsh32, rand 0.9.2.
// Call sites: crates/raytracing-cpu/src/sample.rs:29-181.
// ---------------------------------------------------------------------------------------------
struct FxHasher {
    uint64_t hash = 0;
    static constexpr uint64_t K = 0xf1357aea2e62a9c5ull;
    void add(uint64_t i) { hash = (hash + i) * K; }
    void write_u32(uint32_t v) { add(v); }
    void write_u64(uint64_t v) { add(v); }
    uint64_t finish() const { return (hash << 26) | (hash >> 38); }  // rotate_left(26)
};

struct Pcg32 {
    uint64_t state = 0, inc = 1;
    static constexpr uint64_t MUL = 6364136223846793005ull;
    Pcg32() = default;
    Pcg32(uint64_t st, uint64_t stream) {
        inc = (stream << 1) | 1;
        state = st + inc;
        step();
    }
    void step() { state = state * MUL + inc; }
    uint32_t next_u32() {
        uint64_t old = state;
        step();
        uint32_t rot = (uint32_t)(old >> 59);
        uint32_t xsh = (uint32_t)(((old >> 18) ^ old) >> 27);
        return (xsh >> rot) | (xsh << ((32 - rot) & 31));
    }
    // rand 0.9.2 StandardUniform for f32: 24 random mantissa bits, [0,1)
    float next_f32() { return (float)(next_u32() >> 8) * (1.0f / 16777216.0f); }
    // rand 0.9.2 UniformInt<u32>::sample_single (Canon's method, one bias-reduction draw)
    uint32_t range_u32(uint32_t lo, uint32_t hi) {
        uint32_t range = hi - lo;  // hi exclusive; sample_single_inclusive(lo, hi-1): range = hi-1-lo+1
        if (range == 0) return next_u32();
        uint64_t m = (uint64_t)next_u32() * range;
        uint32_t result = (uint32_t)(m >> 32), lo_order = (uint32_t)m;
        if (lo_order > (uint32_t)(0u - range)) {
            uint64_t m2 = (uint64_t)next_u32() * range;
            uint32_t new_hi = (uint32_t)(m2 >> 32);
            if ((uint64_t)lo_order + new_hi > 0xffffffffull) result += 1;
        }
        return lo + result;
    }
};

// crates/raytracing-cpu/src/sample.rs:228-254
uint32_t permute(uint32_t index, uint32_t length, uint32_t seed) {
    uint32_t npot = 1;  // length.next_power_of_two()
    while (npot < length) npot <<= 1;
    uint32_t mask = npot - 1;
    for (;;) {
        index ^= seed;
        index *= 0xe170893du;
        index ^= seed >> 16;
        index ^= (index & mask) >> 4;
        index ^= seed >> 8;
        index *= 0x0929eb3fu;
        index ^= seed >> 23;
        index ^= (index & mask) >> 1;
        index *= (1u | seed >> 27);
        index *= 0x6935fa69u;
        index ^= (index & mask) >> 11;
        index *= 0x74dcb303u;
        index ^= (index & mask) >> 2;
        index *= 0x9e501cc3u;
        index ^= (index & mask) >> 2;
        index *= 0xc860a3dfu;
        index &= mask;
        index ^= index >> 5;
        if (index < length) return (index + seed) % length;  // release-mode wrapping add
    }
}

// crates/raytracing-cpu/src/sample.rs:8-181 (CpuSampler)
struct Sampler {
    bool stratified = false;
    bool jitter = true;
    uint32_t x_strata = 1, y_strata = 1;
    uint64_t seed = 0;  // already hashed
    Pcg32 rng;
    uint32_t dimension = 0, sample_index = 0;

    static Sampler from_settings(const rtcuda_settings& s) {  // sample.rs:29-57
        Sampler r;
        uint64_t seed = s.has_seed ? s.seed : 42;
        FxHasher h;
        h.write_u64(seed);
        r.seed = h.finish();
        r.stratified = s.sampler_kind == RTCUDA_SAMPLER_STRATIFIED;
        r.jitter = s.stratified_jitter != 0;
        r.x_strata = s.x_strata;
        r.y_strata = s.y_strata;
        r.rng = Pcg32(r.seed, 0);
        return r;
    }
    static Sampler one_off(uint64_t seed) {  // sample.rs:59-64 (seed NOT re-hashed)
        Sampler r;
        r.seed = seed;
        r.rng = Pcg32(seed, 0);
        return r;
    }
    void start_sample(uint32_t px, uint32_t py, uint64_t sidx) {  // sample.rs:69-87
        FxHasher h;
        h.write_u32(px);
        h.write_u32(py);
        h.write_u32((uint32_t)sidx);
        rng = Pcg32(seed, h.finish());
        dimension = 0;
        sample_index = (uint32_t)sidx;
    }
    uint32_t dim_hash() const {
        FxHasher h;
        h.write_u32(dimension);
        h.write_u64(seed);
        return (uint32_t)h.finish();
    }
    float uniform() {  // sample.rs:89-121
        if (!stratified) return rng.next_f32();
        uint32_t total = x_strata * y_strata;
        uint32_t strata = permute(sample_index, total, dim_hash());
        float delta = jitter ? rng.next_f32() : 0.5f;
        dimension += 1;
        return ((float)strata + delta) / (float)total;
    }
    uint32_t u32_range(uint32_t lo, uint32_t hi) {  // sample.rs:123-138
        if (!stratified) return rng.range_u32(lo, hi);
        float u = uniform();
        float offset = u * (float)(hi - lo);
        return lo + rs_as_u32(offset);
    }
    Vec2 uniform2() {  // sample.rs:140-180
        if (!stratified) {
            float a = rng.next_f32();
            float b = rng.next_f32();
            return {a, b};
        }
        uint32_t hash = dim_hash();
        uint32_t total = x_strata * y_strata;
        uint32_t strata = permute(sample_index, total, hash);
        dimension += 2;
        uint32_t y = strata / x_strata, x = strata % x_strata;
        float dx = 0.5f, dy = 0.5f;
        if (jitter) { dx = rng.next_f32(); dy = rng.next_f32(); }
        return {((float)x + dx) / (float)x_strata, ((float)y + dy) / (float)y_strata};
    }
};

// sample.rs:184-224
inline Vec2 sample_unit_disk(Vec2 u) {
    float r = std::sqrt(u.x);
    float theta = 2.0f * PI * u.y;
    return {r * std::cos(theta), r * std::sin(theta)};
}
inline Vec2 sample_unit_disk_concentric(Vec2 u) {
    Vec2 o = 2.0f * u - Vec2{1.0f, 1.0f};
    if (o.x == 0.0f && o.y == 0.0f) return {0, 0};
    float theta, r;
    if (std::fabs(o.x) > std::fabs(o.y)) { theta = FRAC_PI_4 * (o.y / o.x); r = o.x; }
    else { theta = FRAC_PI_2 - FRAC_PI_4 * (o.x / o.y); r = o.y; }
    return r * Vec2{std::cos(theta), std::sin(theta)};
}
inline Vec3 sample_cosine_hemisphere(Vec2 u) {
    Vec2 d = sample_unit_disk(u);
    float z = std::sqrt(std::fmax(1.0f - d.x * d.x - d.y * d.y, 0.0f));
    return {d.x, d.y, z};
}
inline float sample_exponential(float u, float a) { return -std::log(1.0f - u) / a; }
inline float power_heuristic(uint32_t na, float pa, uint32_t nb, float pb) {
    float wa = ((float)na * pa) * ((float)na * pa);
    float wb = ((float)nb * pb) * ((float)nb * pb);
    return wa / (wa + wb);
}

// ---------------------------------------------------------------------------------------------
// Geometry. crates/raytracing-cpu/src/geometry.rs
// ---------------------------------------------------------------------------------------------
// geometry.rs:8-20
inline void make_orthonormal_basis(Vec3 z, Vec3& x, Vec3& y) {
    Vec3 a = std::fabs(z.z) < 0.8f ? Vec3{0, 0, 1} : Vec3{0, 1, 0};
    x = unit(cross(a, z));
    y = cross(z, x);
}

// geometry.rs:51-78. Returns false for None; t0 may be negative.
inline bool intersect_aabb(const AABB& b, const Ray& r, float& t0, float& t1) {
    float a = (b.mn.x - r.origin.x) / r.direction.x, bb = (b.mx.x - r.origin.x) / r.direction.x;
    float t0x = std::fmin(a, bb), t1x = std::fmax(a, bb);
    float c = (b.mn.y - r.origin.y) / r.direction.y, d = (b.mx.y - r.origin.y) / r.direction.y;
    float t0y = std::fmin(c, d), t1y = std::fmax(c, d);
    float e = (b.mn.z - r.origin.z) / r.direction.z, f = (b.mx.z - r.origin.z) / r.direction.z;
    float t0z = std::fmin(e, f), t1z = std::fmax(e, f);
    t0 = std::fmax(std::fmax(t0x, t0y), t0z);
    t1 = std::fmin(std::fmin(t1x, t1y), t1z);
    return t0 <= t1;
}

struct IntersectResult { float t; Vec2 uv; Vec3 point, normal, dpdu, dpdv; };

struct MeshView {
    const float* vertices = nullptr;   // 3 f32
    const uint32_t* tris = nullptr;    // 3 u32
    const float* normals = nullptr;    // may be null
    const float* uvs = nullptr;        // may be null
    uint32_t tri_count = 0;
    Vec3 v(uint32_t i) const { return {vertices[3 * i], vertices[3 * i + 1], vertices[3 * i + 2]}; }
    Vec3 n(uint32_t i) const { return {normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]}; }
    Vec2 uv(uint32_t i) const { return {uvs[2 * i], uvs[2 * i + 1]}; }
    float tri_area(uint32_t t) const {  // mesh.rs:271-278
        Vec3 p0 = v(tris[3 * t]), p1 = v(tris[3 * t + 1]), p2 = v(tris[3 * t + 2]);
        return length(cross(p1 - p0, p2 - p0)) / 2.0f;
    }
};

// geometry.rs:301-340 (Möller–Trumbore, inclusive bounds, no culling)
inline bool ray_triangle_intersect(Vec3 p0, Vec3 p1, Vec3 p2, const Ray& r, float t_min, float t_max,
                                   float& t, float& u, float& v) {
    Vec3 e1 = p1 - p0, e2 = p2 - p0;
    Vec3 P = cross(r.direction, e2);
    float denom = dot(P, e1);
    if (denom == 0.0f) return false;
    Vec3 T = r.origin - p0;
    u = dot(P, T) / denom;
    if (u < 0.0f || u > 1.0f) return false;
    Vec3 Q = cross(T, e1);
    v = dot(Q, r.direction) / denom;
    if (v < 0.0f || u + v > 1.0f) return false;
    t = dot(Q, e2) / denom;
    if (t < t_min || t > t_max) return false;
    return true;
}

// geometry.rs:229-298
inline bool ray_mesh_intersect(const MeshView& mesh, uint32_t tri, const Ray& r, float t_min, float t_max,
                               IntersectResult& out) {
    uint32_t i0 = mesh.tris[3 * tri], i1 = mesh.tris[3 * tri + 1], i2 = mesh.tris[3 * tri + 2];
    Vec3 p0 = mesh.v(i0), p1 = mesh.v(i1), p2 = mesh.v(i2);
    float t, u, v;
    if (!ray_triangle_intersect(p0, p1, p2, r, t_min, t_max, t, u, v)) return false;
    float w = 1.0f - u - v;
    Vec3 n;
    if (!mesh.normals) n = unit(cross(p2 - p0, p1 - p0));
    else n = unit(w * mesh.n(i0) + u * mesh.n(i1) + v * mesh.n(i2));
    Vec2 uv0{0, 0}, uv1{1, 0}, uv2{0, 1};
    if (mesh.uvs) { uv0 = mesh.uv(i0); uv1 = mesh.uv(i1); uv2 = mesh.uv(i2); }
    Vec2 uv = w * uv0 + u * uv1 + v * uv2;
    Vec2 duv02 = uv0 - uv2, duv12 = uv1 - uv2;
    Vec3 dp02 = p0 - p2, dp12 = p1 - p2;
    float det = duv02.x * duv12.y - duv02.y * duv12.x;
    Vec3 dpdu{}, dpdv{};
    if (!(std::fabs(det) < 1.0e-9f)) {
        float inv_det = 1.0f / det;
        dpdu = inv_det * (duv12.y * dp02 - duv02.y * dp12);
        dpdv = inv_det * (duv02.x * dp12 - duv12.x * dp02);
    }
    out = {t, uv, r.at(t), n, dpdu, dpdv};
    return true;
}

// geometry.rs:139-227
inline bool ray_sphere_intersect(Vec3 center, float radius, const Ray& r, float t_min, float t_max,
                                 IntersectResult& out) {
    Vec3 omc = r.origin - center;
    float a = sqmag(r.direction);
    float b = 2.0f * dot(r.direction, omc);
    float c = sqmag(omc) - radius * radius;
    float disc = b * b - 4.0f * a * c;
    float t1, t2;
    if (disc < 0.0f) return false;
    else if (disc == 0.0f) { float t = -b / (2.0f * a); t1 = t2 = t; }
    else {
        float q = -0.5f * (b + rs_signum(b) * std::sqrt(disc));
        t1 = q / a;
        t2 = c / q;
    }
    if (t1 > t2) std::swap(t1, t2);
    float t;
    if (t1 >= t_min && t1 <= t_max) t = t1;
    else if (t2 >= t_min && t2 <= t_max) t = t2;
    else return false;
    Vec3 point = r.at(t);
    Vec3 local = point - center;
    float theta = std::acos(local.z / radius);
    float cos_phi = local.x / (radius * std::sin(theta));
    float sin_phi = local.y / (radius * std::sin(theta));
    float phi = local.y > 0.0f ? std::acos(cos_phi) : 2.0f * PI - std::acos(cos_phi);
    Vec2 uv{phi / (2.0f * PI), theta / PI};
    Vec3 dpdu{-2.0f * PI * local.y, 2.0f * PI * local.x, 0.0f};
    float sin_theta = std::sin(theta);
    Vec3 dpdv = PI * Vec3{local.z * cos_phi, local.z * sin_phi, -radius * sin_theta};
    out = {t, uv, point, local / radius, dpdu, dpdv};
    return true;
}

// ---------------------------------------------------------------------------------------------
// Scene view over the flat descriptor (mirror of crates/raytracing/src/scene/scene.rs accessors).
// ---------------------------------------------------------------------------------------------
struct ImageView {
    uint32_t w = 0, h = 0, channels = 0, format = 0;
    const uint8_t* data = nullptr;
    std::vector<uint8_t> owned;  // for generated mips
    // crates/raytracing/src/materials/image.rs:56-121
    float channel(uint32_t x, uint32_t y, uint32_t c) const {
        if (c >= channels) return 0.0f;
        size_t idx = ((size_t)y * w + x) * channels + c;
        switch (format) {
            case RTCUDA_IMAGE_U8: return (float)data[idx] / 255.0f;
            case RTCUDA_IMAGE_U16: return (float)((const uint16_t*)data)[idx] / 65535.0f;
            default: return ((const float*)data)[idx] / 1.0f;
        }
    }
    Vec4 pixel(uint32_t x, uint32_t y) const { return {channel(x, y, 0), channel(x, y, 1), channel(x, y, 2), channel(x, y, 3)}; }
};

struct Mipmap { ImageView mip0; std::vector<ImageView> mips; };

// image 0.25.8 imageops::resize(.., FilterType::Lanczos3): vertical pass into f32 (no clamp), then
// horizontal pass clamped to the f32 pixel range [0,1]. Works on interleaved f32 with `ch` channels.
inline float lanczos3(float x) {
    auto sinc = [](float t) { float a = t * PI; return t == 0.0f ? 1.0f : std::sin(a) / a; };
    return std::fabs(x) < 3.0f ? sinc(x) * sinc(x / 3.0f) : 0.0f;
}
std::vector<float> resize_lanczos3(const std::vector<float>& src, uint32_t w, uint32_t h, uint32_t ch, uint32_t nw, uint32_t nh) {
    if (nw == w && nh == h) return src;
    std::vector<float> tmp((size_t)w * nh * ch);
    std::vector<float> ws;
    {
        float ratio = (float)h / (float)nh;
        float sratio = ratio < 1.0f ? 1.0f : ratio;
        float support = 3.0f * sratio;
        for (uint32_t oy = 0; oy < nh; oy++) {
            float in = ((float)oy + 0.5f) * ratio;
            int64_t left = (int64_t)std::floor(in - support);
            left = std::min<int64_t>(std::max<int64_t>(left, 0), (int64_t)h - 1);
            int64_t right = (int64_t)std::ceil(in + support);
            right = std::min<int64_t>(std::max<int64_t>(right, left + 1), (int64_t)h);
            in = in - 0.5f;
            ws.clear();
            float sum = 0.0f;
            for (int64_t i = left; i < right; i++) { float wgt = lanczos3(((float)i - in) / sratio); ws.push_back(wgt); sum += wgt; }
            for (auto& x : ws) x /= sum;
            for (uint32_t x = 0; x < w; x++)
                for (uint32_t c = 0; c < ch; c++) {
                    float t = 0.0f;
                    for (size_t i = 0; i < ws.size(); i++) t += src[((size_t)(left + i) * w + x) * ch + c] * ws[i];
                    tmp[((size_t)oy * w + x) * ch + c] = t;
                }
        }
    }
    std::vector<float> out((size_t)nw * nh * ch);
    {
        float ratio = (float)w / (float)nw;
        float sratio = ratio < 1.0f ? 1.0f : ratio;
        float support = 3.0f * sratio;
        for (uint32_t ox = 0; ox < nw; ox++) {
            float in = ((float)ox + 0.5f) * ratio;
            int64_t left = (int64_t)std::floor(in - support);
            left = std::min<int64_t>(std::max<int64_t>(left, 0), (int64_t)w - 1);
            int64_t right = (int64_t)std::ceil(in + support);
            right = std::min<int64_t>(std::max<int64_t>(right, left + 1), (int64_t)w);
            in = in - 0.5f;
            ws.clear();
            float sum = 0.0f;
            for (int64_t i = left; i < right; i++) { float wgt = lanczos3(((float)i - in) / sratio); ws.push_back(wgt); sum += wgt; }
            for (auto& x : ws) x /= sum;
            for (uint32_t y = 0; y < nh; y++)
                for (uint32_t c = 0; c < ch; c++) {
                    float t = 0.0f;
                    for (size_t i = 0; i < ws.size(); i++) t += tmp[((size_t)y * w + (left + i)) * ch + c] * ws[i];
                    out[((size_t)y * nw + ox) * ch + c] = rs_clamp(t, 0.0f, 1.0f);
                }
        }
    }
    return out;
}

// crates/raytracing-cpu/src/texture.rs:86-111 (`cast`): f32 -> original sample type, round & clamp.
ImageView cast_image(const std::vector<float>& f, uint32_t w, uint32_t h, uint32_t ch, uint32_t format) {
    ImageView v;
    v.w = w; v.h = h; v.channels = ch; v.format = format;
    size_t n = (size_t)w * h * ch;
    if (format == RTCUDA_IMAGE_U8) {
        v.owned.resize(n);
        for (size_t i = 0; i < n; i++) v.owned[i] = (uint8_t)std::round(rs_clamp(f[i], 0.0f, 1.0f) * 255.0f);
    } else if (format == RTCUDA_IMAGE_U16) {
        v.owned.resize(n * 2);
        for (size_t i = 0; i < n; i++) ((uint16_t*)v.owned.data())[i] = (uint16_t)std::round(rs_clamp(f[i], 0.0f, 1.0f) * 65535.0f);
    } else {
        v.owned.resize(n * 4);
        std::memcpy(v.owned.data(), f.data(), n * 4);
    }
    v.data = v.owned.data();
    return v;
}

// crates/raytracing-cpu/src/texture.rs:114-165
Mipmap generate_mips(const ImageView& base) {
    uint32_t ch = base.channels;
    std::vector<float> cur((size_t)base.w * base.h * ch);
    for (uint32_t y = 0; y < base.h; y++)
        for (uint32_t x = 0; x < base.w; x++)
            for (uint32_t c = 0; c < ch; c++) cur[((size_t)y * base.w + x) * ch + c] = base.channel(x, y, c);
    uint32_t w = base.w, h = base.h;
    auto is_pot = [](uint32_t v) { return v && !(v & (v - 1)); };
    auto npot = [](uint32_t v) { uint32_t p = 1; while (p < v) p <<= 1; return p; };
    if (!(is_pot(w) && is_pot(h)) || w != h) {
        uint32_t s = std::max(npot(w), npot(h));
        cur = resize_lanczos3(cur, w, h, ch, s, s);
        w = h = s;
    }
    Mipmap m;
    m.mip0 = cast_image(cur, w, h, ch, base.format);
    while (w > 1 && h > 1) {
        std::vector<float> next = resize_lanczos3(cur, w, h, ch, w / 2, h / 2);
        w /= 2; h /= 2;
        m.mips.push_back(cast_image(next, w, h, ch, base.format));
        cur.swap(next);
    }
    // fix up data pointers after vector moves
    m.mip0.data = m.mip0.owned.data();
    for (auto& i : m.mips) i.data = i.owned.data();
    return m;
}

// crates/raytracing-cpu/src/materials.rs:702-796
struct MaterialEvalContext { Vec2 uv; float dudx = 0, dudy = 0, dvdx = 0, dvdy = 0; };

struct HitInfo {  // crates/raytracing-cpu/src/accel.rs:13-25
    float t = 0;
    Vec2 uv;
    Vec3 point, normal, dpdu, dpdv;
    uint32_t material_idx = 0;
    uint32_t light_idx = RTCUDA_NONE;
    uint32_t geom_id = RTCUDA_NONE, prim_id = RTCUDA_NONE;  // debug planes only
};

inline MaterialEvalContext mec_new(const HitInfo& hit, Vec3 dpdx, Vec3 dpdy) {  // materials.rs:715-763
    Vec3 dpdu = hit.dpdu, dpdv = hit.dpdv;
    float ata00 = dot(dpdu, dpdu), ata11 = dot(dpdv, dpdv), ata01 = dot(dpdu, dpdv);
    float det = ata00 * ata11 - ata01 * ata01;
    float inv_det = 1.0f / det;
    float atb0x = dot(dpdu, dpdx), atb1x = dot(dpdv, dpdx), atb0y = dot(dpdu, dpdy), atb1y = dot(dpdv, dpdy);
    float dudx = inv_det * (ata11 * atb0x - ata01 * atb1x);
    float dvdx = inv_det * (ata00 * atb1x - ata01 * atb0x);
    float dudy = inv_det * (ata11 * atb0y - ata01 * atb1y);
    float dvdy = inv_det * (ata00 * atb1y - ata01 * atb0y);
    auto cl = [](float v) { return std::isfinite(v) ? rs_clamp(v, -1.0e8f, 1.0e8f) : 0.0f; };
    return {hit.uv, cl(dudx), cl(dudy), cl(dvdx), cl(dvdy)};
}
inline MaterialEvalContext mec_no_aa(Vec2 uv) { return {uv, 0, 0, 0, 0}; }
inline MaterialEvalContext mec_from_differentials(const HitInfo& hit, const Ray& ray, const RayDifferentials& rd) {  // :771-796
    Vec3 n = hit.normal, p = hit.point;
    Vec3 rx_o = ray.origin + rd.x_origin, rx_d = ray.direction + rd.x_direction;
    Vec3 ry_o = ray.origin + rd.y_origin, ry_d = ray.direction + rd.y_direction;
    float d = -dot(n, p);
    float tx = -(dot(n, rx_o) + d) / dot(n, rx_d);
    float ty = -(dot(n, ry_o) + d) / dot(n, ry_d);
    Vec3 px = rx_o + tx * rx_d, py = ry_o + ty * ry_d;
    return mec_new(hit, px - hit.point, py - hit.point);
}

// crates/raytracing/src/materials/texture.rs:45-68
inline float wrap_apply(uint32_t mode, float x) {
    switch (mode) {
        case RTCUDA_WRAP_REPEAT: { float f = rs_fract(x); return f < 0.0f ? 1.0f + f : f; }
        case RTCUDA_WRAP_MIRROR: {
            float f = rs_fract(x);
            float rep = f < 0.0f ? 1.0f + f : f;
            int32_t fl = rs_as_i32(std::floor(x));
            int32_t r = fl % 2; if (r < 0) r += 2;  // rem_euclid
            return r == 1 ? 1.0f - rep : rep;
        }
        default: return rs_clamp(x, 0.0f, 1.0f);
    }
}

struct SceneView {
    const rtcuda_scene_desc* d = nullptr;
    std::vector<MeshView> meshes;           // per shape
    std::vector<Transform> inst_tf;         // per instance
    std::vector<ImageView> images;
    std::vector<std::unique_ptr<Mipmap>> mipmaps;  // per image, null unless trilinear-referenced

    explicit SceneView(const rtcuda_scene_desc* desc) : d(desc) {
        meshes.resize(d->shape_count);
        for (uint32_t i = 0; i < d->shape_count; i++) {
            const rtcuda_shape& s = d->shapes[i];
            if (s.kind != RTCUDA_SHAPE_TRIANGLE_MESH) continue;
            MeshView& m = meshes[i];
            m.vertices = d->vertices + 3 * (size_t)s.vertex_offset;
            m.tris = d->tris + 3 * (size_t)s.tri_offset;
            m.tri_count = s.tri_count;
            m.normals = s.normal_offset == RTCUDA_NONE ? nullptr : d->normals + 3 * (size_t)s.normal_offset;
            m.uvs = s.uv_offset == RTCUDA_NONE ? nullptr : d->uvs + 2 * (size_t)s.uv_offset;
        }
        inst_tf.resize(d->instance_count);
        for (uint32_t i = 0; i < d->instance_count; i++) inst_tf[i] = tf_from(d->instances[i].object_to_world);
        images.resize(d->image_count);
        for (uint32_t i = 0; i < d->image_count; i++) {
            const rtcuda_image& im = d->images[i];
            images[i].w = im.width; images[i].h = im.height; images[i].channels = im.channels; images[i].format = im.format;
            images[i].data = d->image_bytes + im.byte_offset;
        }
        // CpuTextures::new, crates/raytracing-cpu/src/texture.rs:214-233
        mipmaps.resize(d->image_count);
        for (uint32_t t = 0; t < d->texture_count; t++) {
            const rtcuda_texture& tx = d->textures[t];
            if (tx.kind == RTCUDA_TEXTURE_IMAGE && tx.filter == RTCUDA_FILTER_TRILINEAR && !mipmaps[tx.image])
                mipmaps[tx.image] = std::make_unique<Mipmap>(generate_mips(images[tx.image]));
        }
    }

    // texture.rs:235-268
    static Vec4 point_sample(const ImageView& im, float u, float v) {
        float w = (float)im.w, h = (float)im.h;
        float x = u * w - 0.5f, y = v * h - 0.5f;
        uint32_t xi = rs_as_u32(rs_clamp(std::round(x), 0.0f, w - 1.0f));
        uint32_t yi = rs_as_u32(rs_clamp(std::round(y), 0.0f, h - 1.0f));
        return im.pixel(xi, yi);
    }
    static Vec4 bilerp_sample(const ImageView& im, float u, float v) {
        float w = (float)im.w, h = (float)im.h;
        float x = u * w - 0.5f, y = v * h - 0.5f;
        uint32_t x0 = rs_as_u32(rs_clamp(std::floor(x), 0.0f, w - 1.0f));
        uint32_t x1 = rs_as_u32(rs_clamp(std::ceil(x), 0.0f, w - 1.0f));
        uint32_t y0 = rs_as_u32(rs_clamp(std::floor(y), 0.0f, h - 1.0f));
        uint32_t y1 = rs_as_u32(rs_clamp(std::ceil(y), 0.0f, h - 1.0f));
        float xf = rs_clamp(rs_fract(x), 0.0f, 1.0f), yf = rs_clamp(rs_fract(y), 0.0f, 1.0f);
        Vec4 p00 = im.pixel(x0, y0), p01 = im.pixel(x1, y0), p10 = im.pixel(x0, y1), p11 = im.pixel(x1, y1);
        Vec4 u0 = p00 * (1.0f - xf) + p01 * xf;
        Vec4 u1 = p10 * (1.0f - xf) + p11 * xf;
        return u0 * (1.0f - yf) + u1 * yf;
    }
    // texture.rs:270-297; returns false for None
    static bool mip_level(const ImageView& mip0, const MaterialEvalContext& c, float& level) {
        float dx = std::sqrt(c.dudx * c.dudx + c.dvdx * c.dvdx);
        float dy = std::sqrt(c.dudy * c.dudy + c.dvdy * c.dvdy);
        float larger = std::fmax(dx, dy);
        if (larger <= 0.0f) return false;
        float half_pixel = 1.0f / (2.0f * (float)mip0.w);
        level = std::log2(larger / half_pixel);
        return true;
    }
    // texture.rs:299-357
    Vec4 sample_image_texture(const rtcuda_texture& tx, const MaterialEvalContext& c) const {
        const ImageView& im = images[tx.image];
        float u = wrap_apply(tx.wrap, c.uv.x), v = wrap_apply(tx.wrap, c.uv.y);
        switch (tx.filter) {
            case RTCUDA_FILTER_NEAREST: return point_sample(im, u, v);
            case RTCUDA_FILTER_BILINEAR: return bilerp_sample(im, u, v);
            default: {
                const Mipmap& mm = *mipmaps[tx.image];
                float level;
                if (!mip_level(mm.mip0, c, level)) return bilerp_sample(im, u, v);
                float maxl = (float)mm.mips.size();
                uint32_t lower = rs_as_u32(std::floor(rs_clamp(level, 0.0f, maxl)));
                uint32_t upper = rs_as_u32(std::ceil(rs_clamp(level, 0.0f, maxl)));
                const ImageView& lo = lower == 0 ? mm.mip0 : mm.mips[lower - 1];
                const ImageView& up = upper == 0 ? mm.mip0 : mm.mips[upper - 1];
                float t = rs_fract(level);
                Vec4 a = bilerp_sample(lo, u, v), b = bilerp_sample(up, u, v);
                return t * b + (1.0f - t) * a;
            }
        }
    }
    // texture.rs:359-459
    Vec4 sample(uint32_t tex_id, const MaterialEvalContext& c) const {
        const rtcuda_texture& tx = d->textures[tex_id];
        float u = c.uv.x, v = c.uv.y;
        switch (tx.kind) {
            case RTCUDA_TEXTURE_IMAGE: return sample_image_texture(tx, c);
            case RTCUDA_TEXTURE_CONSTANT: return {tx.value[0], tx.value[1], tx.value[2], tx.value[3]};
            case RTCUDA_TEXTURE_CHECKER: {
                Vec4 color1{tx.value[0], tx.value[1], tx.value[2], tx.value[3]};
                Vec4 color2{tx.value2[0], tx.value2[1], tx.value2[2], tx.value2[3]};
                u = u - std::floor(u);
                v = v - std::floor(v);
                if ((c.dudx == 0.0f && c.dvdx == 0.0f) || (c.dudy == 0.0f && c.dvdy == 0.0f))
                    return ((u > 0.5f) != (v > 0.5f)) ? color1 : color2;
                float srx = std::sqrt(c.dudx * c.dudx + c.dvdx * c.dvdx);
                float sry = std::sqrt(c.dudy * c.dudy + c.dvdy * c.dvdy);
                float sigma = 0.1f * std::fmax(srx, sry);
                float a = u < 0.25f ? u : (u < 0.75f ? -(u - 0.5f) : u - 1.0f);
                float b = v < 0.25f ? v : (v < 0.75f ? -(v - 0.5f) : v - 1.0f);
                float xz = a / (std::sqrt(2.0f) * sigma), yz = b / (std::sqrt(2.0f) * sigma);
                float xf = 0.5f * (1.0f + std::erf(xz)), yf = 0.5f * (1.0f + std::erf(yz));
                xf = v > 0.5f ? xf : 1.0f - xf;
                yf = u > 0.5f ? yf : 1.0f - yf;
                float factor = xf * yf;
                return factor * color1 + (1.0f - factor) * color2;
            }
            case RTCUDA_TEXTURE_SCALE: return sample(tx.a, c) * sample(tx.b, c);
            default: {  // MIX
                Vec4 one{1, 1, 1, 1}, zero{0, 0, 0, 0};
                Vec4 cv = sample(tx.c, c);
                Vec4 bv = (cv == zero) ? zero : sample(tx.b, c);
                Vec4 av = (cv == one) ? zero : sample(tx.a, c);
                return (one - cv) * av + cv * bv;
            }
        }
    }
    // texture.rs:461-480
    bool texture_mip_level(uint32_t tex_id, const MaterialEvalContext& c, float& level) const {
        const rtcuda_texture& tx = d->textures[tex_id];
        if (tx.kind == RTCUDA_TEXTURE_IMAGE && tx.filter == RTCUDA_FILTER_TRILINEAR && mipmaps[tx.image])
            return mip_level(mipmaps[tx.image]->mip0, c, level);
        return false;
    }
};

// geometry.rs:92-136
inline bool intersect_shape(const SceneView& sv, const Ray& w_ray, float t_min, float t_max, const Transform& o2w,
                            uint32_t shape_idx, uint32_t prim_index, IntersectResult& out) {
    Transform w2o = o2w.invert();
    Ray o_ray = ray_transform(w_ray, w2o);
    const rtcuda_shape& s = sv.d->shapes[shape_idx];
    IntersectResult o;
    bool hit = s.kind == RTCUDA_SHAPE_TRIANGLE_MESH
                   ? ray_mesh_intersect(sv.meshes[shape_idx], prim_index, o_ray, t_min, t_max, o)
                   : ray_sphere_intersect({s.center[0], s.center[1], s.center[2]}, s.radius, o_ray, t_min, t_max, o);
    if (!hit) return false;
    out.t = o.t;
    out.uv = o.uv;
    out.point = o2w.apply_point(o.point);
    out.normal = unit(o2w.apply_normal(o.normal));
    out.dpdu = o2w.apply_vector(o.dpdu);
    out.dpdv = o2w.apply_vector(o.dpdv);
    return true;
}

// ---------------------------------------------------------------------------------------------
// BVH2: build-primitive conventions of crates/raytracing/src/accel/bvh2.rs:174-275, a binned-SAH
// builder standing in for Embree's rtcBuildBVH (embree4/src/bvh.rs:204-227: SAH, maxLeaf 8,
// branching 2, traversal/intersection cost 1), the pre-order layout of bvh2.rs:403-535 and the
// traversal of crates/raytracing-cpu/src/accel.rs:65-258.
// ---------------------------------------------------------------------------------------------
struct BuildPrim { AABB box; uint32_t geom_id, prim_id; };
struct BvhNode {
    AABB bounds;
    uint32_t right_child = 0;     // internal
    uint32_t prim_offset = 0;     // leaf
    uint32_t prim_count = 0;      // 0 => internal
};
struct PrimPtr { uint32_t geom_id, prim_id; };

struct Bvh2 {
    std::vector<BvhNode> nodes;
    std::vector<PrimPtr> prim_ptrs;
    AABB bounds() const { return nodes[0].bounds; }

    void build(std::vector<BuildPrim>& prims) {
        nodes.clear();
        prim_ptrs.clear();
        nodes.reserve(prims.size() * 2 + 1);
        if (prims.empty()) { BvhNode n; n.bounds = AABB{{INF, INF, INF}, {-INF, -INF, -INF}}; n.prim_count = 0; nodes.push_back(n); return; }
        AABB all;
        for (auto& p : prims) all.grow(p.box);
        build_rec(prims, 0, prims.size(), all, 0);
        // linearize_bvh: root bounds = union of children; a leaf root gets infinite bounds (bvh2.rs:442-456)
        if (nodes[0].prim_count != 0) nodes[0].bounds = AABB{{-INF, -INF, -INF}, {INF, INF, INF}};
    }

    uint32_t build_rec(std::vector<BuildPrim>& prims, size_t lo, size_t hi, const AABB& box, int depth) {
        uint32_t me = (uint32_t)nodes.size();
        nodes.emplace_back();
        nodes[me].bounds = box;
        size_t n = hi - lo;
        auto make_leaf = [&]() {
            nodes[me].prim_offset = (uint32_t)prim_ptrs.size();
            nodes[me].prim_count = (uint32_t)n;
            for (size_t i = lo; i < hi; i++) prim_ptrs.push_back({prims[i].geom_id, prims[i].prim_id});
            return me;
        };
        if (n == 1 || (depth >= 31 && n <= 8)) return make_leaf();
        // binned SAH over the centroid bounds
        AABB cb;
        for (size_t i = lo; i < hi; i++) cb.grow(prims[i].box.center());
        constexpr int NB = 16;
        float best_cost = INF;
        int best_axis = -1, best_bin = -1;
        for (int axis = 0; axis < 3; axis++) {
            float cmin = (&cb.mn.x)[axis], cmax = (&cb.mx.x)[axis];
            if (!(cmax > cmin)) continue;
            AABB bb[NB];
            uint32_t cnt[NB] = {};
            float scale = (float)NB / (cmax - cmin);
            for (size_t i = lo; i < hi; i++) {
                Vec3 c = prims[i].box.center();
                int b = std::min(NB - 1, (int)(((&c.x)[axis] - cmin) * scale));
                bb[b].grow(prims[i].box);
                cnt[b]++;
            }
            float right_area[NB];
            uint32_t right_cnt[NB];
            AABB acc;
            uint32_t c = 0;
            for (int b = NB - 1; b > 0; b--) { acc.grow(bb[b]); c += cnt[b]; right_area[b] = c ? acc.half_area() : 0; right_cnt[b] = c; }
            acc = AABB();
            c = 0;
            for (int b = 0; b < NB - 1; b++) {
                acc.grow(bb[b]);
                c += cnt[b];
                if (c == 0 || right_cnt[b + 1] == 0) continue;
                float cost = acc.half_area() * (float)c + right_area[b + 1] * (float)right_cnt[b + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = b; }
            }
        }
        size_t mid;
        if (best_axis < 0) {
            if (n <= 8 && depth > 0) return make_leaf();
            mid = lo + n / 2;  // all centroids coincide: median split
        } else {
            // SAH termination: leaf cost (intersection cost 1 per prim) vs split cost (traversal cost 1)
            float leaf_cost = (float)n * box.half_area();
            float split_cost = 1.0f * box.half_area() + best_cost;
            // (the root is only ever a leaf for a single primitive: Embree's own decision for 2..8
            //  primitives is unpinned, and a leaf root changes scene bounds to infinity, bvh2.rs:448-452)
            if (depth > 0 && n <= 8 && leaf_cost <= split_cost) return make_leaf();
            float cmin = (&cb.mn.x)[best_axis], cmax = (&cb.mx.x)[best_axis];
            float scale = (float)NB / (cmax - cmin);
            auto it = std::partition(prims.begin() + lo, prims.begin() + hi, [&](const BuildPrim& p) {
                Vec3 c = p.box.center();
                int b = std::min(NB - 1, (int)(((&c.x)[best_axis] - cmin) * scale));
                return b <= best_bin;
            });
            mid = it - prims.begin();
            if (mid == lo || mid == hi) mid = lo + n / 2;
        }
        AABB lb, rb;
        for (size_t i = lo; i < mid; i++) lb.grow(prims[i].box);
        for (size_t i = mid; i < hi; i++) rb.grow(prims[i].box);
        build_rec(prims, lo, mid, lb, depth + 1);  // left child = me + 1
        uint32_t r = build_rec(prims, mid, hi, rb, depth + 1);
        nodes[me].right_child = r;
        return me;
    }
};

struct Bsdf;
struct Context {
    const SceneView* sv = nullptr;
    Bvh2 bvh;
    AABB scene_bounds;
    Vec3 scene_center;
    float scene_radius = 0;
    bool brute_force = false;
    mutable std::atomic<uint64_t> n_primary{0}, n_bounce{0}, n_shadow{0}, n_aov{0};
};

// prepare_cpu_acceleration_structures, crates/raytracing-cpu/src/scene.rs:14-73 (single root aggregate)
void build_context(Context& ctx, const SceneView& sv, bool brute_force) {
    ctx.sv = &sv;
    ctx.brute_force = brute_force;
    std::vector<BuildPrim> prims;
    for (uint32_t g = 0; g < sv.d->instance_count; g++) {
        const rtcuda_instance& inst = sv.d->instances[g];
        const rtcuda_shape& s = sv.d->shapes[inst.shape];
        const Transform& tf = sv.inst_tf[g];
        if (s.kind == RTCUDA_SHAPE_TRIANGLE_MESH) {  // bvh2.rs:238-246
            const MeshView& m = sv.meshes[inst.shape];
            for (uint32_t t = 0; t < m.tri_count; t++) {
                AABB b;
                b.grow(tf.apply_point(m.v(m.tris[3 * t])));
                b.grow(tf.apply_point(m.v(m.tris[3 * t + 1])));
                b.grow(tf.apply_point(m.v(m.tris[3 * t + 2])));
                prims.push_back({b, g, t});
            }
        } else {  // bvh2.rs:208-236
            Vec3 c{s.center[0], s.center[1], s.center[2]}, r{s.radius, s.radius, s.radius};
            AABB b{c - r, c + r};
            prims.push_back({transform_aabb(b, tf), g, 0});
        }
    }
    ctx.bvh.build(prims);
    // CpuRaytracingContext::new, crates/raytracing-cpu/src/lib.rs:81-105
    ctx.scene_bounds = ctx.bvh.bounds();
    ctx.scene_center = ctx.scene_bounds.center();
    ctx.scene_radius = ctx.scene_bounds.radius();
}

inline void fill_hit(const Context& ctx, const IntersectResult& ir, uint32_t geom_id, uint32_t prim_id, HitInfo& h) {
    // accel.rs:144-164: root BVH => local_to_root is the identity, applied anyway (normal re-normalised)
    static const Transform ident = tf_identity();
    const rtcuda_shape& s = ctx.sv->d->shapes[ctx.sv->d->instances[geom_id].shape];
    h.t = ir.t;
    h.uv = ir.uv;
    h.point = ident.apply_point(ir.point);
    h.normal = unit(ident.apply_normal(ir.normal));
    h.dpdu = ident.apply_vector(ir.dpdu);
    h.dpdv = ident.apply_vector(ir.dpdv);
    h.material_idx = s.material;
    h.light_idx = s.area_light;
    h.geom_id = geom_id;
    h.prim_id = prim_id;
}

// accel.rs:65-258
bool traverse_bvh(const Context& ctx, const Ray& ray, float t_min, float t_max, bool early_exit, HitInfo& hit) {
    const SceneView& sv = *ctx.sv;
    float closest_t = t_max;
    bool found = false;
    float t0, t1;
    if (ctx.bvh.prim_ptrs.empty()) return false;
    if (!intersect_aabb(ctx.bvh.bounds(), ray, t0, t1)) return false;
    if (ctx.brute_force) {
        // same acceptance rule, every primitive tested in build order (self-check mode)
        for (const PrimPtr& pp : ctx.bvh.prim_ptrs) {
            IntersectResult ir;
            if (intersect_shape(sv, ray, t_min, closest_t, sv.inst_tf[pp.geom_id], sv.d->instances[pp.geom_id].shape, pp.prim_id, ir)) {
                fill_hit(ctx, ir, pp.geom_id, pp.prim_id, hit);
                closest_t = ir.t;
                found = true;
                if (early_exit) return true;
            }
        }
        return found;
    }
    struct Entry { uint32_t node, progress; };
    Entry stack[128];
    int sp = 0;
    stack[sp++] = {0, 0};
    const auto& nodes = ctx.bvh.nodes;
    while (sp > 0) {
        Entry& top = stack[sp - 1];
        const BvhNode& node = nodes[top.node];
        if (node.prim_count != 0) {
            const PrimPtr& pp = ctx.bvh.prim_ptrs[node.prim_offset + top.progress];
            IntersectResult ir;
            if (intersect_shape(sv, ray, t_min, closest_t, sv.inst_tf[pp.geom_id], sv.d->instances[pp.geom_id].shape, pp.prim_id, ir)) {
                fill_hit(ctx, ir, pp.geom_id, pp.prim_id, hit);
                closest_t = ir.t;
                found = true;
                if (early_exit) return true;
            }
            top.progress += 1;
            if (top.progress == node.prim_count) sp--;
        } else {
            if (top.progress == 0) {
                uint32_t left = top.node + 1;
                top.progress += 1;
                if (intersect_aabb(nodes[left].bounds, ray, t0, t1) && t0 < closest_t) stack[sp++] = {left, 0};
            } else {
                uint32_t right = node.right_child;
                sp--;
                if (intersect_aabb(nodes[right].bounds, ray, t0, t1) && t0 < closest_t) stack[sp++] = {right, 0};
            }
        }
    }
    return found;
}

// ---------------------------------------------------------------------------------------------
// BSDFs. crates/raytracing-cpu/src/materials.rs
// ---------------------------------------------------------------------------------------------
enum : uint8_t {  // materials.rs:90-103
    NONSPEC_REFL = 1, SPEC_REFL = 2, NONSPEC_TRANS = 4, SPEC_TRANS = 8,
    REFLECTION = NONSPEC_REFL | SPEC_REFL, TRANSMISSION = NONSPEC_TRANS | SPEC_TRANS,
    SPECULAR = SPEC_REFL | SPEC_TRANS, NONSPECULAR = NONSPEC_REFL | NONSPEC_TRANS, ALL = 15
};
struct BsdfSample { Vec3 wi, bsdf; float pdf = 0; uint8_t component = 0; };
enum SampleStatus { VALID, NULL_SAMPLE, INVALID };

inline SampleStatus validate(const BsdfSample& s) {  // materials.rs:28-39
    auto bad = [](float f) { return std::isinf(f) || std::isnan(f); };
    auto badv = [&](Vec3 v) { return bad(v.x) || bad(v.y) || bad(v.z); };
    if (badv(s.bsdf) || bad(s.pdf) || s.pdf <= 0.0f || badv(s.wi) || __builtin_popcount(s.component) != 1) return INVALID;
    return VALID;
}

struct Complex { float re, im; };
inline Complex operator*(Complex a, Complex b) { return {a.re * b.re - a.im * b.im, a.im * b.re + a.re * b.im}; }
inline Complex operator*(Complex a, float s) { return {a.re * s, a.im * s}; }
inline Complex operator+(Complex a, Complex b) { return {a.re + b.re, a.im + b.im}; }
inline Complex operator-(Complex a, Complex b) { return {a.re - b.re, a.im - b.im}; }
inline Complex operator+(Complex a, float s) { return {a.re + s, a.im}; }
inline Complex operator-(Complex a) { return {-a.re, -a.im}; }
inline Complex operator/(Complex a, Complex b) {
    float den = b.re * b.re + b.im * b.im;
    return {(a.re * b.re + a.im * b.im) / den, (a.im * b.re - a.re * b.im) / den};
}
inline float csqmag(Complex a) { return a.re * a.re + a.im * a.im; }
inline Complex csqrt(Complex a) {  // complex.rs:197-216 (polar form)
    float r = std::sqrt(csqmag(a)), theta = std::atan2(a.im, a.re);
    float sr = std::sqrt(r), ht = theta / 2.0f;
    return {sr * std::cos(ht), sr * std::sin(ht)};
}

// materials.rs:992-1009; false => None
inline bool refract(float eta, Vec3 wo, Vec3 normal, Vec3& out) {
    float cos_i = dot(wo, normal);
    if (cos_i < 0.0f) { eta = 1.0f / eta; cos_i = -cos_i; normal = -normal; }
    float sin2_i = 1.0f - cos_i * cos_i;
    float sin2_t = sin2_i / (eta * eta);
    if (sin2_t >= 1.0f) return false;
    float cos_t = std::sqrt(1.0f - sin2_t);
    out = -wo / eta + (cos_i / eta - cos_t) * normal;
    return true;
}
// materials.rs:1018-1041
inline float fresnel_dielectric(float cos_i, float eta) {
    if (cos_i < 0.0f) { eta = 1.0f / eta; cos_i = -cos_i; }
    float sin2_i = 1.0f - cos_i * cos_i;
    float sin2_t = sin2_i / (eta * eta);
    if (sin2_t >= 1.0f) return 1.0f;
    float cos_t = std::sqrt(1.0f - sin2_t);
    float r_parl = (eta * cos_i - cos_t) / (eta * cos_i + cos_t);
    float r_perp = (cos_i - eta * cos_t) / (cos_i + eta * cos_t);
    return (r_parl * r_parl + r_perp * r_perp) / 2.0f;
}
// materials.rs:1045-1065
inline float fresnel_complex(float cos_i, Complex eta) {
    float sin2_i = 1.0f - cos_i * cos_i;
    Complex sin2_t = Complex{sin2_i, 0.0f} / (eta * eta);
    Complex cos2_t = -sin2_t + 1.0f;
    Complex cos_t = csqrt(cos2_t);
    Complex r_parl = (eta * cos_i - cos_t) / (eta * cos_i + cos_t);
    Complex r_perp = (Complex{cos_i, 0.0f} - eta * cos_t) / (Complex{cos_i, 0.0f} + eta * cos_t);
    return (csqmag(r_parl) + csqmag(r_perp)) / 2.0f;
}

namespace microfacet {  // materials.rs:1068-1474
inline float distribution(Vec3 wm, float ax, float ay) {
    float c2 = wm.z * wm.z, s2 = 1.0f - c2;
    float e = (wm.x * wm.x) / (ax * ax) + (wm.y * wm.y) / (ay * ay);
    float t = (1.0f + (s2 / c2) * e) * (1.0f + (s2 / c2) * e);
    return 1.0f / (PI * ax * ay * c2 * c2 * t);
}
inline float lambda(Vec3 w, float ax, float ay) {
    float c2 = w.z * w.z, s2 = 1.0f - c2, tan2 = s2 / c2;
    float a2 = ax * ax * w.x * w.x + ay * ay * w.y * w.y;
    return (std::sqrt(1.0f + a2 * tan2) - 1.0f) / 2.0f;
}
inline float G1(Vec3 w, float ax, float ay) { return 1.0f / (1.0f + lambda(w, ax, ay)); }
inline float G(Vec3 wo, Vec3 wi, float ax, float ay) { return 1.0f / (1.0f + lambda(wo, ax, ay) + lambda(wi, ax, ay)); }
inline float visible_distribution(Vec3 w, Vec3 wm, float ax, float ay) {
    float cos_theta = std::fabs(w.z);
    return (G1(w, ax, ay) / cos_theta) * distribution(wm, ax, ay) * std::fabs(dot(w, wm));
}
inline Vec3 sample_wm(Vec3 w, float ax, float ay, Vec2 u) {
    Vec3 wh = unit(Vec3{ax * w.x, ay * w.y, w.z});
    if (wh.z < 0.0f) wh = -wh;
    Vec2 p = sample_unit_disk(u);
    Vec3 t1 = wh.z < 0.9999f ? cross(Vec3{0, 0, 1}, wh) : Vec3{1, 0, 0};
    Vec3 t2 = cross(wh, t1);
    float h = std::sqrt(1.0f - p.x * p.x);
    float offset = 0.5f * h * (1.0f - wh.z);
    float scale = 0.5f * (1.0f + wh.z);
    p = Vec2{p.x, offset + scale * p.y};
    float pz = std::sqrt(std::fmax(0.0f, 1.0f - sqmag(p)));
    Vec3 nh = p.x * t1 + p.y * t2 + pz * wh;
    return unit(Vec3{ax * nh.x, ay * nh.y, std::fmax(1.0e-6f, nh.z)});
}
inline float refl_pdf(Vec3 wo, Vec3 wi, float ax, float ay) {
    if ((wo + wi) == Vec3{0, 0, 0}) return 0.0f;
    Vec3 wm = unit(wo + wi);
    if (wm.z < 0.0f) wm = -wm;
    return visible_distribution(wo, wm, ax, ay) / (4.0f * std::fabs(dot(wo, wm)));
}
inline Vec3 refl_bsdf(Vec3 wo, Vec3 wi, Vec3 eta, Vec3 kappa, float ax, float ay) {
    if ((wo + wi) == Vec3{0, 0, 0}) return {0, 0, 0};
    Vec3 wm = unit(wo + wi);
    float cos_theta = dot(wm, wi);
    Vec3 fr{fresnel_complex(std::fabs(cos_theta), {eta.x, kappa.x}), fresnel_complex(std::fabs(cos_theta), {eta.y, kappa.y}),
            fresnel_complex(std::fabs(cos_theta), {eta.z, kappa.z})};
    return distribution(wm, ax, ay) * fr * G(wo, wi, ax, ay) / (4.0f * wo.z * wi.z);
}
inline SampleStatus refl_sample(Vec3 wo, Vec3 eta, Vec3 kappa, float ax, float ay, Sampler& s, BsdfSample& out) {
    Vec2 u = s.uniform2();
    Vec3 wm = sample_wm(wo, ax, ay, u);
    Vec3 wi = reflect(wo, wm);
    if (wo.z * wi.z < 0.0f) return NULL_SAMPLE;
    out.wi = wi;
    out.pdf = refl_pdf(wo, wi, ax, ay);
    out.bsdf = refl_bsdf(wo, wi, eta, kappa, ax, ay);
    out.component = NONSPEC_REFL;
    return validate(out);
}
inline float ts_pdf(Vec3 wo, Vec3 wi, float eta, float ax, float ay, uint8_t component) {
    bool refl = wo.z * wi.z > 0.0f;
    float eta_wm = !refl ? (wo.z > 0.0f ? eta : 1.0f / eta) : 1.0f;
    Vec3 wm = unit(wi * eta_wm + wo);
    if (wm.z < 0.0f) wm = -wm;
    if (wi.z == 0.0f || wo.z == 0.0f || wm == Vec3{0, 0, 0}) return 0.0f;
    if (dot(wm, wi) * wi.z < 0.0f || dot(wm, wo) * wo.z < 0.0f) return 0.0f;
    float R = fresnel_dielectric(dot(wo, wm), eta), T = 1.0f - R;
    float pr = (component & NONSPEC_REFL) ? R : 0.0f;
    float pt = (component & NONSPEC_TRANS) ? T : 0.0f;
    float ptot = pr + pt;
    if (refl) return (pr / ptot) * visible_distribution(wo, wm, ax, ay) / (4.0f * std::fabs(dot(wo, wm)));
    float d = dot(wi, wm) + dot(wo, wm) / eta_wm;
    float denom = d * d;
    float dwm_dwi = std::fabs(dot(wi, wm)) / denom;
    return (pt / ptot) * visible_distribution(wo, wm, ax, ay) * dwm_dwi;
}
inline Vec3 ts_bsdf(Vec3 wo, Vec3 wi, float eta, float ax, float ay) {
    bool refl = wo.z * wi.z > 0.0f;
    float eta_wm = !refl ? (wo.z > 0.0f ? eta : 1.0f / eta) : 1.0f;
    Vec3 wm = unit(wi * eta_wm + wo);
    if (wm.z < 0.0f) wm = -wm;
    if (wi.z == 0.0f || wo.z == 0.0f || wm == Vec3{0, 0, 0}) return {0, 0, 0};
    if (dot(wm, wi) * wi.z < 0.0f || dot(wm, wo) * wo.z < 0.0f) return {0, 0, 0};
    float F = fresnel_dielectric(dot(wo, wm), eta);
    if (refl) {
        float brdf = distribution(wm, ax, ay) * F * G(wo, wi, ax, ay) / std::fabs(4.0f * wo.z * wi.z);
        return {brdf, brdf, brdf};
    }
    float d = dot(wi, wm) + dot(wo, wm) / eta_wm;
    float denom = wi.z * wo.z * (d * d);
    float btdf = distribution(wm, ax, ay) * (1.0f - F) * G(wo, wi, ax, ay) * std::fabs(dot(wi, wm) * dot(wo, wm) / denom) / (eta_wm * eta_wm);
    return {btdf, btdf, btdf};
}
inline SampleStatus ts_sample(Vec3 wo, float eta, float ax, float ay, uint8_t component, Sampler& s, BsdfSample& out) {
    Vec2 u = s.uniform2();
    Vec3 wm = sample_wm(wo, ax, ay, u);
    float R = fresnel_dielectric(dot(wo, wm), eta), T = 1.0f - R;
    float pr = (component & REFLECTION) == REFLECTION ? R : 0.0f;      // contains(REFLECTION)
    float pt = (component & TRANSMISSION) == TRANSMISSION ? T : 0.0f;  // contains(TRANSMISSION)
    float ptot = pr + pt;
    Vec3 wi;
    bool reflected;
    if (s.uniform() * ptot < pr) {
        wi = reflect(wo, wm);
        if (wo.z * wi.z < 0.0f) return NULL_SAMPLE;
        reflected = true;
    } else {
        if (!refract(eta, wo, wm, wi)) return INVALID;
        if (wo.z * wi.z > 0.0f || wi.z == 0.0f) return NULL_SAMPLE;
        reflected = false;
    }
    out.wi = wi;
    out.pdf = ts_pdf(wo, wi, eta, ax, ay, component);
    out.bsdf = ts_bsdf(wo, wi, eta, ax, ay);
    out.component = reflected ? NONSPEC_REFL : NONSPEC_TRANS;
    return validate(out);
}
}  // namespace microfacet

namespace phase {  // materials.rs:1477-1535
inline float hg(float cos_theta, float g) {
    float denom = 1.0f + g * g + 2.0f * g * cos_theta;
    return FRAC_1_PI * 0.25f * (1.0f - g * g) / (denom * std::sqrt(denom));
}
struct PhaseSample { Vec3 wi; float p, pdf; };
inline PhaseSample sample_p(Vec3 wo, float g, Vec2 u) {
    float cos_theta;
    if (std::fabs(g) < 1.0e-3f) cos_theta = 1.0f - 2.0f * u.x;
    else {
        float term = (1.0f - g * g) / (1.0f + g - 2.0f * g * u.x);
        cos_theta = -1.0f / (2.0f * g) * (1.0f + g * g - term * term);
    }
    float phi = 2.0f * PI * u.y;
    float sin_theta = std::sqrt(1.0f - cos_theta * cos_theta);
    Vec3 d{std::cos(phi) * sin_theta, std::sin(phi) * sin_theta, cos_theta};
    Vec3 wx, wy;
    make_orthonormal_basis(wo, wx, wy);
    Vec3 wi = d.x * wx + d.y * wy + d.z * wo;
    float p = hg(cos_theta, g);
    return {wi, p, p};
}
inline float p(Vec3 wo, Vec3 wi, float g) { return hg(dot(wo, wi), g); }
}  // namespace phase

enum BsdfKind { B_DIFFUSE, B_SMOOTH_DIELECTRIC, B_SMOOTH_CONDUCTOR, B_ROUGH_CONDUCTOR, B_ROUGH_DIELECTRIC, B_LAYERED };

struct Bsdf {  // materials.rs:43-80
    BsdfKind kind = B_DIFFUSE;
    Vec3 albedo;           // Diffuse; Layered: medium albedo
    float eta = 1;         // dielectrics
    Vec3 eta3, kappa;      // conductors
    float ax = 0, ay = 0;
    // Layered
    std::shared_ptr<Bsdf> top, bottom;
    uint32_t n_samples = 0, max_depth = 0;
    float thickness = 0, g = 0;

    bool is_delta() const { return kind == B_SMOOTH_DIELECTRIC || kind == B_SMOOTH_CONDUCTOR; }  // :671-680
    uint8_t components() const {  // :682-691 (Layered is todo!() in the reference: never reached for nested layers)
        switch (kind) {
            case B_DIFFUSE: return NONSPEC_REFL;
            case B_SMOOTH_DIELECTRIC: return SPEC_REFL | SPEC_TRANS;
            case B_SMOOTH_CONDUCTOR: return SPEC_REFL;
            case B_ROUGH_CONDUCTOR: return NONSPEC_REFL;
            case B_ROUGH_DIELECTRIC: return NONSPEC_REFL | NONSPEC_TRANS;
            default: return 0;
        }
    }
    static float Tr(float dz, Vec3 w) { return std::exp(-std::fabs(dz / w.z)); }  // :84-87

    float evaluate_pdf(Vec3 wo, Vec3 wi, uint8_t component) const {  // :338-370
        switch (kind) {
            case B_DIFFUSE:
                if (!(component & NONSPEC_REFL)) return 0.0f;
                return wo.z * wi.z > 0.0f ? 1.0f / (2.0f * PI) : 0.0f;
            case B_ROUGH_CONDUCTOR:
                if (!(component & NONSPEC_REFL)) return 0.0f;
                return microfacet::refl_pdf(wo, wi, ax, ay);
            case B_ROUGH_DIELECTRIC: return microfacet::ts_pdf(wo, wi, eta, ax, ay, component);
            default: return 0.0f;
        }
    }

    Vec3 evaluate(Vec3 wo, Vec3 wi) const;                                   // :125-336
    SampleStatus sample(Vec3 wo, uint8_t component, Sampler& s, BsdfSample& out) const;  // :377-668
};

SampleStatus Bsdf::sample(Vec3 wo, uint8_t component, Sampler& s, BsdfSample& out) const {
    switch (kind) {
        case B_DIFFUSE: {
            if (!(component & NONSPEC_REFL)) return INVALID;
            Vec2 u = s.uniform2();
            Vec3 wi = sample_cosine_hemisphere(u);
            out = {wi, albedo / PI, wi.z / PI, NONSPEC_REFL};
            return validate(out);
        }
        case B_SMOOTH_DIELECTRIC: {
            if (!(component & (SPEC_REFL | SPEC_TRANS))) return INVALID;
            Vec3 normal{0, 0, 1};
            float R = fresnel_dielectric(wo.z, eta), T = 1.0f - R;
            float pr = (component & SPEC_REFL) ? R : 0.0f, pt = (component & SPEC_TRANS) ? T : 0.0f;
            float ptot = pr + pt;
            float smp = s.uniform();
            if (smp * ptot < pr) {
                Vec3 rd = reflect(wo, normal);
                float f = R / std::fabs(rd.z);
                out = {rd, {f, f, f}, R / ptot, SPEC_REFL};
            } else {
                Vec3 rd;
                if (!refract(eta, wo, normal, rd)) return INVALID;
                float e = wo.z < 0.0f ? 1.0f / eta : eta;
                float f = (T / std::fabs(rd.z)) / (e * e);
                out = {rd, {f, f, f}, T / ptot, SPEC_TRANS};
            }
            return validate(out);
        }
        case B_SMOOTH_CONDUCTOR: {
            Vec3 rd = reflect(wo, Vec3{0, 0, 1});
            Vec3 f{fresnel_complex(wo.z, {eta3.x, kappa.x}) / wo.z, fresnel_complex(wo.z, {eta3.y, kappa.y}) / wo.z,
                   fresnel_complex(wo.z, {eta3.z, kappa.z}) / wo.z};
            out = {rd, f, 1.0f, SPEC_REFL};
            return validate(out);
        }
        case B_ROUGH_CONDUCTOR: return microfacet::refl_sample(wo, eta3, kappa, ax, ay, s, out);
        case B_ROUGH_DIELECTRIC: return microfacet::ts_sample(wo, eta, ax, ay, component, s, out);
        case B_LAYERED: {  // :540-666
            bool flip_wi = false;
            if (wo.z < 0.0f) { wo = -wo; flip_wi = true; }
            BsdfSample enter;
            SampleStatus st = top->sample(wo, ALL, s, enter);
            if (st != VALID) return st;
            if (enter.component & REFLECTION) {
                out = enter;
                if (flip_wi) out.wi = -enter.wi;
                return validate(out);
            }
            bool specular_path = (enter.component & SPECULAR) != 0;
            Vec3 w = enter.wi;
            Vec3 f = enter.bsdf * std::fabs(enter.wi.z);
            float pdf = enter.pdf;
            float z = thickness;
            for (uint32_t depth = 0; depth < max_depth; depth++) {
                float rr_beta = max_component(f) / pdf;
                if (depth > 3 && rr_beta < 0.25f) {
                    float q = std::fmax(0.0f, 1.0f - rr_beta);
                    if (s.uniform() < q) return NULL_SAMPLE;
                    pdf *= 1.0f - q;
                }
                if (w.z == 0.0f) return NULL_SAMPLE;
                if (albedo != Vec3{0, 0, 0}) {
                    float sigma_t = 1.0f;
                    float dz = sample_exponential(s.uniform(), sigma_t / std::fabs(w.z));
                    float zp = w.z > 0.0f ? z + dz : z - dz;
                    if (zp == z) return INVALID;
                    if (0.0f < zp && zp < thickness) {
                        phase::PhaseSample ps = phase::sample_p(-w, g, s.uniform2());
                        if (ps.wi.z == 0.0f) return NULL_SAMPLE;
                        f *= albedo * ps.p;
                        pdf *= ps.pdf;
                        specular_path = false;
                        w = ps.wi;
                        z = zp;
                        continue;
                    }
                    z = rs_clamp(zp, 0.0f, thickness);
                } else {
                    z = (z == thickness) ? 0.0f : thickness;
                    f *= Tr(thickness, w);
                }
                const Bsdf& iface = (z == 0.0f) ? *bottom : *top;
                BsdfSample is;
                st = iface.sample(-w, ALL, s, is);
                if (st != VALID) return st;
                f *= is.bsdf;
                pdf *= is.pdf;
                specular_path = specular_path && (is.component & SPECULAR) != 0;
                w = is.wi;
                if (is.component & TRANSMISSION) {
                    bool same_dir = wo.z * w.z > 0.0f;
                    uint8_t comp = same_dir ? (specular_path ? SPEC_REFL : NONSPEC_REFL) : (specular_path ? SPEC_TRANS : NONSPEC_TRANS);
                    if (flip_wi) w = -w;
                    out = {w, f, pdf, comp};
                    return validate(out);
                }
                f *= std::fabs(is.wi.z);
            }
            return NULL_SAMPLE;
        }
    }
    return INVALID;
}

Vec3 Bsdf::evaluate(Vec3 wo, Vec3 wi) const {
    switch (kind) {
        case B_DIFFUSE: return wo.z * wi.z < 0.0f ? Vec3{0, 0, 0} : albedo / PI;
        case B_SMOOTH_DIELECTRIC:
        case B_SMOOTH_CONDUCTOR: return {0, 0, 0};
        case B_ROUGH_CONDUCTOR: return microfacet::refl_bsdf(wo, wi, eta3, kappa, ax, ay);
        case B_ROUGH_DIELECTRIC: return microfacet::ts_bsdf(wo, wi, eta, ax, ay);
        case B_LAYERED: break;
    }
    // LayeredBsdf, materials.rs:171-333
    Vec3 f{0, 0, 0};
    if (wo.z < 0.0f) { wo = -wo; wi = -wi; }
    const Bsdf* enter_if = top.get();
    const Bsdf *exit_if, *non_exit_if;
    float exit_z;
    if (wi.z < 0.0f) { exit_if = bottom.get(); non_exit_if = top.get(); exit_z = 0.0f; }
    else { exit_if = top.get(); non_exit_if = bottom.get(); exit_z = thickness; }
    if (!(enter_if->components() & TRANSMISSION) || !(exit_if->components() & TRANSMISSION)) return {0, 0, 0};
    if (wo.z * wi.z > 0.0f) f += (float)n_samples * enter_if->evaluate(wo, wi);
    FxHasher h;
    auto bits = [](float v) { uint32_t b; std::memcpy(&b, &v, 4); return b; };
    h.write_u32(bits(wi.x)); h.write_u32(bits(wi.y)); h.write_u32(bits(wi.z));
    h.write_u32(bits(wo.x)); h.write_u32(bits(wo.y)); h.write_u32(bits(wo.z));
    Sampler s = Sampler::one_off(h.finish());
    for (uint32_t i = 0; i < n_samples; i++) {
        BsdfSample enter, exit;
        if (enter_if->sample(wo, TRANSMISSION, s, enter) != VALID) continue;
        if (exit_if->sample(wi, TRANSMISSION, s, exit) != VALID) continue;
        Vec3 beta = exit.bsdf * std::fabs(exit.wi.z) / exit.pdf;
        float z = thickness;
        Vec3 w = enter.wi;
        for (uint32_t depth = 0; depth < max_depth; depth++) {
            if (depth > 3 && max_component(beta) < 0.25f) {
                float q = std::fmax(0.0f, max_component(beta));
                if (s.uniform() < q) break;
                beta /= 1.0f - q;
            }
            if (albedo == Vec3{0, 0, 0}) {
                z = (z == thickness) ? 0.0f : thickness;
                beta *= Tr(thickness, w);
            } else {
                float sigma_t = 1.0f;
                float dz = sample_exponential(s.uniform(), sigma_t / std::fabs(w.z));
                float zp = w.z > 0.0f ? z + dz : z - dz;
                if (0.0f < zp && zp < thickness) {
                    float wt = exit_if->is_delta() ? 1.0f : power_heuristic(1, exit.pdf, 1, phase::p(-w, -exit.wi, g));
                    f += beta * albedo * phase::p(-w, -exit.wi, g) * wt * Tr(zp - exit_z, exit.wi) * exit.bsdf / exit.pdf;
                    Vec2 u = s.uniform2();
                    phase::PhaseSample ps = phase::sample_p(-w, g, u);
                    beta *= albedo * ps.p / ps.pdf;
                    w = ps.wi;
                    z = zp;
                    bool facing_exit = (z < exit_z && w.z > 0.0f) || (z > exit_z && w.z < 0.0f);
                    if (!exit_if->is_delta() && facing_exit) {
                        Vec3 exit_f = exit_if->evaluate(-w, wi);
                        if (exit_f != Vec3{0, 0, 0}) {
                            float exit_pdf = exit_if->evaluate_pdf(-w, wi, TRANSMISSION);
                            float wt2 = power_heuristic(1, ps.pdf, 1, exit_pdf);
                            f += beta * Tr(zp - exit_z, ps.wi) * exit_f * wt2;
                        }
                    }
                    continue;
                }
                z = rs_clamp(zp, 0.0f, thickness);
            }
            if (z == exit_z) {
                BsdfSample rs;
                if (exit_if->sample(-w, REFLECTION, s, rs) != VALID) break;
                beta *= rs.bsdf * std::fabs(rs.wi.z) / rs.pdf;
                w = rs.wi;
            } else {
                if (!non_exit_if->is_delta()) {
                    float wt = power_heuristic(1, exit.pdf, 1, non_exit_if->evaluate_pdf(-w, -exit.wi, REFLECTION));
                    f += beta * non_exit_if->evaluate(-w, -exit.wi) * std::fabs(exit.wi.z) * wt * Tr(thickness, exit.wi) * exit.bsdf / exit.pdf;
                }
                BsdfSample ns;
                if (non_exit_if->sample(-w, REFLECTION, s, ns) != VALID) break;
                beta *= ns.bsdf * std::fabs(ns.wi.z) / ns.pdf;
                w = ns.wi;
                if (!exit_if->is_delta()) {
                    Vec3 exit_f = exit_if->evaluate(-w, wi);
                    if (exit_f != Vec3{0, 0, 0}) {
                        float exit_pdf = exit_if->evaluate_pdf(-w, wi, ALL);
                        float wt = non_exit_if->is_delta() ? 1.0f : power_heuristic(1, ns.pdf, 1, exit_pdf);
                        f += beta * Tr(thickness, ns.wi) * exit_f * wt;
                    }
                }
            }
        }
    }
    return f / (float)n_samples;
}

constexpr float MINIMUM_ROUGHNESS = 1.0e-3f;  // materials.rs:1538-1541

// CpuMaterial for Material, materials.rs:823-990
Bsdf get_bsdf(const SceneView& sv, const rtcuda_material& m, const MaterialEvalContext& c) {
    auto rgb = [&](uint32_t id) { Vec4 v = sv.sample(id, c); return Vec3{v.x, v.y, v.z}; };
    auto alphas = [&](uint32_t id, bool remap, float& ax, float& ay) {
        Vec4 r = sv.sample(id, c);
        ax = r.x; ay = r.y;
        if (remap) { ax = std::sqrt(ax); ay = std::sqrt(ay); }
    };
    Bsdf b;
    switch (m.kind) {
        case RTCUDA_MATERIAL_DIFFUSE: b.kind = B_DIFFUSE; b.albedo = rgb(m.albedo); break;
        case RTCUDA_MATERIAL_SMOOTH_DIELECTRIC: b.kind = B_SMOOTH_DIELECTRIC; b.eta = sv.sample(m.eta, c).x; break;
        case RTCUDA_MATERIAL_SMOOTH_CONDUCTOR: b.kind = B_SMOOTH_CONDUCTOR; b.eta3 = rgb(m.eta); b.kappa = rgb(m.kappa); break;
        case RTCUDA_MATERIAL_ROUGH_CONDUCTOR: {
            b.eta3 = rgb(m.eta); b.kappa = rgb(m.kappa);
            alphas(m.roughness, m.remap_roughness, b.ax, b.ay);
            b.kind = std::fmax(b.ax, b.ay) < MINIMUM_ROUGHNESS ? B_SMOOTH_CONDUCTOR : B_ROUGH_CONDUCTOR;
            break;
        }
        case RTCUDA_MATERIAL_ROUGH_DIELECTRIC: {
            b.eta = sv.sample(m.eta, c).x;
            alphas(m.roughness, m.remap_roughness, b.ax, b.ay);
            b.kind = std::fmax(b.ax, b.ay) < MINIMUM_ROUGHNESS ? B_SMOOTH_DIELECTRIC : B_ROUGH_DIELECTRIC;
            break;
        }
        default: {  // CoatedDiffuse
            auto bottom = std::make_shared<Bsdf>();
            bottom->kind = B_DIFFUSE;
            bottom->albedo = rgb(m.albedo);
            auto top = std::make_shared<Bsdf>();
            top->eta = sv.sample(m.eta, c).x;
            top->kind = B_SMOOTH_DIELECTRIC;
            if (m.roughness != RTCUDA_NONE) {
                alphas(m.roughness, m.remap_roughness, top->ax, top->ay);
                if (!(std::fmax(top->ax, top->ay) < MINIMUM_ROUGHNESS)) top->kind = B_ROUGH_DIELECTRIC;
            }
            b.kind = B_LAYERED;
            b.top = top;
            b.bottom = bottom;
            b.n_samples = 8;
            b.max_depth = 8;
            b.thickness = sv.sample(m.thickness, c).x;
            b.albedo = rgb(m.coat_albedo);
            b.g = 0.0f;
        }
    }
    return b;
}
bool get_mip_level(const SceneView& sv, const rtcuda_material& m, const MaterialEvalContext& c, float& level) {
    if (m.kind != RTCUDA_MATERIAL_DIFFUSE) return false;
    return sv.texture_mip_level(m.albedo, c, level);
}
Vec3 get_albedo(const SceneView& sv, const rtcuda_material& m, const MaterialEvalContext& c) {
    if (m.kind == RTCUDA_MATERIAL_DIFFUSE || m.kind == RTCUDA_MATERIAL_COATED_DIFFUSE) {
        Vec4 v = sv.sample(m.albedo, c);
        return {v.x, v.y, v.z};
    }
    return {1, 1, 1};
}

// ---------------------------------------------------------------------------------------------
// Lights. crates/raytracing-cpu/src/lights.rs
// ---------------------------------------------------------------------------------------------
struct LightSample { Vec3 radiance; Ray shadow_ray; float distance; float pdf; };

LightSample sample_light(const Context& ctx, const rtcuda_light& l, Vec3 point, Sampler& s) {  // :14-122
    Vec3 a{l.position_or_direction[0], l.position_or_direction[1], l.position_or_direction[2]};
    Vec3 b{l.intensity_or_radiance[0], l.intensity_or_radiance[1], l.intensity_or_radiance[2]};
    if (l.kind == RTCUDA_LIGHT_POINT) {
        Vec3 dir = point - a;
        float d = length(dir), d2 = d * d;
        return {b / d2, {a, dir / d}, d, 1.0f};
    }
    if (l.kind == RTCUDA_LIGHT_DIRECTION) {
        float diam = ctx.scene_radius * 2.0f;
        Vec3 origin = point - a * diam;
        return {b, {origin, unit(a)}, diam, 1.0f};
    }
    const MeshView& em = ctx.sv->meshes[l.shape];
    float pdf = 1.0f;
    pdf /= (float)em.tri_count;
    uint32_t tri = s.u32_range(0, em.tri_count);
    Vec2 smp = s.uniform2();
    Vec3 bary;
    if (smp.x < smp.y) { float b0 = smp.x / 2.0f, b1 = smp.y - smp.x / 2.0f; bary = {b0, b1, 1.0f - b0 - b1}; }
    else { float b0 = smp.x - smp.y / 2.0f, b1 = smp.y / 2.0f; bary = {b0, b1, 1.0f - b0 - b1}; }
    pdf /= em.tri_area(tri);
    uint32_t i0 = em.tris[3 * tri], i1 = em.tris[3 * tri + 1], i2 = em.tris[3 * tri + 2];
    Vec3 p0 = em.v(i0), p1 = em.v(i1), p2 = em.v(i2);
    Vec3 p_local = bary.x * p0 + bary.y * p1 + bary.z * p2;
    Vec3 p_world = mat_from(l.light_to_world).apply_point(p_local);
    Vec3 dir_world = point - p_world;
    float d = length(dir_world);
    Ray shadow{p_world, dir_world / d};
    Vec3 n = !em.normals ? unit(cross(p2 - p0, p1 - p0)) : unit(bary.x * em.n(i0) + bary.y * em.n(i1) + bary.z * em.n(i2));
    Vec3 radiance = dot(dir_world, n) < 0.0f ? Vec3{0, 0, 0} : b;
    pdf *= (d * d) / std::fabs(dot(dir_world, n));
    return {radiance, shadow, d, pdf};
}

inline Vec3 light_radiance(const rtcuda_light& l) {  // :124-135
    if (l.kind == RTCUDA_LIGHT_DIFFUSE_AREA) return {l.intensity_or_radiance[0], l.intensity_or_radiance[1], l.intensity_or_radiance[2]};
    return {0, 0, 0};
}
Vec3 environment_light_radiance(const Context& ctx, Vec3 direction) {  // :137-157
    direction = unit(direction);
    float t = std::acos(direction.z) * FRAC_1_PI;
    float s = (std::atan2(direction.x, direction.y) + PI) * FRAC_1_PI * 0.5f;
    Vec4 v = ctx.sv->sample(ctx.sv->d->environment_light_texture, mec_no_aa({s, t}));
    return {v.x, v.y, v.z};
}
bool occluded(const Context& ctx, const LightSample& ls) {  // :159-168
    HitInfo h;
    ctx.n_shadow.fetch_add(1, std::memory_order_relaxed);
    return traverse_bvh(ctx, ls.shadow_ray, 0.001f, ls.distance - 0.001f, true, h);
}

// ---------------------------------------------------------------------------------------------
// Camera + integrator. crates/raytracing-cpu/src/lib.rs
// ---------------------------------------------------------------------------------------------
Ray camera_ray(const rtcuda_camera& cam, float x, float y, bool has_lens, Vec2 lens) {  // lib.rs:145-195
    Transform r2c = tf_from(cam.raster_to_camera), c2w = tf_from(cam.camera_to_world);
    Vec3 raster{x, y, 0.0f};
    if (cam.kind == RTCUDA_CAMERA_ORTHOGRAPHIC) {
        Vec3 o = r2c.apply_point(raster);
        return {c2w.apply_point(o), unit(c2w.apply_vector({0, 0, 1}))};
    }
    if (cam.kind == RTCUDA_CAMERA_PINHOLE) {
        Vec3 cp = r2c.apply_point(raster);
        Vec3 dir = unit(cp);
        return {c2w.apply_point({0, 0, 0}), unit(c2w.apply_vector(dir))};
    }
    Vec3 cp = r2c.apply_point(raster);
    float t = cam.focal_distance / cp.z;
    Vec3 focus = cp * t;
    Vec3 co{0, 0, 0}, cd;
    if (has_lens) {
        co = {lens.x * cam.aperture_radius, lens.y * cam.aperture_radius, 0.0f};
        cd = unit(focus - co);
    } else cd = unit(cp);
    return {c2w.apply_point(co), unit(c2w.apply_vector(cd))};
}

void generate_ray(const rtcuda_camera& cam, uint32_t px, uint32_t py, Sampler& s, uint32_t spp, bool jitter, Ray& ray, RayDifferentials& rd) {  // lib.rs:198-245
    float x, y;
    if (jitter) { Vec2 d = s.uniform2(); x = (float)px + d.x; y = (float)py + d.y; }
    else { x = (float)px + 0.5f; y = (float)py + 0.5f; }
    bool has_lens = cam.kind == RTCUDA_CAMERA_THIN_LENS;
    Vec2 lens{};
    if (has_lens) lens = sample_unit_disk_concentric(s.uniform2());
    ray = camera_ray(cam, x, y, has_lens, lens);
    Ray rx = camera_ray(cam, x + 1.0f, y, has_lens, lens);
    Ray ry = camera_ray(cam, x, y + 1.0f, has_lens, lens);
    float scale = std::fmax(0.125f, std::sqrt(1.0f / (float)spp));
    Vec3 sx = ray.direction + (rx.direction - ray.direction) * scale;
    Vec3 sy = ray.direction + (ry.direction - ray.direction) * scale;
    rd = {rx.origin - ray.origin, ry.origin - ray.origin, unit(sx) - ray.direction, unit(sy) - ray.direction};
}

// Matrix4x4::create_from_basis(x,y,n) and its transpose, lib.rs:311-316
struct Frame {
    Vec3 x, y, n;
    Vec3 to_local(Vec3 v) const { return {x.x * v.x + x.y * v.y + x.z * v.z, y.x * v.x + y.y * v.y + y.z * v.z, n.x * v.x + n.y * v.y + n.z * v.z}; }
    Vec3 to_world(Vec3 v) const { return {x.x * v.x + y.x * v.y + n.x * v.z, x.y * v.x + y.y * v.y + n.y * v.z, x.z * v.x + y.z * v.y + n.z * v.z}; }
};

Vec3 ray_radiance(Ray ray, const RayDifferentials& rd, const Context& ctx, Sampler& s, const rtcuda_settings& st) {  // lib.rs:247-393
    const rtcuda_scene_desc& d = *ctx.sv->d;
    uint32_t depth = 0;
    bool specular_bounce = true;
    Vec3 radiance{0, 0, 0}, path_weight{1, 1, 1};
    Ray cam_ray = ray;
    for (;;) {
        float t_min = depth == 0 ? d.camera.near_clip : 0.0001f;
        float t_max = depth == 0 ? d.camera.far_clip : INF;
        HitInfo hit;
        (depth == 0 ? ctx.n_primary : ctx.n_bounce).fetch_add(1, std::memory_order_relaxed);
        if (!traverse_bvh(ctx, ray, t_min, t_max, false, hit)) {
            if (d.environment_light_texture != RTCUDA_NONE) radiance += path_weight * environment_light_radiance(ctx, ray.direction);
            break;
        }
        bool add_zero_bounce = st.accumulate_bounces || st.max_ray_depth == depth;
        if (specular_bounce && add_zero_bounce && hit.light_idx != RTCUDA_NONE) radiance += path_weight * light_radiance(d.lights[hit.light_idx]);
        const rtcuda_material& mat = d.materials[hit.material_idx];
        MaterialEvalContext mec = (depth == 0 && st.antialias_primary_rays) ? mec_from_differentials(hit, cam_ray, rd) : mec_no_aa(hit.uv);
        Bsdf bsdf = get_bsdf(*ctx.sv, mat, mec);
        Frame fr;
        fr.n = hit.normal;
        make_orthonormal_basis(hit.normal, fr.x, fr.y);
        Vec3 wo = fr.to_local(-ray.direction);
        depth += 1;
        bool delta = bsdf.is_delta();
        if (depth > st.max_ray_depth) break;
        bool add_direct = st.accumulate_bounces || st.max_ray_depth == depth;
        if (!delta && add_direct) {
            Vec3 direct{0, 0, 0};
            for (uint32_t li = 0; li < d.light_count; li++) {
                const rtcuda_light& light = d.lights[li];
                Vec3 contrib{0, 0, 0};
                uint32_t n = light.kind == RTCUDA_LIGHT_DIFFUSE_AREA ? st.light_sample_count : 1;
                for (uint32_t k = 0; k < n; k++) {
                    LightSample ls = sample_light(ctx, light, hit.point, s);
                    if (!occluded(ctx, ls)) {
                        Vec3 wi = fr.to_local(-ls.shadow_ray.direction);
                        Vec3 bv = bsdf.evaluate(wo, wi);
                        contrib += bv * ls.radiance * std::fmax(0.0f, wi.z) / ls.pdf;
                    }
                }
                contrib /= (float)n;
                direct += contrib;
            }
            radiance += path_weight * direct;
        }
        BsdfSample bs;
        if (bsdf.sample(wo, ALL, s, bs) != VALID) break;
        if (bs.bsdf == Vec3{0, 0, 0} || bs.pdf == 0.0f) break;
        path_weight *= bs.bsdf * std::fabs(bs.wi.z) / bs.pdf;
        specular_bounce = (bs.component & SPECULAR) != 0;
        ray = {hit.point, fr.to_world(bs.wi)};
    }
    return radiance;
}

struct FirstHit { bool hit = false; Vec2 uv; Vec3 normal, albedo; bool has_mip = false; float mip = 0; uint32_t geom = RTCUDA_NONE, prim = RTCUDA_NONE; float t = 0; };
FirstHit first_hit_aovs(const Ray& ray, const RayDifferentials& rd, const Context& ctx) {  // lib.rs:403-444
    const rtcuda_scene_desc& d = *ctx.sv->d;
    FirstHit r;
    HitInfo hit;
    ctx.n_aov.fetch_add(1, std::memory_order_relaxed);
    if (!traverse_bvh(ctx, ray, d.camera.near_clip, d.camera.far_clip, false, hit)) return r;
    const rtcuda_material& mat = d.materials[hit.material_idx];
    MaterialEvalContext mec = mec_from_differentials(hit, ray, rd);
    r.hit = true;
    r.uv = hit.uv;
    r.normal = hit.normal;
    r.albedo = get_albedo(*ctx.sv, mat, mec);
    r.has_mip = get_mip_level(*ctx.sv, mat, mec, r.mip);
    r.geom = hit.geom_id;
    r.prim = hit.prim_id;
    r.t = hit.t;
    return r;
}

struct Tile { uint32_t x0, x1, y0, y1; };

}  // namespace

// ---------------------------------------------------------------------------------------------
// C entry points (loaded by tests / bench via ctypes)
// ---------------------------------------------------------------------------------------------
extern "C" {

struct oracle_stats { uint64_t primary_rays, bounce_rays, shadow_rays, aov_rays; double render_ms, build_ms; uint64_t bvh_nodes; };

enum { ORACLE_FLAG_BRUTE_FORCE = 1 };

// raytracing_cpu::render, crates/raytracing-cpu/src/lib.rs:645-858. tile_rank/tile_world select the
// 64x64 tiles this call renders (i % world == rank), mirroring rtcuda_backend_settings.
__attribute__((visibility("default")))
int oracle_render(const rtcuda_scene_desc* desc, const rtcuda_settings* st, rtcuda_outputs* out, uint32_t num_threads,
                  uint32_t flags, uint32_t tile_rank, uint32_t tile_world, oracle_stats* stats) {
    if (!desc || !st || !out) return 1;
    uint32_t W = desc->camera.raster_width, H = desc->camera.raster_height;
    if (out->width != W || out->height != H) return 1;
    auto tb0 = std::chrono::steady_clock::now();
    SceneView sv(desc);
    Context ctx;
    build_context(ctx, sv, flags & ORACLE_FLAG_BRUTE_FORCE);
    auto tb1 = std::chrono::steady_clock::now();
    if (tile_world == 0) tile_world = 1;

    constexpr uint32_t TS = 64;  // lib.rs:481-504
    uint32_t tiles_x = (W + TS - 1) / TS, tiles_y = (H + TS - 1) / TS;
    auto tile_mine = [&](uint32_t x, uint32_t y) { return ((y / TS) * tiles_x + (x / TS)) % tile_world == tile_rank; };

    auto t0 = std::chrono::steady_clock::now();
    Sampler base = Sampler::from_settings(*st);
    if (st->outputs & (RTCUDA_AOV_FIRST_HIT | RTCUDA_AOV_DEBUG_IDS | RTCUDA_AOV_DEBUG_DEPTH)) {  // render_aovs, lib.rs:556-625
        Sampler s = base;
        for (uint32_t y = 0; y < H; y++)
            for (uint32_t x = 0; x < W; x++) {
                size_t i = (size_t)y * W + x;
                FirstHit fh;
                if (tile_mine(x, y)) {
                    s.start_sample(x, y, 0);
                    Ray ray; RayDifferentials rd;
                    generate_ray(desc->camera, x, y, s, st->samples_per_pixel, false, ray, rd);
                    fh = first_hit_aovs(ray, rd, ctx);
                }
                if ((st->outputs & RTCUDA_AOV_NORMALS) && out->normals) { out->normals[3 * i] = fh.normal.x; out->normals[3 * i + 1] = fh.normal.y; out->normals[3 * i + 2] = fh.normal.z; }
                if ((st->outputs & RTCUDA_AOV_ALBEDO) && out->albedo) { out->albedo[3 * i] = fh.albedo.x; out->albedo[3 * i + 1] = fh.albedo.y; out->albedo[3 * i + 2] = fh.albedo.z; }
                if ((st->outputs & RTCUDA_AOV_UV_COORDS) && out->uv) { out->uv[2 * i] = fh.uv.x; out->uv[2 * i + 1] = fh.uv.y; }
                if ((st->outputs & RTCUDA_AOV_MIP_LEVEL) && out->mip_level) out->mip_level[i] = fh.has_mip ? fh.mip : 0.0f;
                if ((st->outputs & RTCUDA_AOV_DEBUG_IDS) && out->debug_ids) { out->debug_ids[2 * i] = fh.geom; out->debug_ids[2 * i + 1] = fh.prim; }
                if ((st->outputs & RTCUDA_AOV_DEBUG_DEPTH) && out->debug_depth) out->debug_depth[i] = fh.hit ? fh.t : 0.0f;
            }
    }
    if ((st->outputs & RTCUDA_AOV_BEAUTY) && out->beauty) {
        std::vector<Tile> jobs;
        for (uint32_t j = 0; j < tiles_y; j++)
            for (uint32_t i = 0; i < tiles_x; i++)
                jobs.push_back({i * TS, std::min(W, (i + 1) * TS), j * TS, std::min(H, (j + 1) * TS)});
        std::memset(out->beauty, 0, sizeof(float) * 3 * (size_t)W * H);
        std::mutex mu;
        auto worker = [&]() {
            Sampler s = base;
            for (;;) {
                Tile t;
                {
                    std::lock_guard<std::mutex> g(mu);
                    if (jobs.empty()) return;
                    t = jobs.back();  // popped from the back, lib.rs:731-733
                    jobs.pop_back();
                }
                if (!tile_mine(t.x0, t.y0)) continue;
                for (uint32_t y = t.y0; y < t.y1; y++)
                    for (uint32_t x = t.x0; x < t.x1; x++) {  // render_tile, lib.rs:506-554
                        Vec3 rad{0, 0, 0};
                        for (uint32_t k = 0; k < st->samples_per_pixel; k++) {
                            s.start_sample(x, y, k);
                            Ray ray; RayDifferentials rd;
                            generate_ray(desc->camera, x, y, s, st->samples_per_pixel, true, ray, rd);
                            rad += ray_radiance(ray, rd, ctx, s, *st);
                        }
                        rad /= (float)st->samples_per_pixel;
                        size_t i = (size_t)y * W + x;
                        out->beauty[3 * i] = rad.x; out->beauty[3 * i + 1] = rad.y; out->beauty[3 * i + 2] = rad.z;
                    }
            }
        };
        if (num_threads <= 1) worker();
        else {
            std::vector<std::thread> th;
            for (uint32_t i = 0; i < num_threads; i++) th.emplace_back(worker);
            for (auto& t : th) t.join();
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    if (stats) {
        stats->primary_rays = ctx.n_primary; stats->bounce_rays = ctx.n_bounce; stats->shadow_rays = ctx.n_shadow; stats->aov_rays = ctx.n_aov;
        stats->render_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        stats->build_ms = std::chrono::duration<double, std::milli>(tb1 - tb0).count();
        stats->bvh_nodes = ctx.bvh.nodes.size();
    }
    return 0;
}

// raytracing_cpu::render_single_pixel, lib.rs:860-931, over a sample range.
__attribute__((visibility("default")))
int oracle_render_pixel(const rtcuda_scene_desc* desc, const rtcuda_settings* st, uint32_t x, uint32_t y, uint32_t lo, uint32_t hi,
                        rtcuda_pixel_output* out, uint32_t flags) {
    if (!desc || !st || !out) return 1;
    uint32_t W = desc->camera.raster_width, H = desc->camera.raster_height;
    if (x >= W || y >= H) { x = std::min(x, W - 1); y = std::min(y, H - 1); }
    SceneView sv(desc);
    Context ctx;
    build_context(ctx, sv, flags & ORACLE_FLAG_BRUTE_FORCE);
    for (uint32_t k = lo; k < hi; k++) {
        Sampler s = Sampler::from_settings(*st);
        s.start_sample(x, y, k);
        Ray ray; RayDifferentials rd;
        generate_ray(desc->camera, x, y, s, st->samples_per_pixel, false, ray, rd);
        FirstHit fh = first_hit_aovs(ray, rd, ctx);
        s.start_sample(x, y, k);
        generate_ray(desc->camera, x, y, s, st->samples_per_pixel, true, ray, rd);
        Vec3 rad = ray_radiance(ray, rd, ctx, s, *st);
        rtcuda_pixel_output& o = out[k - lo];
        o.sample_index = k;
        o.hit = fh.hit;
        o.uv[0] = fh.uv.x; o.uv[1] = fh.uv.y;
        o.normal[0] = fh.normal.x; o.normal[1] = fh.normal.y; o.normal[2] = fh.normal.z;
        o.radiance[0] = rad.x; o.radiance[1] = rad.y; o.radiance[2] = rad.z;
    }
    return 0;
}

__attribute__((visibility("default"))) uint32_t oracle_abi_struct_sizes(uint32_t* out, uint32_t capacity) {
    const uint32_t sizes[] = {sizeof(rtcuda_camera), sizeof(rtcuda_shape), sizeof(rtcuda_instance), sizeof(rtcuda_light), sizeof(rtcuda_material),
                              sizeof(rtcuda_texture), sizeof(rtcuda_image), sizeof(rtcuda_scene_desc), sizeof(rtcuda_settings),
                              sizeof(rtcuda_backend_settings), sizeof(rtcuda_outputs), sizeof(rtcuda_pixel_output), sizeof(rtcuda_stats)};
    uint32_t n = sizeof(sizes) / sizeof(sizes[0]);
    for (uint32_t i = 0; i < n && i < capacity; i++) out[i] = sizes[i];
    return n;
}

// ---- unit hooks for the known-answer tests -------------------------------------------------
__attribute__((visibility("default"))) uint64_t oracle_fxhash_u32x3(uint32_t a, uint32_t b, uint32_t c) { FxHasher h; h.write_u32(a); h.write_u32(b); h.write_u32(c); return h.finish(); }
__attribute__((visibility("default"))) uint64_t oracle_fxhash_u64(uint64_t a) { FxHasher h; h.write_u64(a); return h.finish(); }
__attribute__((visibility("default"))) void oracle_pcg32_stream(uint64_t state, uint64_t stream, uint32_t n, uint32_t* out) { Pcg32 r(state, stream); for (uint32_t i = 0; i < n; i++) out[i] = r.next_u32(); }
__attribute__((visibility("default"))) void oracle_pcg32_raw(uint64_t state, uint64_t inc, uint32_t n, uint32_t* out) { Pcg32 r; r.state = state; r.inc = inc; for (uint32_t i = 0; i < n; i++) out[i] = r.next_u32(); }
__attribute__((visibility("default"))) uint32_t oracle_permute(uint32_t i, uint32_t len, uint32_t seed) { return permute(i, len, seed); }
__attribute__((visibility("default"))) uint32_t oracle_range_u32(uint64_t state, uint64_t stream, uint32_t lo, uint32_t hi) { Pcg32 r(state, stream); return r.range_u32(lo, hi); }
__attribute__((visibility("default"))) void oracle_make_orthonormal_basis(const float* z, float* x, float* y) {
    Vec3 xx, yy;
    make_orthonormal_basis({z[0], z[1], z[2]}, xx, yy);
    x[0] = xx.x; x[1] = xx.y; x[2] = xx.z; y[0] = yy.x; y[1] = yy.y; y[2] = yy.z;
}
// out: t, u, v, nx, ny, nz ; returns hit
__attribute__((visibility("default"))) int oracle_ray_sphere(const float* center, float radius, const float* o, const float* d, float t_min, float t_max, float* out) {
    IntersectResult r;
    if (!ray_sphere_intersect({center[0], center[1], center[2]}, radius, {{o[0], o[1], o[2]}, {d[0], d[1], d[2]}}, t_min, t_max, r)) return 0;
    out[0] = r.t; out[1] = r.uv.x; out[2] = r.uv.y; out[3] = r.normal.x; out[4] = r.normal.y; out[5] = r.normal.z;
    return 1;
}
__attribute__((visibility("default"))) int oracle_ray_triangle(const float* p, const float* o, const float* d, float t_min, float t_max, float* out) {
    float t, u, v;
    if (!ray_triangle_intersect({p[0], p[1], p[2]}, {p[3], p[4], p[5]}, {p[6], p[7], p[8]}, {{o[0], o[1], o[2]}, {d[0], d[1], d[2]}}, t_min, t_max, t, u, v)) return 0;
    out[0] = t; out[1] = u; out[2] = v;
    return 1;
}
// sampler stream of one (pixel, sample): n uniform() draws
__attribute__((visibility("default"))) void oracle_sampler_stream(const rtcuda_settings* st, uint32_t x, uint32_t y, uint32_t sample, uint32_t n, float* out) {
    Sampler s = Sampler::from_settings(*st);
    s.start_sample(x, y, sample);
    for (uint32_t i = 0; i < n; i++) out[i] = s.uniform();
}
// texture lookup hook: out = rgba
__attribute__((visibility("default"))) void oracle_sample_texture(const rtcuda_scene_desc* desc, uint32_t tex, const float* uv_and_derivs, float* out) {
    SceneView sv(desc);
    MaterialEvalContext c{{uv_and_derivs[0], uv_and_derivs[1]}, uv_and_derivs[2], uv_and_derivs[3], uv_and_derivs[4], uv_and_derivs[5]};
    Vec4 v = sv.sample(tex, c);
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
// mip pyramid hook: returns level count; copies level `level` (0 = mip0) as u8/u16/f32 bytes when out != NULL
__attribute__((visibility("default"))) uint32_t oracle_mip_level(const rtcuda_scene_desc* desc, uint32_t image, uint32_t level, uint8_t* out, uint32_t* w, uint32_t* h) {
    SceneView sv(desc);
    Mipmap m = generate_mips(sv.images[image]);
    const ImageView& iv = level == 0 ? m.mip0 : m.mips[level - 1];
    if (w) *w = iv.w;
    if (h) *h = iv.h;
    if (out) std::memcpy(out, iv.data, iv.owned.size());
    return (uint32_t)m.mips.size() + 1;
}
// BSDF hook: evaluates / samples a material at uv=(0,0) with no AA. mode 0: evaluate(wo,wi) -> out[0..3];
// mode 1: sample(wo) with the path sampler of (x=0,y=0,sample=seed_idx) -> out = status, wi.xyz, f.xyz, pdf, component
__attribute__((visibility("default"))) void oracle_bsdf(const rtcuda_scene_desc* desc, const rtcuda_settings* st, uint32_t material, int mode, const float* wo, const float* wi, uint32_t seed_idx, float* out) {
    SceneView sv(desc);
    Bsdf b = get_bsdf(sv, desc->materials[material], mec_no_aa({0, 0}));
    if (mode == 0) {
        Vec3 f = b.evaluate({wo[0], wo[1], wo[2]}, {wi[0], wi[1], wi[2]});
        out[0] = f.x; out[1] = f.y; out[2] = f.z;
    } else {
        Sampler s = Sampler::from_settings(*st);
        s.start_sample(0, 0, seed_idx);
        BsdfSample bs;
        SampleStatus stt = b.sample({wo[0], wo[1], wo[2]}, ALL, s, bs);
        out[0] = (float)stt; out[1] = bs.wi.x; out[2] = bs.wi.y; out[3] = bs.wi.z; out[4] = bs.bsdf.x; out[5] = bs.bsdf.y; out[6] = bs.bsdf.z; out[7] = bs.pdf; out[8] = (float)bs.component;
    }
}

// ---- unit hooks shared with oracle/_ref (the reference's own C++ device headers compiled for the host, oracle/ref_shim/
// ref_units.cpp): same kinds, same row layouts (24 floats in, 8 floats out per row). tests/test_oracle_ref.py compares the two.
__attribute__((visibility("default"))) int oracle_unit_batch(int kind, uint32_t n, const float* in, float* out) {
    constexpr int IN_W = 24, OUT_W = 8;
    auto v3 = [](const float* p) { return Vec3{p[0], p[1], p[2]}; };
    auto put3 = [](float* o, Vec3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; };
    for (uint32_t i = 0; i < n; i++) {
        const float* a = in + (size_t)i * IN_W;
        float* o = out + (size_t)i * OUT_W;
        for (int k = 0; k < OUT_W; k++) o[k] = 0.0f;
        switch (kind) {
            case 0: o[0] = fresnel_dielectric(a[0], a[1]); break;
            case 1: o[0] = fresnel_complex(a[0], Complex{a[1], a[2]}); break;
            case 2: { Vec3 r; bool ok = refract(a[0], v3(a + 1), v3(a + 4), r); o[0] = ok ? 1.0f : 0.0f; if (ok) put3(o + 1, r); break; }
            case 3: put3(o, reflect(v3(a), v3(a + 3))); break;
            case 4: o[0] = microfacet::distribution(v3(a), a[3], a[4]); break;
            case 5: o[0] = microfacet::lambda(v3(a), a[3], a[4]); break;
            case 6: o[0] = microfacet::G1(v3(a), a[3], a[4]); break;
            case 7: o[0] = microfacet::G(v3(a), v3(a + 3), a[6], a[7]); break;
            case 8: o[0] = microfacet::visible_distribution(v3(a), v3(a + 3), a[6], a[7]); break;
            case 9: put3(o, microfacet::sample_wm(v3(a), a[3], a[4], Vec2{a[5], a[6]})); break;
            case 10: put3(o, microfacet::refl_bsdf(v3(a + 8), v3(a + 11), v3(a), v3(a + 3), a[6], a[7])); break;
            case 11: o[0] = microfacet::refl_pdf(v3(a + 8), v3(a + 11), a[6], a[7]); break;
            case 12: put3(o, microfacet::ts_bsdf(v3(a + 3), v3(a + 6), a[0], a[1], a[2])); break;
            case 13: o[0] = microfacet::ts_pdf(v3(a + 3), v3(a + 6), a[0], a[1], a[2], ALL); break;
            case 14: { Bsdf b; b.kind = B_DIFFUSE; b.albedo = v3(a); put3(o, b.evaluate(v3(a + 3), v3(a + 6))); break; }
            case 15: { Vec3 x, y; make_orthonormal_basis(v3(a), x, y); put3(o, x); put3(o + 3, y); break; }
            case 16: { Vec3 p0 = v3(a), p1 = v3(a + 3), p2 = v3(a + 6); o[0] = length(cross(p1 - p0, p2 - p0)) / 2.0f; break; }   // Mesh::tri_area, mesh.rs:271-278
            case 17: { Vec2 d = sample_unit_disk(Vec2{a[0], a[1]}); o[0] = d.x; o[1] = d.y; break; }
            case 18: { Vec2 d = sample_unit_disk_concentric(Vec2{a[0], a[1]}); o[0] = d.x; o[1] = d.y; break; }
            case 19: put3(o, sample_cosine_hemisphere(Vec2{a[0], a[1]})); break;
            case 20: o[0] = sample_exponential(a[0], a[1]); break;
            case 21: { Complex c = csqrt(Complex{a[0], a[1]}); o[0] = c.re; o[1] = c.im; break; }
            case 22: {   // SmoothConductor::sample_bsdf (materials.rs:440-466): no draw
                Bsdf b; b.kind = B_SMOOTH_CONDUCTOR; b.eta3 = v3(a); b.kappa = v3(a + 3);
                rtcuda_settings st{}; st.samples_per_pixel = 1;
                Sampler s = Sampler::from_settings(st);
                s.start_sample(0, 0, 0);
                BsdfSample bs;
                SampleStatus stt = b.sample(v3(a + 6), ALL, s, bs);
                put3(o, bs.wi); put3(o + 3, bs.bsdf); o[6] = bs.pdf; o[7] = stt == VALID ? 1.0f : 0.0f;
                break;
            }
            case 23: { Mat4 m; std::memcpy(m.m, a, sizeof m.m); put3(o, m.apply_point(v3(a + 16))); break; }
            case 24: { Mat4 m; std::memcpy(m.m, a, sizeof m.m); put3(o, m.apply_vector(v3(a + 16))); break; }
            default: return 1;
        }
    }
    return 0;
}
// camera_ray for raster positions (x, y) exactly as given (lib.rs:145-195): kind 0 orthographic, 1 pinhole. out: o.xyz, d.xyz
__attribute__((visibility("default"))) int oracle_camera_rays(int kind, const float* raster_to_camera, const float* camera_to_world, uint32_t n,
                                                              const uint32_t* xy, float* out) {
    rtcuda_camera cam{};
    cam.kind = kind == 0 ? RTCUDA_CAMERA_ORTHOGRAPHIC : RTCUDA_CAMERA_PINHOLE;
    std::memcpy(cam.raster_to_camera.forward.m, raster_to_camera, 64);
    std::memcpy(cam.camera_to_world.forward.m, camera_to_world, 64);
    for (uint32_t i = 0; i < n; i++) {
        Ray r = camera_ray(cam, (float)xy[2 * i], (float)xy[2 * i + 1], false, Vec2{0, 0});
        out[6 * (size_t)i] = r.origin.x; out[6 * (size_t)i + 1] = r.origin.y; out[6 * (size_t)i + 2] = r.origin.z;
        out[6 * (size_t)i + 3] = r.direction.x; out[6 * (size_t)i + 4] = r.direction.y; out[6 * (size_t)i + 5] = r.direction.z;
    }
    return 0;
}
}
