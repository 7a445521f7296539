// rt_common.h — vocabulary shared by every kernel of libraytracing_cuda: vector math with the
// reference's operation order, bit casts, and the host/device portability shims that let the
// per-thread kernel bodies also be compiled by g++ for the CPU test harness (tests/hostsim —
// test infrastructure, never linked into the product library).
//
// Math conventions follow crates/raytracing/src/geometry/{vec2,vec3,vec4}.rs: `v / s` is
// `v * (1.0 / s)` (vec3.rs:122-126,179-184), `unit(v) = v / length(v)`.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define RT_HD __host__ __device__ __forceinline__
#define RT_D __device__ __forceinline__
// heavy, multiply-called bodies (BSDF walks, texture fetch): real calls keep cicc's inlining tractable
#define RT_HD_CALL __host__ __device__ __noinline__ inline
#else
#define RT_HD inline
#define RT_D inline
#define RT_HD_CALL inline
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
struct alignas(8) uint2 { uint32_t x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
#endif

// Bounds checks of the queue / stack indices for debug builds (`make NVCCFLAGS_EXTRA=-DRT_DEBUG_CHECKS`): the pool's
// compute-sanitizer is closed, so the invariants are asserted by the kernels themselves and the GPU suite is run once
// per round with the checks on (profiles/r1_notes.md).
#if defined(RT_DEBUG_CHECKS)
#include <cassert>
#define RT_CHECK(cond) assert(cond)
#else
#define RT_CHECK(cond) ((void)0)
#endif

namespace rt {

constexpr float PI = 3.14159265358979323846f;
constexpr float FRAC_1_PI = 0.318309886183790671537767526745028724f;
constexpr float FRAC_PI_2 = 1.57079632679489661923132169163975144f;
constexpr float FRAC_PI_4 = 0.785398163397448309615660845819875721f;
constexpr uint32_t NONE = 0xffffffffu;

#if defined(__CUDA_ARCH__)
#define RT_INF __int_as_float(0x7f800000)
RT_HD uint32_t f2u(float f) { return __float_as_uint(f); }
RT_HD float u2f(uint32_t u) { return __uint_as_float(u); }
RT_HD int clz32(uint32_t v) { return __clz((int)v); }
RT_HD int clz64(uint64_t v) { return __clzll((long long)v); }
RT_HD int popc32(uint32_t v) { return __popc(v); }
RT_HD int bfind32(uint32_t v) { return 31 - __clz((int)v); }  // index of highest set bit (v != 0)
template <typename T> RT_HD T ldg(const T* p) { return __ldg(p); }
#else
#define RT_INF (__builtin_inff())
RT_HD uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
RT_HD float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
RT_HD int clz32(uint32_t v) { return v ? __builtin_clz(v) : 32; }
RT_HD int clz64(uint64_t v) { return v ? __builtin_clzll(v) : 64; }
RT_HD int popc32(uint32_t v) { return __builtin_popcount(v); }
RT_HD int bfind32(uint32_t v) { return 31 - __builtin_clz(v); }
template <typename T> RT_HD T ldg(const T* p) { return *p; }
#endif

struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

RT_HD V2 mk2(float x, float y) { V2 r; r.x = x; r.y = y; return r; }
RT_HD V3 mk3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
RT_HD V3 mk3(float s) { return mk3(s, s, s); }
RT_HD V4 mk4(float x, float y, float z, float w) { V4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
RT_HD V3 xyz(float4 v) { return mk3(v.x, v.y, v.z); }
RT_HD V3 xyz(V4 v) { return mk3(v.x, v.y, v.z); }

RT_HD V2 operator+(V2 a, V2 b) { return mk2(a.x + b.x, a.y + b.y); }
RT_HD V2 operator-(V2 a, V2 b) { return mk2(a.x - b.x, a.y - b.y); }
RT_HD V2 operator*(V2 a, float s) { return mk2(a.x * s, a.y * s); }
RT_HD V2 operator*(float s, V2 a) { return a * s; }
RT_HD float sqmag(V2 a) { return a.x * a.x + a.y * a.y; }

RT_HD V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
RT_HD V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
RT_HD V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
RT_HD V3 operator*(V3 a, V3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
RT_HD V3 operator*(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
RT_HD V3 operator*(float s, V3 a) { return a * s; }
RT_HD V3 operator/(V3 a, float s) { return a * (1.0f / s); }
RT_HD V3& operator+=(V3& a, V3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
RT_HD V3& operator*=(V3& a, V3 b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; return a; }
RT_HD V3& operator*=(V3& a, float s) { a.x *= s; a.y *= s; a.z *= s; return a; }
RT_HD V3& operator/=(V3& a, float s) { a *= (1.0f / s); return a; }
RT_HD bool operator==(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
RT_HD bool operator!=(V3 a, V3 b) { return !(a == b); }
RT_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
RT_HD V3 cross(V3 u, V3 v) { return mk3(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x); }
RT_HD float sqmag(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
RT_HD float length(V3 a) { return sqrtf(sqmag(a)); }
RT_HD V3 unit(V3 a) { return a / length(a); }
RT_HD float max_component(V3 a) { return fmaxf(a.x, fmaxf(a.y, a.z)); }
RT_HD V3 vmin(V3 a, V3 b) { return mk3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
RT_HD V3 vmax(V3 a, V3 b) { return mk3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
RT_HD V3 reflect(V3 v, V3 n) { return -v + 2.0f * dot(v, n) * n; }  // vec3.rs:93-95
RT_HD bool is_zero(V3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }

RT_HD V4 operator+(V4 a, V4 b) { return mk4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
RT_HD V4 operator-(V4 a, V4 b) { return mk4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
RT_HD V4 operator*(V4 a, V4 b) { return mk4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
RT_HD V4 operator*(V4 a, float s) { return mk4(a.x * s, a.y * s, a.z * s, a.w * s); }
RT_HD V4 operator*(float s, V4 a) { return a * s; }
RT_HD bool operator==(V4 a, V4 b) { return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w; }

// Rust f32 helpers with their exact semantics.
RT_HD float rs_clamp(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }  // NaN stays NaN
RT_HD float rs_fract(float v) { return v - truncf(v); }
RT_HD float rs_signum(float v) { return v != v ? v : copysignf(1.0f, v); }
RT_HD uint32_t rs_as_u32(float v) {  // saturating `as u32`
    if (!(v > 0.0f)) return 0;
    if (v >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)v;
}
RT_HD int32_t rs_as_i32(float v) {
    if (v != v) return 0;
    if (v <= -2147483648.0f) return INT32_MIN;
    if (v >= 2147483648.0f) return INT32_MAX;
    return (int32_t)v;
}
RT_HD bool finite_f(float f) { return (f2u(f) & 0x7f800000u) != 0x7f800000u; }

// Row-major 4x4 (crates/raytracing/src/geometry/matrix4x4.rs:326-359)
struct M4 { float m[16]; };
RT_HD V3 apply_point(const M4& a, V3 p) {
    float x = a.m[0] * p.x + a.m[1] * p.y + a.m[2] * p.z + a.m[3] * 1.0f;
    float y = a.m[4] * p.x + a.m[5] * p.y + a.m[6] * p.z + a.m[7] * 1.0f;
    float z = a.m[8] * p.x + a.m[9] * p.y + a.m[10] * p.z + a.m[11] * 1.0f;
    float w = a.m[12] * p.x + a.m[13] * p.y + a.m[14] * p.z + a.m[15] * 1.0f;
    return mk3(x / w, y / w, z / w);
}
RT_HD V3 apply_vector(const M4& a, V3 v) {
    return mk3(a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[4] * v.x + a.m[5] * v.y + a.m[6] * v.z,
               a.m[8] * v.x + a.m[9] * v.y + a.m[10] * v.z);
}
// rows 0..2 of a matrix that lives in device memory at a 16-byte aligned address (Instance arrays), as three 128-bit loads
RT_HD void load_rows3(const M4& a, float4& r0, float4& r1, float4& r2) {
#if defined(__CUDA_ARCH__)
    const float4* p = reinterpret_cast<const float4*>(a.m);
    r0 = __ldg(p); r1 = __ldg(p + 1); r2 = __ldg(p + 2);
#else
    r0 = make_float4(a.m[0], a.m[1], a.m[2], a.m[3]); r1 = make_float4(a.m[4], a.m[5], a.m[6], a.m[7]); r2 = make_float4(a.m[8], a.m[9], a.m[10], a.m[11]);
#endif
}
// four consecutive 32-bit fields at a 16-byte aligned device address
RT_HD uint4 load_u4(const uint32_t* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(reinterpret_cast<const uint4*>(p));
#else
    return make_uint4(p[0], p[1], p[2], p[3]);
#endif
}
// Streaming (evict-first) accesses for the wavefront queues: every entry is written once and read once by another kernel, so it
// should not displace the BVH nodes, primitive records and scene tables the kernels re-read — measured: C3 +0.3 %, C4 and the
// rough-metal box -2 % (profiles/r5g_ab.log), so the hints stay off (A/B switch RT_STREAM_HINTS).
#ifndef RT_STREAM_HINTS
#define RT_STREAM_HINTS 0
#endif
RT_HD float4 ld_stream(const float4* p) {
#if defined(__CUDA_ARCH__) && RT_STREAM_HINTS
    return __ldcs(p);
#else
    return *p;
#endif
}
RT_HD void st_stream(float4* p, float4 v) {
#if defined(__CUDA_ARCH__) && RT_STREAM_HINTS
    __stcs(p, v);
#else
    *p = v;
#endif
}
RT_HD V3 apply_vector_transposed(const M4& a, V3 v) {
    return mk3(a.m[0] * v.x + a.m[4] * v.y + a.m[8] * v.z, a.m[1] * v.x + a.m[5] * v.y + a.m[9] * v.z,
               a.m[2] * v.x + a.m[6] * v.y + a.m[10] * v.z);
}

// geometry.rs:8-20
RT_HD void make_orthonormal_basis(V3 z, V3& x, V3& y) {
    V3 a = fabsf(z.z) < 0.8f ? mk3(0, 0, 1) : mk3(0, 1, 0);
    x = unit(cross(a, z));
    y = cross(z, x);
}

struct Frame {  // Matrix4x4::create_from_basis(x, y, n) and its transpose (lib.rs:311-316)
    V3 x, y, n;
    RT_HD V3 to_local(V3 v) const { return mk3(dot(x, v), dot(y, v), dot(n, v)); }
    RT_HD V3 to_world(V3 v) const {
        return mk3(x.x * v.x + y.x * v.y + n.x * v.z, x.y * v.x + y.y * v.y + n.y * v.z, x.z * v.x + y.z * v.y + n.z * v.z);
    }
};

}  // namespace rt

// Atomics used by the builder / queues. On the device these are the hardware atomics; the CPU test
// harness runs kernel bodies sequentially, where a plain read-modify-write is equivalent.
namespace rt {
#if defined(__CUDA_ARCH__)
RT_HD uint32_t atomic_add_u32(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
RT_HD uint32_t atomic_min_u32(uint32_t* p, uint32_t v) { return atomicMin(p, v); }
RT_HD uint32_t atomic_max_u32(uint32_t* p, uint32_t v) { return atomicMax(p, v); }
RT_HD void mem_fence() { __threadfence(); }
#else
RT_HD uint32_t atomic_add_u32(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
RT_HD uint32_t atomic_min_u32(uint32_t* p, uint32_t v) { uint32_t o = *p; if (v < o) *p = v; return o; }
RT_HD uint32_t atomic_max_u32(uint32_t* p, uint32_t v) { uint32_t o = *p; if (v > o) *p = v; return o; }
RT_HD void mem_fence() {}
#endif
// order-preserving float <-> u32 key (for atomic min / max of floats)
RT_HD uint32_t float_key(float f) { uint32_t b = f2u(f); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
RT_HD float key_float(uint32_t k) { return u2f((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }
}  // namespace rt
