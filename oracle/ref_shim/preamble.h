// preamble.h — the CUDA / OptiX vocabulary the reference's device headers assume, for a plain g++ host build.
//
// TEST INFRASTRUCTURE (oracle/_ref): lets oracle/ref_shim/Makefile compile the reference's own C++ restatement of the CPU
// renderer's math (/root/reference/crates/raytracing-optix/csrc/kernels/{kernel_math,materials,geometry,camera,sample}.hpp,
// every function annotated `@raytracing_cpu::...`) UNMODIFIED, where the headers lie, into oracle/_ref/libref_units.so.
// Forced in front of every translation unit with `-include`. Nothing of the reference is copied: this file only supplies
// what nvcc and the OptiX SDK would (qualifiers, vector types, intrinsics), with the host libm behind the math calls.
#pragma once
#include <math.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <type_traits>

#define __device__
#define __host__
#define __constant__
#define __forceinline__ inline
#define __align__(n) __attribute__((aligned(n)))

struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct uint2 { unsigned int x, y; };
struct uint3 { unsigned int x, y, z; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline uint2 make_uint2(unsigned int x, unsigned int y) { return uint2{x, y}; }
inline uint3 make_uint3(unsigned int x, unsigned int y, unsigned int z) { return uint3{x, y, z}; }

// CUDA's rsqrtf is a 2-ulp approximation; the host stand-in is the correctly rounded quotient (tests allow for it)
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline float __uint_as_float(unsigned int u) { float f; memcpy(&f, &u, 4); return f; }
inline unsigned int __float_as_uint(float f) { unsigned int u; memcpy(&u, &f, 4); return u; }

typedef unsigned long long cudaTextureObject_t;
template <typename T> inline T tex2D(cudaTextureObject_t, float, float) { return T{}; }   // image textures are not part of the unit comparisons
