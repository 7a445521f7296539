// shim: <optix_device.h> — the payload / launch intrinsics pathtracer.hpp wraps. None of them is reached by the unit
// comparisons (no ray is traced through OptiX here); they exist so that the reference's headers compile as they are.
#pragma once
#include "optix_types.h"
#define RT_SHIM_PAYLOAD(N) \
    inline unsigned int optixGetPayload_##N() { return 0u; } \
    inline void optixSetPayload_##N(unsigned int) {}
RT_SHIM_PAYLOAD(0) RT_SHIM_PAYLOAD(1) RT_SHIM_PAYLOAD(2) RT_SHIM_PAYLOAD(3) RT_SHIM_PAYLOAD(4) RT_SHIM_PAYLOAD(5) RT_SHIM_PAYLOAD(6) RT_SHIM_PAYLOAD(7)
RT_SHIM_PAYLOAD(8) RT_SHIM_PAYLOAD(9) RT_SHIM_PAYLOAD(10) RT_SHIM_PAYLOAD(11) RT_SHIM_PAYLOAD(12) RT_SHIM_PAYLOAD(13) RT_SHIM_PAYLOAD(14)
#undef RT_SHIM_PAYLOAD
template <typename... Args> inline void optixTrace(Args&&...) {}
inline uint3 optixGetLaunchDimensions() { return uint3{1, 1, 1}; }
inline uint3 optixGetLaunchIndex() { return uint3{0, 0, 0}; }
