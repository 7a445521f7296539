// shim: <optix_types.h> — only the names the reference's shared host/device headers mention
#pragma once
typedef unsigned long long OptixTraversableHandle;
typedef unsigned int OptixVisibilityMask;
struct OptixAabb { float minX, minY, minZ, maxX, maxY, maxZ; };
#define OPTIX_SBT_RECORD_ALIGNMENT 16
#define OPTIX_SBT_RECORD_HEADER_SIZE 32
enum { OPTIX_PAYLOAD_TYPE_ID_0 = 1, OPTIX_PAYLOAD_TYPE_ID_1 = 2 };
enum { OPTIX_RAY_FLAG_NONE = 0, OPTIX_RAY_FLAG_TERMINATE_ON_FIRST_HIT = 4 };
