// TEST INFRASTRUCTURE ONLY -- never included by the product (fredholm_b200/).
//
// Minimal stand-in for <optix.h> so that the reference's OptiX device programs
// (/root/reference/fredholm/modules/pt.cu and everything it includes) compile
// *verbatim* as host C++.  It only DECLARES the 13 device-API
// entry points the reference calls; oracle_host.cpp implements them on the CPU
// (BVH traversal + program dispatch + tex2D emulation).
#pragma once

#include <cuda_runtime.h>  // float3/float4/uint3, cudaTextureObject_t
#include <sys/types.h>     // uint

#include <cmath>
#include <cstdint>
#include <cstring>

typedef unsigned long long OptixTraversableHandle;
typedef unsigned int OptixVisibilityMask;
typedef void* OptixDeviceContext;

enum OptixRayFlags {
  OPTIX_RAY_FLAG_NONE = 0u,
  OPTIX_RAY_FLAG_TERMINATE_ON_FIRST_HIT = 1u << 2,
};

#define OPTIX_SBT_RECORD_ALIGNMENT 16ull
#define OPTIX_SBT_RECORD_HEADER_SIZE ((size_t)32)

// ---- device API used by pt.cu (pt.cu:76-77, 89, 420-421, 548-556, 686-696) ----
void optixTrace(OptixTraversableHandle handle, float3 rayOrigin,
                float3 rayDirection, float tmin, float tmax, float rayTime,
                OptixVisibilityMask visibilityMask, unsigned int rayFlags,
                unsigned int SBToffset, unsigned int SBTstride,
                unsigned int missSBTIndex, unsigned int& p0, unsigned int& p1);
uint3 optixGetLaunchIndex();
uint3 optixGetLaunchDimensions();
unsigned int optixGetPayload_0();
unsigned int optixGetPayload_1();
unsigned long long optixGetSbtDataPointer();
unsigned int optixGetPrimitiveIndex();
unsigned int optixGetInstanceIndex();
float2 optixGetTriangleBarycentrics();
float3 optixGetWorldRayOrigin();
float3 optixGetWorldRayDirection();
float optixGetRayTmax();
void optixIgnoreIntersection();

// ---- CUDA device intrinsics the reference uses on the device side ----
template <typename T>
T tex2D(cudaTextureObject_t tex, float x, float y);
template <>
float4 tex2D<float4>(cudaTextureObject_t tex, float x, float y);

static inline int __float_as_int(float f)
{
  int i;
  std::memcpy(&i, &f, 4);
  return i;
}
static inline float __int_as_float(int i)
{
  float f;
  std::memcpy(&f, &i, 4);
  return f;
}
