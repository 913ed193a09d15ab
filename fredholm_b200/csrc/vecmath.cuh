// Small float3 toolkit for the wavefront kernels.
//
// Semantics follow the helpers the reference integrator is written against
// (externals/sutil/sutil/vec_math.h in the reference tree) where they change
// rounding: normalize() multiplies by 1/sqrt(dot) (vec_math.h:568-572) and
// division by a scalar multiplies by the reciprocal (vec_math.h:498-511).
#pragma once
#include <cuda_runtime.h>

#define FR_HD __host__ __device__ __forceinline__
#define FR_D __device__ __forceinline__

namespace frd
{

constexpr float kPi = 3.14159265358979323846f;
constexpr float kInvPi = 0.318309886183790671538f;

FR_HD float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
FR_HD float3 f3(float s) { return make_float3(s, s, s); }
FR_HD float3 f3(const float4& v) { return make_float3(v.x, v.y, v.z); }

FR_HD float3 operator+(const float3& a, const float3& b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
FR_HD float3 operator-(const float3& a, const float3& b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
FR_HD float3 operator*(const float3& a, const float3& b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
FR_HD float3 operator/(const float3& a, const float3& b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
FR_HD float3 operator*(const float3& a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
FR_HD float3 operator*(float s, const float3& a) { return f3(a.x * s, a.y * s, a.z * s); }
FR_HD float3 operator+(const float3& a, float s) { return f3(a.x + s, a.y + s, a.z + s); }
FR_HD float3 operator+(float s, const float3& a) { return f3(a.x + s, a.y + s, a.z + s); }
FR_HD float3 operator-(const float3& a, float s) { return f3(a.x - s, a.y - s, a.z - s); }
FR_HD float3 operator-(float s, const float3& a) { return f3(s - a.x, s - a.y, s - a.z); }
FR_HD float3 operator/(float s, const float3& a) { return f3(s / a.x, s / a.y, s / a.z); }
FR_HD float3 operator/(const float3& a, float s)
{
  const float inv = 1.0f / s;
  return a * inv;
}
FR_HD float3 operator-(const float3& a) { return f3(-a.x, -a.y, -a.z); }
FR_HD void operator+=(float3& a, const float3& b) { a = a + b; }
FR_HD void operator*=(float3& a, const float3& b) { a = a * b; }
FR_HD void operator*=(float3& a, float s) { a = a * s; }

FR_HD float2 operator+(const float2& a, const float2& b) { return make_float2(a.x + b.x, a.y + b.y); }
FR_HD float2 operator*(float s, const float2& a) { return make_float2(a.x * s, a.y * s); }

FR_HD float dot(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
FR_HD float3 cross(const float3& a, const float3& b)
{
  return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
FR_HD float length(const float3& v) { return sqrtf(dot(v, v)); }
FR_HD float3 normalize(const float3& v)
{
  const float inv = 1.0f / sqrtf(dot(v, v));
  return v * inv;
}
FR_HD float clampf(float f, float a, float b) { return fmaxf(a, fminf(f, b)); }
// clamp(x, 0, 1) with the DEVICE semantics of the reference: nvcc lowers sutil's fmaxf(0, fminf(x, 1)) to the
// .sat modifier, which maps NaN to 0 (oracle/Makefile, `make sat-evidence`).  __saturatef states that
// explicitly instead of relying on the compiler's pattern match.
#ifdef __CUDACC__
FR_D float saturate1(float x) { return __saturatef(x); }
#else
inline float saturate1(float x) { return x != x ? 0.0f : fmaxf(0.0f, fminf(x, 1.0f)); }
#endif
FR_D float3 saturate3(const float3& v) { return make_float3(saturate1(v.x), saturate1(v.y), saturate1(v.z)); }
FR_HD float3 clamp3(const float3& v, float a, float b)
{
  return f3(clampf(v.x, a, b), clampf(v.y, a, b), clampf(v.z, a, b));
}
FR_HD float3 max3(const float3& a, const float3& b) { return f3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
FR_HD float3 sqrt3(const float3& v) { return f3(sqrtf(v.x), sqrtf(v.y), sqrtf(v.z)); }
FR_HD bool any_nan(const float3& v) { return isnan(v.x) || isnan(v.y) || isnan(v.z); }
FR_HD bool any_inf(const float3& v) { return isinf(v.x) || isinf(v.y) || isinf(v.z); }
FR_HD bool bad3(const float3& v) { return any_nan(v) || any_inf(v); }

// reference math.cu:90-93 (Lindbloom sRGB->Y row)
FR_HD float luminance(const float3& rgb)
{
  return dot(rgb, f3(0.2126729f, 0.7151522f, 0.0721750f));
}

// Duff et al. 2017 branchless ONB, as used by the reference (math.cu:7-17)
FR_HD void onb(const float3& n, float3& t, float3& b)
{
  const float sign = copysignf(1.0f, n.z);
  const float a = -1.0f / (sign + n.z);
  const float c = n.x * n.y * a;
  t = f3(1.0f + sign * n.x * n.x * a, sign * c, -sign * n.x);
  b = f3(c, sign + n.y * n.y * a, -n.y);
}

// shading frame: local = (dot(v,t), dot(v,n), dot(v,b)), y is up (math.cu:19-35)
struct Frame {
  float3 t, n, b;
  FR_HD float3 to_local(const float3& v) const { return f3(dot(v, t), dot(v, n), dot(v, b)); }
  FR_HD float3 to_world(const float3& v) const
  {
    return f3(v.x * t.x + v.y * n.x + v.z * b.x, v.x * t.y + v.y * n.y + v.z * b.y,
              v.x * t.z + v.y * n.z + v.z * b.z);
  }
};

}  // namespace frd
