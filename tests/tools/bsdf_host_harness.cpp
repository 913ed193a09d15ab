// TEST TOOLING ONLY (never part of libfredholm_b200.so): compiles the device BSDF header csrc/bsdf.cuh for the
// HOST, so that its control flow can be compared with the reference's BSDF (oracle/_ref) on this CPU-only
// container -- in particular for directions the GPU golden table does not contain (wo below the shading
// horizon, which normal maps produce).  Arithmetic differs from the device build in the last bits (no
// -use_fast_math, glibc libm); structural differences (NaN vs value, zero vs value) do not.
#define __device__
#define __host__
#define __global__
#define __constant__
#define __forceinline__ inline
#include <algorithm>
#include <cmath>
#include <cstdint>
using std::isinf;
using std::isnan;
using std::max;
using std::min;
#include "tables.cuh"
//
#include "bsdf.cuh"

using namespace frd;

extern "C" void host_bsdf_eval_sample(const float* in, uint32_t n, float* out)
{
  for (uint32_t i = 0; i < n; ++i) {
    const float* p = in + 40ull * i;
    SurfaceParams s;
    s.diffuse = p[0];
    s.base_color = f3(p[1], p[2], p[3]);
    s.diffuse_roughness = p[4];
    s.specular = p[5];
    s.specular_color = f3(p[6], p[7], p[8]);
    s.specular_roughness = p[9];
    s.metalness = p[10];
    s.coat = p[11];
    s.coat_color = f3(p[12], p[13], p[14]);
    s.coat_roughness = p[15];
    s.transmission = p[16];
    s.transmission_color = f3(p[17], p[18], p[19]);
    s.sheen = p[20];
    s.sheen_color = f3(p[21], p[22], p[23]);
    s.sheen_roughness = p[24];
    s.subsurface = p[25];
    s.subsurface_color = f3(p[26], p[27], p[28]);
    s.thin_walled = p[29];
    const float3 wo = f3(p[30], p[31], p[32]);
    const bool entering = p[33] != 0.0f;
    const float3 wi = f3(p[34], p[35], p[36]);
    Closure<M_ALL> b;
    b.init(wo, s, entering);
    float3 f;
    float pdf;
    b.eval(wi, f, pdf);
    float* o = out + 11ull * i;
    o[0] = f.x, o[1] = f.y, o[2] = f.z, o[3] = pdf;
    float3 fs;
    float pdfs;
    const float3 ws = b.sample(p[37], make_float2(p[38], p[39]), fs, pdfs);
    o[4] = ws.x, o[5] = ws.y, o[6] = ws.z, o[7] = fs.x, o[8] = fs.y, o[9] = fs.z, o[10] = pdfs;
  }
}
