// NVTX ranges around the wavefront stages: an nsys / ncu timeline then shows
// generate / trace_closest / shade / trace_shadow / trace_light / film per bounce next to the kernels, lined
// up with Renderer::get_stage_times.  Header-only NVTX v3 (ships with the CUDA toolkit, no library to link);
// without a profiler attached a range costs a few nanoseconds.  -DFRD_NO_NVTX compiles them out.
#pragma once
#if !defined(FRD_NO_NVTX) && __has_include(<nvtx3/nvToolsExt.h>)
#include <nvtx3/nvToolsExt.h>
namespace frd
{
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
}  // namespace frd
#define FR_NVTX_CAT2(a, b) a##b
#define FR_NVTX_CAT(a, b) FR_NVTX_CAT2(a, b)
#define FR_NVTX_RANGE(name) ::frd::NvtxRange FR_NVTX_CAT(fr_nvtx_range_, __LINE__)(name)
#define FR_HAVE_NVTX 1
#else
#define FR_NVTX_RANGE(name) ((void)0)
#define FR_HAVE_NVTX 0
#endif
