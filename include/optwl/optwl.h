// optwl/optwl.h of the B200 core: source compatibility for applications written against the reference's OptiX
// context wrapper (reference optwl/include/optwl/optwl.h:43-92).  There is no OptiX here -- the traversal is a
// hand-written CUDA kernel -- so the "context" an application creates and hands to fredholm::Renderer /
// fredholm::Denoiser is just the CUDA device index:
//     optwl::Context context;                       // rtcamp8.cpp:72, controller.cpp:10
//     fredholm::Renderer renderer(context.m_context);
// compiles unchanged; OptixDeviceContext is an int.
#pragma once
#include <cuda_runtime.h>

typedef int OptixDeviceContext;  // CUDA device index

namespace optwl
{

struct Context {
  OptixDeviceContext m_context = 0;
  // the reference takes a CUcontext (0 = current); here: the CUDA device to render on
  explicit Context(int cuda_device = 0) : m_context(cuda_device) {}
};

}  // namespace optwl
