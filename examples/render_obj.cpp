// Minimal C++ application against the fredholm::Renderer mirror (include/fredholm/renderer.h):
// the same call sequence the reference's apps use (app/controller.cpp:60-68,126-134;
// app/rtcamp8.cpp:79-124,186-215), then post-process and read-back.
//
//   render_obj scene.obj out.ppm [width height spp depth]
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <vector>

#include "fredholm/renderer.h"
#include "kernels/post-process.h"

static void check(cudaError_t e)
{
  if (e != cudaSuccess) throw std::runtime_error(cudaGetErrorString(e));
}

int main(int argc, char** argv)
{
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s scene.obj out.ppm [width height spp depth]\n", argv[0]);
    return 2;
  }
  const int width = argc > 3 ? std::atoi(argv[3]) : 512, height = argc > 4 ? std::atoi(argv[4]) : 512;
  const int spp = argc > 5 ? std::atoi(argv[5]) : 16, depth = argc > 6 ? std::atoi(argv[6]) : 5;
  try {
    fredholm::Renderer renderer(0);
    renderer.create_module("pt.ptx");
    renderer.create_program_group();
    renderer.create_pipeline();
    renderer.set_resolution(width, height);
    renderer.load_scene(argv[1]);
    renderer.build_gas();
    renderer.build_ias();
    renderer.create_sbt();
    renderer.set_directional_light(make_float3(20, 20, 20), make_float3(-0.1f, 1.0f, 0.1f), 1.0f);
    renderer.load_arhosek_sky(3.0f, 0.3f);

    const size_t n = (size_t)width * height;
    float4 *beauty, *high, *temp, *out;
    for (float4** p : {&beauty, &high, &temp, &out}) {
      check(cudaMalloc(p, n * sizeof(float4)));
      check(cudaMemset(*p, 0, n * sizeof(float4)));
    }
    fredholm::RenderLayer layers{};
    layers.beauty = beauty;

    fredholm::Camera camera(make_float3(0.0f, 1.0f, 5.0f));
    renderer.init_render_states();
    renderer.render(camera, make_float3(0, 0, 0), layers, spp, depth);
    renderer.wait_for_completion();

    PostProcessParams pp{true, 2.0f, 5.0f, 80.0f, 1.0f};
    post_process_kernel_launch(beauty, high, temp, width, height, pp, out);
    std::vector<float4> host(n);
    check(cudaMemcpy(host.data(), out, n * sizeof(float4), cudaMemcpyDeviceToHost));

    const fredholm::RenderStatistics st = renderer.get_statistics();
    std::printf("%llu paths, %llu rays, %llu kernel launches\n", st.paths, st.total_rays(), st.kernel_launches);
    FILE* f = std::fopen(argv[2], "wb");
    if (!f) throw std::runtime_error("cannot write output");
    std::fprintf(f, "P6\n%d %d\n255\n", width, height);
    for (const float4& c : host) {
      const unsigned char px[3] = {(unsigned char)std::clamp(255.0f * c.x, 0.0f, 255.0f),
                                   (unsigned char)std::clamp(255.0f * c.y, 0.0f, 255.0f),
                                   (unsigned char)std::clamp(255.0f * c.z, 0.0f, 255.0f)};
      std::fwrite(px, 1, 3, f);
    }
    std::fclose(f);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
