// Exercises the cwl mirror (include/cwl/{buffer,util,texture}.h) the way the reference's applications use it:
// construct / fill / clear / upload / read back CUDABuffer<T>, a DeviceObject<T>, a CUDATexture<uchar4>.
// Prints "ok" and returns 0; any mismatch or CUDA error is an exception (exit code 1).
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "cwl/buffer.h"
#include "cwl/texture.h"
#include "cwl/util.h"

#define REQUIRE(x) \
  if (!(x)) throw std::runtime_error("check failed: " #x)

int main()
{
  try {
    CUDA_CHECK(cudaFree(0));
    // value constructor: every 32-bit word (buffer.h:27-33)
    cwl::CUDABuffer<float4> a(1000, 0x3f800000u);
    std::vector<float4> h;
    a.copy_from_device_to_host(h);
    REQUIRE(h.size() == 1000 && h[0].x == 1.0f && h[999].w == 1.0f);
    REQUIRE(a.get_size() == 1000 && a.get_size_in_bytes() == 16000);
    a.clear();
    a.copy_from_device_to_host(h);
    REQUIRE(h[500].y == 0.0f);
    // vector constructor + move
    std::vector<float> v(257);
    for (size_t i = 0; i < v.size(); ++i) v[i] = float(i);
    cwl::CUDABuffer<float> b(v);
    cwl::CUDABuffer<float> c(std::move(b));
    REQUIRE(b.get_size() == 0 && b.get_device_ptr() == nullptr && c.get_size() == 257);
    std::vector<float> back;
    c.copy_from_device_to_host(back);
    REQUIRE(back == v);
    REQUIRE(c.get_const_device_ptr() == c.get_device_ptr());
    cwl::CUDABuffer<float> empty(0);
    empty.clear();
    empty.copy_from_device_to_host(back);
    REQUIRE(back.empty());
    // one object on the device
    struct P {
      int a;
      float b[3];
    } p{7, {1.f, 2.f, 3.f}}, q{};
    cwl::DeviceObject<P> dp(p);
    CUDA_CHECK(cudaMemcpy(&q, dp.get_device_ptr(), sizeof(P), cudaMemcpyDeviceToHost));
    REQUIRE(std::memcmp(&p, &q, sizeof(P)) == 0);
    // texture object with the reference's sampler state
    std::vector<uchar4> texels(16 * 8, make_uchar4(255, 128, 0, 255));
    cwl::CUDATexture<uchar4> tex(16, 8, texels.data(), /*srgb_to_linear=*/true);
    REQUIRE(tex.get_texture_object() != 0 && tex.get_size().x == 16 && tex.get_size().y == 8);
    cudaTextureDesc td;
    CUDA_CHECK(cudaGetTextureObjectTextureDesc(&td, tex.get_texture_object()));
    REQUIRE(td.addressMode[0] == cudaAddressModeWrap && td.filterMode == cudaFilterModeLinear &&
            td.readMode == cudaReadModeNormalizedFloat && td.normalizedCoords == 1 && td.sRGB == 1);
    // error text of a failing call: the reference's format (cwl/util.h:11-22)
    bool threw = false;
    try {
      CUDA_CHECK(cudaMemcpy(nullptr, nullptr, 16, cudaMemcpyDeviceToHost));
    } catch (const std::runtime_error& e) {
      threw = std::strstr(e.what(), "CUDA call (cudaMemcpy") != nullptr && std::strstr(e.what(), "cwl_check.cpp") != nullptr;
      cudaGetLastError();
    }
    REQUIRE(threw);
    CUDA_SYNC_CHECK();
    std::puts("ok");
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 1;
  }
}
