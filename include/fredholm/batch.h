// fredholm::FrameBatch -- multi-frame batch driver (render -> denoise -> post-process ->
// read-back -> PNG) around a Renderer.
//
// Library form of the reference's batch application app/rtcamp8.cpp:47-303: a render
// thread that, per frame, clears the six AOV layers, resets the render states, calls
// set_time(t) and render(n_spp, max_depth), denoises, post-processes, copies the image to
// the host and hands it to a saver thread that converts to RGBA8 and writes
// "<output_dir>/<frame>.png"; frames advance by 1/fps until max_time, and a wall-clock
// watchdog (kill_time) stops the batch early.
//
// B200 design: everything of one frame is enqueued on the renderer's stream; the
// float4 -> RGBA8 conversion runs on the GPU (bit-identical to the reference's host loop,
// 4x less PCIe traffic), the read-back goes through a second stream into a ring of pinned
// host slots so that it overlaps the next frame's rendering, and a pool of saver threads
// encodes / writes the PNGs.  Frames are independent, so a multi-GPU batch needs no
// collective: rank r of G runs first_frame = r, frame_stride = G.
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "fredholm/camera.h"
#include "fredholm/denoiser.h"
#include "fredholm/renderer.h"
#include "kernels/post-process.h"

namespace fredholm
{

struct BatchConfig {
  uint32_t width = 1920, height = 1080;
  uint32_t n_spp = 16, max_depth = 5;                      // rtcamp8.cpp:49-54
  PostProcessParams post{true, 2.0f, 5.0f, 80.0f, 1.0f};    // bloom 2 / 5, ISO 80, CA 1 (rtcamp8.cpp:55-58)
  bool denoise = true;                                     // false: post-process the beauty layer
  bool upscale = false;                                    // 2x output (rtcamp8.cpp:48)
  DenoiserParams denoiser;
  float fps = 24.0f;                                       // time step = 1 / fps
  float start_time = 0.0f;
  float max_time = 9.5f;                                   // frames with time > max_time are not rendered
  float kill_time_s = 590.0f;                              // wall-clock watchdog
  uint32_t first_frame = 0, frame_stride = 1;              // this rank's frames: first, first+stride, ...
  uint32_t max_frames = 0xffffffffu;                       // cap on frames rendered by this call
  float bg_color[3] = {0.0f, 0.0f, 0.0f};
  bool animate = true;                                     // call Renderer::set_time(t) every frame
  std::string output_dir = "output";                       // "" = do not write files
  uint32_t n_save_threads = 2;
  uint32_t n_slots = 3;                                    // pinned host slots in flight
  bool keep_frames = false;                                // keep the RGBA8 frames in the records
};

struct FrameRecord {
  uint32_t frame_idx = 0;
  float time = 0.0f;
  float accel_ms = 0.0f;     // set_time: animation + acceleration-structure update (host clock)
  float render_ms = 0.0f;    // device time, CUDA events
  float denoise_ms = 0.0f;
  float post_ms = 0.0f;      // post-process + RGBA8 conversion
  float transfer_ms = 0.0f;  // device -> pinned host
  float encode_ms = 0.0f;    // PNG encode (host)
  float save_ms = 0.0f;      // file write (host)
  uint64_t png_bytes = 0;
  std::string path;
  std::vector<uint8_t> rgba8;  // only with keep_frames
};

struct BatchResult {
  std::vector<FrameRecord> frames;  // in frame order
  double wall_s = 0.0;              // first enqueue -> last file written
  bool killed = false;              // watchdog fired
  uint32_t out_width = 0, out_height = 0;
};

class FrameBatch
{
 public:
  // called on the render thread before a frame is enqueued; may move the camera
  using FrameHook = std::function<void(uint32_t frame_idx, float time, Camera& camera)>;

  FrameBatch(Renderer& renderer, const BatchConfig& config);
  ~FrameBatch() noexcept(false);
  FrameBatch(const FrameBatch&) = delete;
  FrameBatch& operator=(const FrameBatch&) = delete;

  BatchResult run(const Camera& camera, const FrameHook& hook = {});

 private:
  struct Impl;
  std::unique_ptr<Impl> m_impl;
};

}  // namespace fredholm
