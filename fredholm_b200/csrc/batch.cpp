// fredholm::FrameBatch: pipelined multi-frame driver.  Behavioural spec: app/rtcamp8.cpp:47-303
// of the reference (see include/fredholm/batch.h).
//
//   render stream : clear layers -> render -> denoise -> post-process -> RGBA8   (frame k)
//   copy stream   :                     wait(ready[k]) -> D2H into pinned slot   (frame k)
//   saver threads :                                  sync(copied[k]) -> PNG -> file
// A slot (device RGBA8 image + pinned host image + events) is reused only after its saver is
// done, so up to n_slots frames are in flight; the AOV layers exist once because the render
// stream serialises their use.
#include "fredholm/batch.h"

#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <exception>
#include <filesystem>
#include <fstream>
#include <mutex>
#include <thread>

#include "cuda_util.h"
#include "image_codec.h"
#include "post_process_async.h"

namespace fredholm
{

namespace
{
using Clock = std::chrono::steady_clock;
double seconds_since(const Clock::time_point& t0)
{
  return std::chrono::duration<double>(Clock::now() - t0).count();
}
float ms_since(const Clock::time_point& t0) { return (float)(1e3 * seconds_since(t0)); }

struct Slot {
  frd::DevBuf<uchar4> d_rgba8;
  uint8_t* h_rgba8 = nullptr;  // pinned
  cudaEvent_t e_start = nullptr, e_render = nullptr, e_denoise = nullptr, e_post = nullptr;
  cudaEvent_t e_copy0 = nullptr, e_copy1 = nullptr;
  bool busy = false;
};
}  // namespace

struct FrameBatch::Impl {
  Renderer& renderer;
  BatchConfig cfg;
  uint32_t out_w, out_h;
  cudaStream_t render_stream = nullptr, copy_stream = nullptr;
  frd::DevBuf<float4> beauty, position, normal, texcoord, albedo, denoised, pp_out, high_lum, temp;
  frd::DevBuf<float> depth;
  std::unique_ptr<Denoiser> denoiser;
  std::vector<Slot> slots;

  Impl(Renderer& r, const BatchConfig& c) : renderer(r), cfg(c)
  {
    if (cfg.width == 0 || cfg.height == 0) throw std::runtime_error("FrameBatch: empty image");
    if (cfg.n_spp == 0) throw std::runtime_error("FrameBatch: n_spp must be positive");
    if (cfg.frame_stride == 0) throw std::runtime_error("FrameBatch: frame_stride must be positive");
    if (!(cfg.fps > 0.0f)) throw std::runtime_error("FrameBatch: fps must be positive");
    if (cfg.upscale && !cfg.denoise) throw std::runtime_error("FrameBatch: upscale needs the denoise stage");
    cfg.n_slots = std::max(cfg.n_slots, 1u);
    cfg.n_save_threads = std::max(cfg.n_save_threads, 1u);
    FR_CUDA_CHECK(cudaSetDevice(renderer.get_device()));
    render_stream = renderer.get_stream();
    FR_CUDA_CHECK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    out_w = cfg.upscale ? 2 * cfg.width : cfg.width;
    out_h = cfg.upscale ? 2 * cfg.height : cfg.height;
    const size_t n = (size_t)cfg.width * cfg.height, n_out = (size_t)out_w * out_h;
    beauty.alloc(n);
    position.alloc(n);
    normal.alloc(n);
    texcoord.alloc(n);
    albedo.alloc(n);
    depth.alloc(n);
    pp_out.alloc(n_out);
    high_lum.alloc(n_out);
    temp.alloc(n_out);
    // the post-process grid quirk leaves border pixels unwritten (post-process.cu:9-11): the
    // reference's buffers are zero-initialised (cwl/buffer.h:25-30), so are these
    pp_out.zero(render_stream);
    high_lum.zero(render_stream);
    temp.zero(render_stream);
    if (cfg.denoise) {
      denoised.alloc(n_out);
      denoiser = std::make_unique<Denoiser>(cfg.width, cfg.height, beauty.get(), normal.get(), albedo.get(),
                                            denoised.get(), cfg.upscale, render_stream);
      denoiser->set_params(cfg.denoiser);
    }
    renderer.set_resolution(cfg.width, cfg.height);
    slots.resize(cfg.n_slots);
    for (Slot& s : slots) {
      s.d_rgba8.alloc(n_out);
      FR_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&s.h_rgba8), n_out * 4, cudaHostAllocDefault));
      for (cudaEvent_t* e : {&s.e_start, &s.e_render, &s.e_denoise, &s.e_post, &s.e_copy0, &s.e_copy1})
        FR_CUDA_CHECK(cudaEventCreate(e));
    }
    FR_CUDA_CHECK(cudaStreamSynchronize(render_stream));
  }

  ~Impl()
  {
    cudaSetDevice(renderer.get_device());
    if (copy_stream) cudaStreamSynchronize(copy_stream);
    for (Slot& s : slots) {
      if (s.h_rgba8) cudaFreeHost(s.h_rgba8);
      for (cudaEvent_t e : {s.e_start, s.e_render, s.e_denoise, s.e_post, s.e_copy0, s.e_copy1})
        if (e) cudaEventDestroy(e);
    }
    if (copy_stream) cudaStreamDestroy(copy_stream);
  }

  RenderLayer layers() const
  {
    RenderLayer l;
    l.beauty = beauty.get();
    l.position = position.get();
    l.normal = normal.get();
    l.depth = depth.get();
    l.texcoord = texcoord.get();
    l.albedo = albedo.get();
    return l;
  }

  void enqueue_frame(Slot& s, const Camera& camera)
  {
    // clear render layers + render states (rtcamp8.cpp:170-178)
    beauty.zero(render_stream);
    position.zero(render_stream);
    normal.zero(render_stream);
    depth.zero(render_stream);
    texcoord.zero(render_stream);
    albedo.zero(render_stream);
    renderer.init_render_states();
    FR_CUDA_CHECK(cudaEventRecord(s.e_start, render_stream));
    renderer.render(camera, make_float3(cfg.bg_color[0], cfg.bg_color[1], cfg.bg_color[2]), layers(), cfg.n_spp,
                    cfg.max_depth);
    FR_CUDA_CHECK(cudaEventRecord(s.e_render, render_stream));
    if (denoiser) denoiser->denoise();
    FR_CUDA_CHECK(cudaEventRecord(s.e_denoise, render_stream));
    const float4* src = denoiser ? denoised.get() : beauty.get();
    frd::post_process_async(src, high_lum.get(), temp.get(), (int)out_w, (int)out_h, cfg.post, pp_out.get(),
                            render_stream);
    frd::quantize_rgba8_async(pp_out.get(), (size_t)out_w * out_h, s.d_rgba8.get(), render_stream);
    FR_CUDA_CHECK(cudaEventRecord(s.e_post, render_stream));
    // read-back on the copy stream, overlapping the next frame
    FR_CUDA_CHECK(cudaStreamWaitEvent(copy_stream, s.e_post, 0));
    FR_CUDA_CHECK(cudaEventRecord(s.e_copy0, copy_stream));
    FR_CUDA_CHECK(cudaMemcpyAsync(s.h_rgba8, s.d_rgba8.get(), (size_t)out_w * out_h * 4, cudaMemcpyDeviceToHost,
                                  copy_stream));
    FR_CUDA_CHECK(cudaEventRecord(s.e_copy1, copy_stream));
  }
};

FrameBatch::FrameBatch(Renderer& renderer, const BatchConfig& config) : m_impl(std::make_unique<Impl>(renderer, config))
{
}
FrameBatch::~FrameBatch() noexcept(false) {}

BatchResult FrameBatch::run(const Camera& camera_in, const FrameHook& hook)
{
  Impl& d = *m_impl;
  const BatchConfig& cfg = d.cfg;
  FR_CUDA_CHECK(cudaSetDevice(d.renderer.get_device()));
  if (!cfg.output_dir.empty()) std::filesystem::create_directories(cfg.output_dir);

  BatchResult result;
  result.out_width = d.out_w;
  result.out_height = d.out_h;
  std::deque<FrameRecord> records;  // stable addresses while savers fill them

  struct Job {
    size_t slot;
    FrameRecord* rec;
  };
  std::mutex mu;
  std::condition_variable cv_jobs, cv_slots;
  std::deque<Job> jobs;
  bool producer_done = false;
  std::exception_ptr saver_error;

  auto saver = [&] {
    try {
      FR_CUDA_CHECK(cudaSetDevice(d.renderer.get_device()));
      for (;;) {
        Job job;
        {
          std::unique_lock<std::mutex> lock(mu);
          cv_jobs.wait(lock, [&] { return !jobs.empty() || producer_done; });
          if (jobs.empty()) return;
          job = jobs.front();
          jobs.pop_front();
        }
        Slot& s = d.slots[job.slot];
        FrameRecord& rec = *job.rec;
        FR_CUDA_CHECK(cudaEventSynchronize(s.e_copy1));
        FR_CUDA_CHECK(cudaEventElapsedTime(&rec.render_ms, s.e_start, s.e_render));
        FR_CUDA_CHECK(cudaEventElapsedTime(&rec.denoise_ms, s.e_render, s.e_denoise));
        FR_CUDA_CHECK(cudaEventElapsedTime(&rec.post_ms, s.e_denoise, s.e_post));
        FR_CUDA_CHECK(cudaEventElapsedTime(&rec.transfer_ms, s.e_copy0, s.e_copy1));
        const size_t n_bytes = (size_t)d.out_w * d.out_h * 4;
        if (cfg.keep_frames) rec.rgba8.assign(s.h_rgba8, s.h_rgba8 + n_bytes);
        if (!cfg.output_dir.empty()) {
          // "<dir>/<frame>.png" (rtcamp8.cpp:284-288)
          auto t0 = Clock::now();
          const std::vector<uint8_t> png = codec::encode_png(s.h_rgba8, (int)d.out_w, (int)d.out_h, 4);
          rec.encode_ms = ms_since(t0);
          rec.path = (std::filesystem::path(cfg.output_dir) / (std::to_string(rec.frame_idx) + ".png")).string();
          t0 = Clock::now();
          std::ofstream f(rec.path, std::ios::binary);
          f.write(reinterpret_cast<const char*>(png.data()), (std::streamsize)png.size());
          f.close();
          if (!f) throw std::runtime_error("FrameBatch: failed to write " + rec.path);
          rec.save_ms = ms_since(t0);
          rec.png_bytes = png.size();
        }
        {
          std::lock_guard<std::mutex> lock(mu);
          s.busy = false;
        }
        cv_slots.notify_all();
      }
    } catch (...) {
      std::lock_guard<std::mutex> lock(mu);
      if (!saver_error) saver_error = std::current_exception();
      for (Slot& s : d.slots) s.busy = false;  // never leave the producer waiting
      cv_slots.notify_all();
    }
  };
  std::vector<std::thread> savers;
  for (uint32_t i = 0; i < cfg.n_save_threads; ++i) savers.emplace_back(saver);

  const auto t_begin = Clock::now();
  std::exception_ptr producer_error;
  try {
    Camera camera = camera_in;
    const float time_step = 1.0f / cfg.fps;
    // the reference advances time by repeated float addition (rtcamp8.cpp:251); a strided
    // rank reproduces the same values by stepping through the frames it skips
    float time = cfg.start_time;
    uint32_t frame_idx = 0;
    for (; frame_idx < cfg.first_frame; ++frame_idx) time += time_step;
    for (uint32_t k = 0; k < cfg.max_frames; ++k) {
      if (time > cfg.max_time) break;
      if (seconds_since(t_begin) > cfg.kill_time_s) {
        result.killed = true;
        break;
      }
      const size_t si = k % d.slots.size();
      Slot& s = d.slots[si];
      {
        std::unique_lock<std::mutex> lock(mu);
        cv_slots.wait(lock, [&] { return !s.busy || saver_error; });
        if (saver_error) break;
        s.busy = true;
      }
      records.emplace_back();
      FrameRecord& rec = records.back();
      rec.frame_idx = frame_idx;
      rec.time = time;
      if (cfg.animate) {
        const auto t0 = Clock::now();
        d.renderer.set_time(time);
        rec.accel_ms = ms_since(t0);
      }
      if (hook) hook(frame_idx, time, camera);
      d.enqueue_frame(s, camera);
      {
        std::lock_guard<std::mutex> lock(mu);
        jobs.push_back(Job{si, &rec});
      }
      cv_jobs.notify_one();
      for (uint32_t j = 0; j < cfg.frame_stride; ++j) time += time_step;
      frame_idx += cfg.frame_stride;
    }
  } catch (...) {
    producer_error = std::current_exception();
  }
  {
    std::lock_guard<std::mutex> lock(mu);
    producer_done = true;
  }
  cv_jobs.notify_all();
  for (std::thread& t : savers) t.join();
  result.wall_s = seconds_since(t_begin);
  if (producer_error) std::rethrow_exception(producer_error);
  if (saver_error) std::rethrow_exception(saver_error);
  result.frames.assign(std::make_move_iterator(records.begin()), std::make_move_iterator(records.end()));
  return result;
}

}  // namespace fredholm
