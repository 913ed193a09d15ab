// Coherence sort of the ray queues ("sorting by ray direction and origin").
//
// Bounce rays leave the shade stage in the order their paths were shaded, i.e. scattered over
// the whole scene.  Before a traversal stage the queue is re-ordered by a key made of the Morton
// code of the ray-origin cell (scene bounds cut into 2^b cells per axis) and the direction octant,
// so that the 32 rays a warp pulls start close together and walk the same BVH sub-trees in the
// same child order.  The sort is a three-kernel counting sort whose item count lives on the
// device (WaveControl), so the host never learns a queue size:
//   k_sort_count    key per item + histogram   (warp-aggregated: __match_any_sync groups equal keys,
//                                               one atomic per group)
//   k_sort_scan     exclusive prefix sum over the bins (one CTA, decoupled from the item count)
//   k_sort_scatter  item -> its slot in the sorted order (same warp aggregation)
// The order of equal keys is not deterministic, but every path's arithmetic is: a path's rays are
// traced and accumulated by kernel sequence, never by queue position.
//
// No counterpart in the reference (OptiX schedules rays itself); HBM-bound: 4 + 32 B read per ray
// for the key (origin + direction), 4 B key write/read, 4 B index write.
#include "cuda_util.h"
#include "queue.cuh"
#include "wavefront.h"
#include "wavefront_kernels.h"

namespace frd
{
namespace
{

constexpr int kBlock = 256;

FR_D uint32_t spread10(uint32_t x)
{
  x &= 0x3ffu;
  x = (x | (x << 16)) & 0x030000ffu;
  x = (x | (x << 8)) & 0x0300f00fu;
  x = (x | (x << 4)) & 0x030c30c3u;
  x = (x | (x << 2)) & 0x09249249u;
  return x;
}

FR_D uint32_t ray_key(const SortGrid& g, const float3& o, const float3& d)
{
  const float hi = (float)((1u << g.cell_bits) - 1u);
  const uint32_t cx = (uint32_t)fminf(fmaxf((o.x - g.lo.x) * g.inv_cell.x, 0.0f), hi);
  const uint32_t cy = (uint32_t)fminf(fmaxf((o.y - g.lo.y) * g.inv_cell.y, 0.0f), hi);
  const uint32_t cz = (uint32_t)fminf(fmaxf((o.z - g.lo.z) * g.inv_cell.z, 0.0f), hi);
  const uint32_t cell = spread10(cx) | (spread10(cy) << 1) | (spread10(cz) << 2);
  if (!g.use_octant) return cell;
  const uint32_t oct = (d.x >= 0.0f ? 1u : 0u) | (d.y >= 0.0f ? 2u : 0u) | (d.z >= 0.0f ? 4u : 0u);
  return (cell << 3) | oct;
}

// origin / direction of item i of queue `which`
FR_D void load_ray(const WaveBuffers& wb, int which, uint32_t i, float3& o, float3& d)
{
  if (which == SORT_RADIANCE0 || which == SORT_RADIANCE1) {
    const uint32_t slot = wb.queue[which][i];
    o = f3(wb.ray_o[slot]);
    d = f3(wb.ray_d[slot]);
  } else {
    const float4* q = reinterpret_cast<const float4*>(which == SORT_LIGHT ? (const void*)wb.light
                                                                         : (const void*)wb.shadow[which - SORT_SHADOW0]);
    o = f3(q[3ull * i]);
    d = f3(q[3ull * i + 1]);
  }
}

FR_D uint32_t queue_size(const WaveControl* ctl, int which)
{
  switch (which) {
    case SORT_RADIANCE0:
    case SORT_RADIANCE1: return ctl->n[Q_CUR];
    case SORT_LIGHT: return ctl->n[Q_LIGHT];
    default: return ctl->n[Q_SHADOW0 + (which - SORT_SHADOW0)];
  }
}

__global__ void __launch_bounds__(kBlock) k_sort_count(WaveBuffers wb, SortGrid g, int which, uint32_t* __restrict__ keys,
                                                       uint32_t* __restrict__ bins)
{
  const uint32_t n = queue_size(wb.ctl, which);
  const uint32_t n_round = (n + 31u) & ~31u;  // whole warps take part in the vote
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
    uint32_t key = 0xffffffffu;
    if (i < n) {
      float3 o, d;
      load_ray(wb, which, i, o, d);
      key = ray_key(g, o, d);
      keys[i] = key;
    }
    const uint32_t same = __match_any_sync(0xffffffffu, key);
    if (i < n && (int)lane_id() == __ffs(same) - 1) atomicAdd(&bins[key], (uint32_t)__popc(same));
  }
}

// exclusive scan of n_bins counters in place, one CTA of 1024 threads
__global__ void __launch_bounds__(1024) k_sort_scan(uint32_t* __restrict__ bins, uint32_t n_bins)
{
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  const uint32_t per_thread = (n_bins + 1023u) / 1024u;
  const uint32_t begin = threadIdx.x * per_thread;
  const uint32_t end = min(begin + per_thread, n_bins);
  uint32_t sum = 0;
  for (uint32_t i = begin; i < end; ++i) sum += bins[i];
  // block-wide exclusive scan of the per-thread sums
  uint32_t incl = sum;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
    if ((int)lane_id() >= off) incl += v;
  }
  if (lane_id() == 31u) s_warp[threadIdx.x >> 5] = incl;
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t w = s_warp[threadIdx.x];
    uint32_t wi = w;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, wi, off);
      if ((int)threadIdx.x >= off) wi += v;
    }
    s_warp[threadIdx.x] = wi - w;
    if (threadIdx.x == 31) s_carry = wi;
  }
  __syncthreads();
  uint32_t run = s_warp[threadIdx.x >> 5] + incl - sum;
  for (uint32_t i = begin; i < end; ++i) {
    const uint32_t c = bins[i];
    bins[i] = run;
    run += c;
  }
  (void)s_carry;
}

__global__ void __launch_bounds__(kBlock) k_sort_scatter(WaveBuffers wb, int which, const uint32_t* __restrict__ keys,
                                                         uint32_t* __restrict__ bins, uint32_t* __restrict__ out)
{
  const uint32_t n = queue_size(wb.ctl, which);
  const uint32_t n_round = (n + 31u) & ~31u;
  const bool radiance = which == SORT_RADIANCE0 || which == SORT_RADIANCE1;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
    const uint32_t key = i < n ? keys[i] : 0xffffffffu;
    const uint32_t same = __match_any_sync(0xffffffffu, key);
    const int leader = __ffs(same) - 1;
    uint32_t base = 0;
    if (i < n && (int)lane_id() == leader) base = atomicAdd(&bins[key], (uint32_t)__popc(same));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (i < n) out[base + __popc(same & ((1u << lane_id()) - 1u))] = radiance ? wb.queue[which][i] : i;
  }
}

GridCache g_sort_grid;

}  // namespace

void launch_coherence_sort(cudaStream_t s, const WaveBuffers& wb, const SortGrid& g, int which, uint32_t* keys,
                           uint32_t* bins, uint32_t* out)
{
  const int grid = g_sort_grid.get(reinterpret_cast<const void*>(k_sort_count), kBlock);
  const uint32_t n_bins = sort_bins(g);
  FR_CUDA_CHECK(cudaMemsetAsync(bins, 0, sizeof(uint32_t) * n_bins, s));
  k_sort_count<<<grid, kBlock, 0, s>>>(wb, g, which, keys, bins);
  FR_CUDA_LAUNCH_CHECK();
  k_sort_scan<<<1, 1024, 0, s>>>(bins, n_bins);
  FR_CUDA_LAUNCH_CHECK();
  k_sort_scatter<<<grid, kBlock, 0, s>>>(wb, which, keys, bins, out);
  FR_CUDA_LAUNCH_CHECK();
}

}  // namespace frd
