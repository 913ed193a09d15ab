// Two-level acceleration structure: per-mesh object-space trees + an instance tree (bvh_build.h TwoLevelBvh).
// Built from the same GPU builder as the flat tree: a BLAS is build_bvh() over one mesh's faces with the identity
// transform; the TLAS is build_bvh() over one PLACEHOLDER triangle per instance whose vertices span the
// instance's world box (so the builder's leaf boxes are exactly the instance boxes and its leaf "triangles" carry
// the instance index).  The traversal kernel (bvh.cuh, TWO = true) enters an instance where the flat kernel would
// test a triangle.
#include <algorithm>
#include <stdexcept>

#include "bvh_build.h"
#include "bvh_quant.cuh"

namespace frd
{
namespace
{

// copies a BLAS into the combined array, making its child / triangle indices absolute
__global__ void k_relocate_nodes(const Node8* __restrict__ src, uint32_t n, uint32_t node_offset, uint32_t tri_offset,
                                 Node8* __restrict__ dst)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Node8 nd = src[i];
  nd.child_base += node_offset;
  nd.tri_base += tri_offset;
  dst[node_offset + i] = nd;
}

// world box of every instance (the 8 corners of its mesh's object-space box through the instance transform)
// as a placeholder triangle: v0 = lo, v1 = hi, v2 = a third corner -- its bounding box is the instance box
__global__ void k_instance_placeholders(const float* __restrict__ mesh_bounds, const fredholm::Matrix3x4* __restrict__ o2w,
                                        uint32_t n, float3* __restrict__ vertices, uint3* __restrict__ indices)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* b = mesh_bounds + 6ull * i;
  const fredholm::Matrix3x4 m = o2w[i];
  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float x = (c & 1) ? b[3] : b[0], y = (c & 2) ? b[4] : b[1], z = (c & 4) ? b[5] : b[2];
    const float4 r[3] = {m.m[0], m.m[1], m.m[2]};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float v = r[a].x * x + r[a].y * y + r[a].z * z + r[a].w;
      lo[a] = fminf(lo[a], v);
      hi[a] = fmaxf(hi[a], v);
    }
  }
  // a few ulps of slack for the rounding of the corner transform and of the ray transform at traversal time
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float pad = 4e-7f * fmaxf(fabsf(lo[a]), fabsf(hi[a])) + 1e-30f;
    lo[a] -= pad;
    hi[a] += pad;
  }
  vertices[3ull * i] = make_float3(lo[0], lo[1], lo[2]);
  vertices[3ull * i + 1] = make_float3(hi[0], hi[1], hi[2]);
  vertices[3ull * i + 2] = make_float3(lo[0], hi[1], lo[2]);
  indices[i] = make_uint3(3u * i, 3u * i + 1u, 3u * i + 2u);
}

// ---- instance-tree refit ----------------------------------------------------------------------------------
constexpr int kRefitThreads = 1024;
constexpr int kMaxRefitLevels = kSmemStack + kLocalStack;
struct LevelTable {
  uint32_t depth;
  uint32_t begin[kMaxRefitLevels + 1];
};

// One block walks the instance tree bottom-up, level by level (nodes are numbered level by level, so a node's
// inner children have already been fitted): child boxes from the refreshed placeholders or from the children's
// node boxes, node box = their union, node frame and quantised child boxes rewritten in place.  Slot assignment
// (octant order) and topology are those of the last rebuild.
__global__ void __launch_bounds__(kRefitThreads) k_refit_tlas(Node8* __restrict__ nodes, float4* __restrict__ tris,
                                                             const float3* __restrict__ placeholder_vertices, uint32_t n_inst,
                                                             LevelTable lv, float* __restrict__ node_box, float* __restrict__ root_out)
{
  // the placeholders keep their place in the leaf order; only their boxes move
  for (uint32_t j = threadIdx.x; j < n_inst; j += blockDim.x) {
    const float4 v0 = tris[3ull * j], v1 = tris[3ull * j + 1];
    const uint32_t id = __float_as_uint(v0.w);
    const float3 lo = placeholder_vertices[3ull * id], hi = placeholder_vertices[3ull * id + 1];
    tris[3ull * j] = make_float4(lo.x, lo.y, lo.z, v0.w);
    tris[3ull * j + 1] = make_float4(hi.x, hi.y, hi.z, v1.w);
    tris[3ull * j + 2] = make_float4(lo.x, hi.y, lo.z, 0.0f);
  }
  __threadfence_block();
  __syncthreads();
  for (int level = (int)lv.depth - 1; level >= 0; --level) {
    for (uint32_t node = lv.begin[level] + threadIdx.x; node < lv.begin[level + 1]; node += blockDim.x) {
      Node8 nd = nodes[node];
      float clo[8][3], chi[8][3];
      float nlo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, nhi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
      for (int s = 0; s < 8; ++s) {
        const uint32_t m = nd.meta[s];
        if (m == 0u) continue;
        float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
        if ((m & 0x1fu) >= 24u) {
          const uint32_t c = nd.child_base + __popc((uint32_t)nd.imask & ((1u << s) - 1u));
          for (int a = 0; a < 3; ++a) {
            lo[a] = node_box[6ull * c + a];
            hi[a] = node_box[6ull * c + 3 + a];
          }
        } else {
          const uint32_t cnt = __popc(m >> 5), off = m & 0x1fu;
          for (uint32_t q = 0; q < cnt; ++q) {
            const float4 a0 = tris[3ull * (nd.tri_base + off + q)], a1 = tris[3ull * (nd.tri_base + off + q) + 1];
            lo[0] = fminf(lo[0], a0.x), lo[1] = fminf(lo[1], a0.y), lo[2] = fminf(lo[2], a0.z);
            hi[0] = fmaxf(hi[0], a1.x), hi[1] = fmaxf(hi[1], a1.y), hi[2] = fmaxf(hi[2], a1.z);
          }
        }
        for (int a = 0; a < 3; ++a) {
          clo[s][a] = lo[a];
          chi[s][a] = hi[a];
          nlo[a] = fminf(nlo[a], lo[a]);
          nhi[a] = fmaxf(nhi[a], hi[a]);
        }
      }
      nd.px = nlo[0];
      nd.py = nlo[1];
      nd.pz = nlo[2];
      const uint32_t ex = grid_exponent(__fsub_ru(nhi[0], nlo[0])), ey = grid_exponent(__fsub_ru(nhi[1], nlo[1])),
                     ez = grid_exponent(__fsub_ru(nhi[2], nlo[2]));
      nd.ex = (uint8_t)ex;
      nd.ey = (uint8_t)ey;
      nd.ez = (uint8_t)ez;
      const float isx = grid_inverse_step(ex), isy = grid_inverse_step(ey), isz = grid_inverse_step(ez);
      for (int s = 0; s < 8; ++s) {
        if (nd.meta[s] == 0u) continue;
        nd.qlox[s] = quantize_lo(clo[s][0], nlo[0], isx);
        nd.qloy[s] = quantize_lo(clo[s][1], nlo[1], isy);
        nd.qloz[s] = quantize_lo(clo[s][2], nlo[2], isz);
        nd.qhix[s] = quantize_hi(chi[s][0], nlo[0], isx);
        nd.qhiy[s] = quantize_hi(chi[s][1], nlo[1], isy);
        nd.qhiz[s] = quantize_hi(chi[s][2], nlo[2], isz);
      }
      nodes[node] = nd;
      for (int a = 0; a < 3; ++a) {
        node_box[6ull * node + a] = nlo[a];
        node_box[6ull * node + 3 + a] = nhi[a];
      }
    }
    __threadfence_block();
    __syncthreads();
  }
  if (threadIdx.x < 6) root_out[threadIdx.x] = node_box[threadIdx.x];
}

float half_area6(const float* b)
{
  const float ex = b[3] - b[0], ey = b[4] - b[1], ez = b[5] - b[2];
  return ex * ey + ey * ez + ez * ex;
}

}  // namespace

bool refit_tlas(cudaStream_t stream, const fredholm::Matrix3x4* d_o2w, TwoLevelBvh& out, float max_growth)
{
  const uint32_t n = out.n_instances;
  const std::vector<uint32_t>& lb = out.tlas.level_begin;
  if (n == 0 || lb.size() < 2 || lb.size() > (size_t)kMaxRefitLevels + 1 || out.rebuilt_root_area <= 0.0f) return false;
  LevelTable lv;
  lv.depth = (uint32_t)lb.size() - 1;
  for (size_t k = 0; k < lb.size(); ++k) lv.begin[k] = lb[k];
  out.refit_boxes.reserve(6ull * out.tlas.n_nodes);
  out.refit_root.reserve(6);
  cudaEvent_t e0, e1;
  FR_CUDA_CHECK(cudaEventCreate(&e0));
  FR_CUDA_CHECK(cudaEventCreate(&e1));
  FR_CUDA_CHECK(cudaEventRecord(e0, stream));
  k_instance_placeholders<<<(n + 127) / 128, 128, 0, stream>>>(out.mesh_bounds.get(), d_o2w, n, out.placeholder_vertices.get(),
                                                               out.placeholder_indices.get());
  FR_CUDA_LAUNCH_CHECK();
  k_refit_tlas<<<1, kRefitThreads, 0, stream>>>(out.nodes.get(), out.tris.get(), out.placeholder_vertices.get(), n, lv,
                                               out.refit_boxes.get(), out.refit_root.get());
  FR_CUDA_LAUNCH_CHECK();
  FR_CUDA_CHECK(cudaEventRecord(e1, stream));
  float root[6];
  FR_CUDA_CHECK(cudaMemcpyAsync(root, out.refit_root.get(), sizeof(root), cudaMemcpyDeviceToHost, stream));
  FR_CUDA_CHECK(cudaStreamSynchronize(stream));
  FR_CUDA_CHECK(cudaEventElapsedTime(&out.tlas_ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  for (int a = 0; a < 3; ++a) {
    out.bounds_lo[a] = root[a];
    out.bounds_hi[a] = root[3 + a];
  }
  out.last_update_was_refit = true;
  // instances that have moved far apart leave a stale topology (siblings that no longer are neighbours): the tree
  // is still correct, but the caller should rebuild it
  return half_area6(root) <= max_growth * out.rebuilt_root_area;
}

void update_tlas(cudaStream_t stream, const fredholm::Matrix3x4* d_o2w, TwoLevelBvh& out, int builder)
{
  const uint32_t n = out.n_instances;
  if (n == 0) throw std::runtime_error("two-level bvh: no instances");
  cudaEvent_t e0, e1;
  FR_CUDA_CHECK(cudaEventCreate(&e0));
  FR_CUDA_CHECK(cudaEventCreate(&e1));
  FR_CUDA_CHECK(cudaEventRecord(e0, stream));
  k_instance_placeholders<<<(n + 127) / 128, 128, 0, stream>>>(out.mesh_bounds.get(), d_o2w, n, out.placeholder_vertices.get(),
                                                               out.placeholder_indices.get());
  FR_CUDA_LAUNCH_CHECK();
  build_bvh(stream, out.placeholder_vertices.get(), out.placeholder_indices.get(), out.zeros.get(), nullptr, out.identity.get(), n,
            out.tlas, builder);
  if (out.tlas.n_nodes > out.tlas_node_capacity) throw std::runtime_error("two-level bvh: TLAS exceeds its reserved node range");
  FR_CUDA_CHECK(cudaMemcpyAsync(out.nodes.get(), out.tlas.nodes.get(), sizeof(Node8) * out.tlas.n_nodes, cudaMemcpyDeviceToDevice, stream));
  FR_CUDA_CHECK(cudaMemcpyAsync(out.tris.get(), out.tlas.tris.get(), sizeof(float4) * 3ull * n, cudaMemcpyDeviceToDevice, stream));
  FR_CUDA_CHECK(cudaEventRecord(e1, stream));
  FR_CUDA_CHECK(cudaEventSynchronize(e1));
  FR_CUDA_CHECK(cudaEventElapsedTime(&out.tlas_ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  for (int a = 0; a < 3; ++a) {
    out.bounds_lo[a] = out.tlas.bounds_lo[a];
    out.bounds_hi[a] = out.tlas.bounds_hi[a];
  }
  const float root6[6] = {out.bounds_lo[0], out.bounds_lo[1], out.bounds_lo[2], out.bounds_hi[0], out.bounds_hi[1], out.bounds_hi[2]};
  out.rebuilt_root_area = half_area6(root6);
  out.last_update_was_refit = false;
  out.depth = out.tlas.depth + out.blas_depth;
  // TLAS levels + what enter_instance parks (3 entries) + BLAS levels must fit the traversal stack
  if (out.depth + 3 + 2 > (uint32_t)(kSmemStack + kLocalStack)) throw std::runtime_error("two-level bvh: tree too deep for the traversal stack");
}

void build_two_level(cudaStream_t stream, const float3* d_vertices, const uint3* d_indices,
                     const std::vector<uint32_t>& submesh_offsets, const std::vector<uint32_t>& submesh_n_faces,
                     const std::vector<uint32_t>& mesh_of_submesh, const std::vector<uint32_t>& representative,
                     const std::vector<std::vector<uint32_t>>* face_flags_of_mesh,
                     const fredholm::Matrix3x4* d_o2w, TwoLevelBvh& out)
{
  const uint32_t n_inst = (uint32_t)submesh_offsets.size(), n_mesh = (uint32_t)representative.size();
  if (n_inst == 0 || n_mesh == 0) throw std::runtime_error("two-level bvh: empty scene");
  out.n_instances = n_inst;
  out.n_meshes = n_mesh;
  uint32_t max_faces = n_inst;
  for (uint32_t m = 0; m < n_mesh; ++m) max_faces = std::max(max_faces, submesh_n_faces[representative[m]]);
  out.zeros.reserve(max_faces);
  FR_CUDA_CHECK(cudaMemsetAsync(out.zeros.get(), 0, sizeof(uint32_t) * max_faces, stream));
  const fredholm::Matrix3x4 ident = fredholm::make_mat3x4(make_float4(1, 0, 0, 0), make_float4(0, 1, 0, 0), make_float4(0, 0, 1, 0));
  out.identity.upload(&ident, 1, stream);

  // ---- one object-space tree per distinct mesh ----
  std::vector<DeviceBvh> blas(n_mesh);
  DevBuf<uint32_t> d_flags;
  for (uint32_t m = 0; m < n_mesh; ++m) {
    const uint32_t s = representative[m], nf = submesh_n_faces[s];
    if (nf == 0) throw std::runtime_error("two-level bvh: sub-mesh without faces");
    const uint32_t* flags = nullptr;
    if (face_flags_of_mesh && !(*face_flags_of_mesh)[m].empty()) {
      d_flags.upload((*face_flags_of_mesh)[m], stream);
      flags = d_flags.get();
    }
    build_bvh(stream, d_vertices, d_indices + submesh_offsets[s], out.zeros.get(), flags, out.identity.get(), nf, blas[m]);
  }

  // ---- layout: [TLAS capacity][BLAS 0][BLAS 1] ... ----
  out.tlas_node_capacity = n_inst / 2 + 2;
  std::vector<uint32_t> node_off(n_mesh), tri_off(n_mesh);
  uint32_t nodes_total = out.tlas_node_capacity, tris_total = n_inst;
  out.blas_depth = 0;
  for (uint32_t m = 0; m < n_mesh; ++m) {
    node_off[m] = nodes_total;
    tri_off[m] = tris_total;
    nodes_total += blas[m].n_nodes;
    tris_total += blas[m].n_faces;
    out.blas_depth = std::max(out.blas_depth, blas[m].depth);
  }
  out.n_nodes = nodes_total;
  out.n_blas_faces = tris_total - n_inst;
  out.nodes.alloc(nodes_total);
  out.tris.alloc(3ull * tris_total);
  FR_CUDA_CHECK(cudaMemsetAsync(out.nodes.get(), 0, sizeof(Node8) * out.tlas_node_capacity, stream));
  for (uint32_t m = 0; m < n_mesh; ++m) {
    const uint32_t n = blas[m].n_nodes;
    k_relocate_nodes<<<(n + 255) / 256, 256, 0, stream>>>(blas[m].nodes.get(), n, node_off[m], tri_off[m], out.nodes.get());
    FR_CUDA_LAUNCH_CHECK();
    FR_CUDA_CHECK(cudaMemcpyAsync(out.tris.get() + 3ull * tri_off[m], blas[m].tris.get(), sizeof(float4) * 3ull * blas[m].n_faces,
                                  cudaMemcpyDeviceToDevice, stream));
  }

  // ---- instances ----
  std::vector<InstanceRecord> rec(n_inst);
  std::vector<float> bounds(6ull * n_inst);
  for (uint32_t i = 0; i < n_inst; ++i) {
    const uint32_t m = mesh_of_submesh[i];
    if (submesh_n_faces[i] != submesh_n_faces[representative[m]]) throw std::runtime_error("two-level bvh: instance / mesh face count mismatch");
    rec[i].blas_root = node_off[m];
    rec[i].face_offset = submesh_offsets[i];
    for (int a = 0; a < 3; ++a) {
      bounds[6ull * i + a] = blas[m].bounds_lo[a];
      bounds[6ull * i + 3 + a] = blas[m].bounds_hi[a];
    }
  }
  out.instances.upload(rec, stream);
  out.mesh_bounds.upload(bounds, stream);
  out.placeholder_vertices.reserve(3ull * n_inst);
  out.placeholder_indices.reserve(n_inst);
  FR_CUDA_CHECK(cudaStreamSynchronize(stream));  // blas[] and the host vectors go out of scope
  update_tlas(stream, d_o2w, out, /*builder=*/-1);  // first build: the default (PLOC) builder
}

}  // namespace frd
