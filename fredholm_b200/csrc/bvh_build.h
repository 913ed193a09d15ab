// Host interface of the GPU BVH builder (bvh_build.cu).
#pragma once
#include <cstdint>
#include <vector>

#include "bvh.cuh"
#include "cuda_util.h"
#include "fredholm/shared.h"

namespace frd
{

struct DeviceBvh {
  DeviceBvh() = default;
  DeviceBvh(DeviceBvh&&) = default;
  DeviceBvh& operator=(DeviceBvh&&) = default;
  DevBuf<Node8> nodes;
  DevBuf<float4> tris;  // LeafTri as float4[3], leaf order
  uint32_t n_nodes = 0;
  uint32_t n_faces = 0;
  uint32_t depth = 0;  // levels of the 8-wide tree
  uint32_t ploc_rounds = 0;  // merge rounds of the PLOC builder (0: LBVH)
  // small trees (<= 16 Ki primitives, built in one launch): first node of every level, level_begin[depth] = n_nodes
  // (nodes are numbered level by level) -- what a bottom-up refit walks
  std::vector<uint32_t> level_begin;
  float bounds_lo[3] = {0, 0, 0}, bounds_hi[3] = {0, 0, 0};

  BvhView view() const
  {
    return BvhView{reinterpret_cast<const float4*>(nodes.get()), tris.get()};
  }
};

// Builds the world-space CWBVH for `n_faces` triangles.  face_submesh[f] selects
// the object-to-world transform of face f; face_flags[f] bit 0 marks alpha-tested
// faces (may be null).  Synchronises `stream` before returning.
// builder: -1 = default (PLOC, or FRD_BVH_BUILDER), 0 = Karras radix tree, 1 = PLOC
void build_bvh(cudaStream_t stream, const float3* d_vertices, const uint3* d_indices,
               const uint32_t* d_face_submesh, const uint32_t* d_face_flags,
               const fredholm::Matrix3x4* d_o2w, uint32_t n_faces, DeviceBvh& out, int builder = -1);

// ---- two-level acceleration structure (accel.cu) -------------------------------------------------------------
// One object-space tree (BLAS) per DISTINCT mesh + an instance tree (TLAS) over the instances' world boxes, in one
// pair of node / triangle arrays (bvh.cuh BvhView): [TLAS, fixed capacity][BLAS 0][BLAS 1]...  A transform change
// only rebuilds the TLAS (update_tlas): a few thousand boxes, no triangle is touched.
// Replaces optixAccelBuild of the IAS (renderer.h:498-552) and its rebuild in set_time (renderer.h:614-640).
struct TwoLevelBvh {
  DevBuf<Node8> nodes;
  DevBuf<float4> tris;
  DevBuf<InstanceRecord> instances;
  DevBuf<float> mesh_bounds;        // 6 floats per instance: object-space bounds of its mesh
  DevBuf<float3> placeholder_vertices;
  DevBuf<uint3> placeholder_indices;
  DevBuf<uint32_t> zeros;           // "sub-mesh 0" for every face of a build that uses the identity transform
  DevBuf<fredholm::Matrix3x4> identity;
  DeviceBvh tlas;                   // scratch of the last TLAS build
  uint32_t n_instances = 0, n_meshes = 0;
  uint32_t tlas_node_capacity = 0;
  uint32_t n_nodes = 0;             // TLAS capacity + all BLAS nodes
  uint32_t n_blas_faces = 0;        // triangles stored (distinct meshes only)
  uint32_t blas_depth = 0, depth = 0;
  float tlas_ms = 0.0f;             // device time of the last update_tlas / refit_tlas
  bool last_update_was_refit = false;
  float rebuilt_root_area = 0.0f;   // half surface area of the TLAS root box at its last rebuild
  DevBuf<float> refit_boxes;        // [6 per TLAS node] + [6 per instance] scratch of refit_tlas
  DevBuf<float> refit_root;         // [6] root box of the last refit
  float bounds_lo[3] = {0, 0, 0}, bounds_hi[3] = {0, 0, 0};
  size_t bytes() const { return nodes.bytes() + tris.bytes() + instances.bytes(); }
  BvhView view(const fredholm::Matrix3x4* d_w2o) const
  {
    return BvhView{reinterpret_cast<const float4*>(nodes.get()), tris.get(), instances.get(),
                   reinterpret_cast<const float4*>(d_w2o)};
  }
};

// mesh_of_submesh[i] = index of the distinct mesh sub-mesh i is a copy of; representative[m] = a sub-mesh holding
// mesh m's geometry.  face_flags_of_mesh (optional, host): per distinct mesh, per local face, the OR of the
// alpha-test flags of its instances.  Synchronises the stream.
void build_two_level(cudaStream_t stream, const float3* d_vertices, const uint3* d_indices,
                     const std::vector<uint32_t>& submesh_offsets, const std::vector<uint32_t>& submesh_n_faces,
                     const std::vector<uint32_t>& mesh_of_submesh, const std::vector<uint32_t>& representative,
                     const std::vector<std::vector<uint32_t>>* face_flags_of_mesh,
                     const fredholm::Matrix3x4* d_o2w, TwoLevelBvh& out);
// instance boxes from the current transforms + TLAS rebuild (radix tree).  Synchronises the stream.
// builder: 0 = radix tree (a handful of launches: the choice for a rebuild between frames), 1 / -1 = PLOC
// (better tree, ~1 ms of merge rounds: the choice for the first build)
void update_tlas(cudaStream_t stream, const fredholm::Matrix3x4* d_o2w, TwoLevelBvh& out, int builder = 0);
// instance boxes from the current transforms + REFIT of the instance tree: the topology of the last rebuild is
// kept, every node's child boxes are recomputed bottom-up and re-quantised in ONE launch (reference: the IAS
// rebuild of Renderer::set_time, renderer.h:614-640).  Returns false (nothing done) when the tree has no level
// table or has grown to more than `max_growth` times the surface area it had at its last rebuild -- the caller
// rebuilds then.  Synchronises the stream.
bool refit_tlas(cudaStream_t stream, const fredholm::Matrix3x4* d_o2w, TwoLevelBvh& out, float max_growth = 2.0f);

}  // namespace frd
