// Host interface of the GPU BVH builder (bvh_build.cu).
#pragma once
#include <cstdint>

#include "bvh.cuh"
#include "cuda_util.h"
#include "fredholm/shared.h"

namespace frd
{

struct DeviceBvh {
  DevBuf<Node8> nodes;
  DevBuf<float4> tris;  // LeafTri as float4[3], leaf order
  uint32_t n_nodes = 0;
  uint32_t n_faces = 0;
  uint32_t depth = 0;  // levels of the 8-wide tree
  uint32_t ploc_rounds = 0;  // merge rounds of the PLOC builder (0: LBVH)
  float bounds_lo[3] = {0, 0, 0}, bounds_hi[3] = {0, 0, 0};

  BvhView view() const
  {
    return BvhView{reinterpret_cast<const float4*>(nodes.get()), tris.get()};
  }
};

// Builds the world-space CWBVH for `n_faces` triangles.  face_submesh[f] selects
// the object-to-world transform of face f; face_flags[f] bit 0 marks alpha-tested
// faces (may be null).  Synchronises `stream` before returning.
void build_bvh(cudaStream_t stream, const float3* d_vertices, const uint3* d_indices,
               const uint32_t* d_face_submesh, const uint32_t* d_face_flags,
               const fredholm::Matrix3x4* d_o2w, uint32_t n_faces, DeviceBvh& out);

}  // namespace frd
