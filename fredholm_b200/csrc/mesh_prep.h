// GPU side of scene ingestion: the per-corner expansion, face-normal generation and
// vertex de-duplication of the .obj loader (reference fredholm/src/scene.cpp:317-393, a serial
// unordered_map<Vertex, uint32_t> loop on the host) as CUDA kernels.  The result is the SAME arrays, bit for bit
// and in the same order -- unique vertices in order of first occurrence -- so the host loop (csrc/scene.cpp)
// and this path are interchangeable; Scene::load_obj takes this one when a CUDA device is present and the mesh is
// large enough to pay for the transfers.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

#include <cuda_runtime.h>

namespace frd
{

struct ObjCorner {
  int v, vt, vn;  // 0-based indices into the attribute pools, -1 = absent
};

struct PreparedMesh {
  std::vector<float3> vertices, normals;
  std::vector<float2> texcoords;
  std::vector<uint3> indices;
  float gpu_ms = 0.0f;
};

// true if a CUDA device can run the kernels (never throws)
bool mesh_prep_available();

// pos / nrm / tex: the file's attribute pools (3, 3, 2 floats per entry); corners: 3 per triangle.
// Throws std::runtime_error("vertex index out of range") like the host loop.
void prepare_mesh_gpu(const std::vector<float>& pos, const std::vector<float>& nrm, const std::vector<float>& tex,
                      const std::vector<ObjCorner>& corners, PreparedMesh& out);

}  // namespace frd
