// Lobe indices and compile-time lobe sets of the layered BSDF (host + device).
#pragma once
#include <cstdint>

namespace frd
{

enum Lobe : int {
  LOBE_COAT = 0,
  LOBE_METAL,
  LOBE_SPECULAR,
  LOBE_TRANSMISSION,
  LOBE_SHEEN,
  LOBE_DIFFUSE_T,
  LOBE_DIFFUSE_R,
  LOBE_COUNT
};

// Compile-time lobe sets.  The shade stage is instantiated once per material class
// (wavefront_kernels.h: ShadeClass) with the set of lobes that class can have, so a
// kernel only carries the code of the lobes it needs and a warp never diverges over
// lobes its material does not have.  A lobe absent from MASK must be one whose
// run-time gate is false for every material of the class (renderer.cpp:
// classify_material), which makes the specialised code bit-identical to the generic one.
constexpr uint32_t M_COAT = 1u << LOBE_COAT, M_METAL = 1u << LOBE_METAL, M_SPECULAR = 1u << LOBE_SPECULAR,
                   M_TRANSMISSION = 1u << LOBE_TRANSMISSION, M_SHEEN = 1u << LOBE_SHEEN,
                   M_DIFFUSE_T = 1u << LOBE_DIFFUSE_T, M_DIFFUSE_R = 1u << LOBE_DIFFUSE_R;
constexpr uint32_t M_ALL = 0x7fu;

}  // namespace frd
