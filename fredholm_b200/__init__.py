"""fredholm_b200 -- Python binding of the B200-native fredholm rendering core.

The product is libfredholm_b200.so (hand-written sm_100a CUDA kernels, C++ host,
C ABI in include/fredholm_b200.h).  This package is a thin ctypes mirror of the
reference's Renderer / Camera call surface plus the procedural scenes used by the
tests and the benchmark.  There is no CPU fallback: importing works anywhere, but
every compute call needs the shared library and a GPU and fails loudly otherwise.
"""
from .types import MATERIAL_DTYPE, SceneArrays, make_material  # noqa: F401
from .api import Renderer, Camera, Scene, DeviceLayers, lib, LibraryNotBuilt  # noqa: F401
