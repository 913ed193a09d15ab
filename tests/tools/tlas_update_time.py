"""Instance-tree update of the two-level structure: device time and (FRD_BVH_VERBOSE=1) its phases."""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np
from fredholm_b200 import Renderer, scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3072
s = scenes.instanced_scene(n_instances=n, mesh_res=(16, 8), terrain_res=64)
r = Renderer(0); r.set_accel_mode("two_level"); r.set_scene(s); r.build_accel()
tr = s.transforms.copy().reshape(-1, 4, 4)
rng = np.random.default_rng(1)
rays = np.concatenate([rng.uniform(-100, 100, (400000, 3)), rng.normal(size=(400000, 3))], 1).astype(np.float32)
for k in range(5):
    tr[7, 3, 1] += 1.0
    r.trace_closest(rays)          # keeps the SM clocks up: an idle GPU runs the update's small kernels 3-6x slower
    t0 = time.perf_counter(); r.set_transforms(tr.reshape(-1, 16)); t1 = time.perf_counter()
    print("update %d: device %.3f ms, host wall %.3f ms" % (k, r.accel_info()["tlas_update_ms"], 1e3 * (t1 - t0)), flush=True)
