"""Structure statistics + validity of the GPU-built CWBVH on the benchmark scene."""
import sys
import numpy as np
sys.path.insert(0, ".")
from fredholm_b200 import Renderer, scenes

small = "--small" in sys.argv
s = scenes.standard_surface_scene(128, 64) if small else scenes.standard_surface_scene()
r = Renderer(0); r.set_scene(s); r.build_accel()
print(r.accel_info())
nodes, tris = r.accel_data()
meta = nodes["meta"]
inner = ((meta & 0x20) != 0) & ((meta >> 5) == 1) & ((meta & 0x1f) >= 24)
empty = meta == 0
leaf = ~inner & ~empty
cnt = np.where(leaf, np.select([(meta >> 5) == 1, (meta >> 5) == 3, (meta >> 5) == 7], [1, 2, 3], 0), 0)
nch = (~empty).sum(1)
print("nodes %d  children/node mean %.2f  hist %s" % (len(nodes), nch.mean(), np.bincount(nch, minlength=9)))
print("inner children/node %.2f  leaf children/node %.2f  tris/leaf %.2f  tris/node %.2f" % (
    inner.sum(1).mean(), leaf.sum(1).mean(), cnt.sum() / max(leaf.sum(), 1), cnt.sum(1).mean()))
bottom = inner.sum(1) == 0
print("bottom nodes (no inner child): %d (%.1f%%), their children mean %.2f, tris mean %.2f" % (
    bottom.sum(), 100 * bottom.mean(), nch[bottom].mean(), cnt.sum(1)[bottom].mean()))
# validity: every leaf triangle inside its dequantised child box
step = np.ldexp(1.0, nodes["e"].astype(np.int32) - 127)              # (N,3)
lo = nodes["p"][:, :, None] + nodes["qlo"].astype(np.float64) * step[:, :, None]   # (N,3,8)
hi = nodes["p"][:, :, None] + nodes["qhi"].astype(np.float64) * step[:, :, None]
bad = 0
ni, si = np.nonzero(leaf)
off = (meta[ni, si] & 0x1f).astype(np.int64)
base = nodes["tri_base"][ni].astype(np.int64) + off
for k in range(3):
    sel = cnt[ni, si] > k
    t = tris[base[sel] + k][:, :, :3]                   # (M,3 verts,3)
    l = lo[ni[sel], :, si[sel]][:, None, :]
    h = hi[ni[sel], :, si[sel]][:, None, :]
    bad += int(((t < l - 1e-9) | (t > h + 1e-9)).any(axis=(1, 2)).sum())
print("triangles outside their leaf box:", bad)
faces = tris[:, 0, 3].view(np.uint32)
print("every face exactly once:", np.array_equal(np.sort(faces), np.arange(len(faces), dtype=np.uint32)))
# depth of every node / leaf (children of node i start at child_base, inner children in slot order)
depth = np.zeros(len(nodes), dtype=np.int32)
order = np.argsort(nodes["child_base"], kind="stable")  # parents precede children in the array (level order)
for i in range(len(nodes)):
    k = int(inner[i].sum())
    if k:
        b = int(nodes["child_base"][i])
        depth[b:b + k] = depth[i] + 1
leaf_depth = np.repeat(depth, leaf.sum(1))
w = np.repeat(cnt.sum(1), 1)
print("node depth hist", np.bincount(depth))
print("mean leaf depth (per triangle) %.2f  max %d" % ((depth * cnt.sum(1)).sum() / cnt.sum(), depth.max()))
