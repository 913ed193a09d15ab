"""Debug helper: image differences against the oracle on small open-mesh scenes (back-face hits, alpha cut-outs)."""
import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from fredholm_b200 import Camera, DeviceLayers, Renderer, api, scenes
from fredholm_b200.scenes import _assemble, _quad
from fredholm_b200.types import make_material
from oracle import binding as ob
yy, xx = np.mgrid[0:16, 0:16]
def texture(holes):
    tex = np.zeros((16, 16, 4), np.uint8); tex[..., :3] = 200
    tex[..., 3] = np.where((xx // 4 + yy // 4) % 2 == 0, 255, 0) if holes else 255
    return tex
back = make_material(base_color=(0.2, 0.6, 0.9))
wall = _quad((-1, 0, 0), (1, 0, 0), (1, 2, 0), (-1, 2, 0))
wall_flipped = _quad((-1, 0, 0), (-1, 2, 0), (1, 2, 0), (1, 0, 0))
behind = [(t, 1) for t in _quad((-2, -1, -1), (2, -1, -1), (2, 3, -1), (-2, 3, -1))]
W = H = 64
c = dict(scenes.CORNELL_CAMERA)
cam = Camera(api.camera_walk(c["origin"], 0.0, 0.0, 0, 0.0), c["fov"], c["F"], c["focus"])
r = Renderer(0); o = ob.Oracle()
def run(name, s, cases=((1, 1), (1, 2), (4, 4))):
    r.set_scene(s); r.build_accel(); r.set_resolution(W, H)
    o.set_scene(s); o.build_accel(); o.set_resolution(W, H)
    for spp, depth in cases:
        layers = DeviceLayers(W, H); r.init_render_states(); r.reset_statistics(); r.render(cam, (1, 1, 1), layers, spp, depth); r.wait()
        o.init_render_states(); o.reset_ray_counts()
        ref, _ = o.render_canonical(cam, (1, 1, 1), spp, depth, n_threads=16)
        g = layers.download("beauty")[..., :3]; q = ref["beauty"][..., :3]
        d = np.abs(g - q).max(axis=2); bad = np.argwhere(d > 1e-3)
        st = r.statistics(); rc = o.ray_counts()
        print("%-28s spp %d depth %d relMSE %.2e bad %4d | rays gpu %d/%d/%d ref %d/%d/%d | mean %.4f %.4f" % (
            name, spp, depth, float(np.mean((g - q) ** 2 / (q ** 2 + 1e-2))), len(bad), st["rays_radiance"], st["rays_shadow"], st["rays_light"],
            rc["rays_radiance"], rc["rays_shadow"], rc["rays_light"], g.mean(), q.mean()))
        for y, x in bad[:3]:
            print("      px", y, x, g[y, x], q[y, x])
plain = make_material(base_color=(0.8, 0.8, 0.8))
run("front face, plain", _assemble([[(t, 0) for t in wall]], [plain, back]))
run("back face, plain", _assemble([[(t, 0) for t in wall_flipped]], [plain, back]))
run("wall + quad, plain", _assemble([[(t, 0) for t in wall], behind], [plain, back]))
for holes in (False, True):
    m = make_material(base_color=(0.8, 0.8, 0.8), base_color_texture_id=0)
    s = _assemble([[(t, 0) for t in wall], behind], [m, back]); s.textures = [(texture(holes), True)]
    run("wall + quad, tex holes=%s" % holes, s)
    s = _assemble([[(t, 0) for t in wall]], [m, back]); s.textures = [(texture(holes), True)]
    run("wall only, tex holes=%s" % holes, s)
