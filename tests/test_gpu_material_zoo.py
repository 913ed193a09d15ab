"""Every lobe of the Standard Surface BSDF through the integrated path: a room of
boxes and open panels, one per material class of tests/bsdf_cases.py (Lambert, Oren-Nayar, dielectric,
glossy, metal, half metal, coat, glass, clear glass, sheen, thin-walled subsurface, everything at once),
lit by an area light, the sun and a constant background -- image against the reference integrator."""
import os

import numpy as np
import pytest

from bsdf_cases import MATERIAL_CLASSES
from conftest import rel_mse
from fredholm_b200 import Camera, DeviceLayers, api, scenes
from fredholm_b200.scenes import _assemble, _box, _quad
from fredholm_b200.types import make_material

pytestmark = pytest.mark.gpu

W, H = 128, 96


def zoo(with_light):
    names = list(MATERIAL_CLASSES)
    mats = [make_material(**MATERIAL_CLASSES[n]) for n in names]
    floor_id = len(mats)
    mats.append(make_material(base_color=(0.6, 0.6, 0.6), specular_color=(0, 0, 0)))
    light_id = len(mats)
    mats.append(make_material(base_color=(0.5, 0.5, 0.5), specular_color=(0, 0, 0), emission=1.0,
                              emission_color=(12, 11, 9)))
    shapes = [[(t, floor_id) for t in _quad((-3, 0, 2), (3, 0, 2), (3, 0, -2), (-3, 0, -2))]]
    for i, n in enumerate(names):
        col, row = i % 6, i // 6
        cx, cz = -2.5 + col * 1.0, -1.0 + row * 1.3
        if n in ("thin_sss", "sheen"):
            # open, two-sided panels: the thin-walled lobes are lit from behind as well
            a, b = (cx - 0.35, 0.05, cz), (cx + 0.35, 0.9, cz + 0.2)
            shapes.append([(t, i) for t in _quad((a[0], a[1], a[2]), (b[0], a[1], a[2]), (b[0], b[1], b[2]),
                                                  (a[0], b[1], b[2]))])
        else:
            shapes.append([(t, i) for t in _box(cx, cz, 0.6, 0.5 + 0.1 * (i % 3), 0.6, 15.0 * i)])
    if with_light:
        shapes.append([(t, light_id) for t in _quad((-1.5, 2.2, -1), (1.5, 2.2, -1), (1.5, 2.2, 1), (-1.5, 2.2, 1))])
    return _assemble(shapes, mats)


def camera():
    return Camera(api.camera_walk((0.0, 2.0, 2.2), 0.0, 330.0, 0, 0.0), np.deg2rad(80.0), 100.0, 10000.0)   # 33 deg down


@pytest.mark.parametrize("mode", ["area_light", "sun_and_background"])
def test_material_zoo_matches_oracle(renderer, oracle, mode):
    s, cam = zoo(mode == "area_light"), camera()
    spp, depth = 32, 6
    bg = (0.0, 0.0, 0.0) if mode == "area_light" else (0.5, 0.6, 0.8)
    for x in (renderer, oracle):
        x.set_scene(s)
        x.build_accel()
        x.set_resolution(W, H)
        if mode != "area_light":
            x.set_directional_light((8, 7.5, 7), (0.3, 1.0, 0.4), 2.0)
    layers = DeviceLayers(W, H)
    renderer.reset_statistics()
    renderer.render(cam, bg, layers, spp, depth)
    renderer.wait()
    oracle.reset_ray_counts()
    ref, _ = oracle.render_canonical(cam, bg, spp, depth, n_threads=os.cpu_count() or 1)
    got, want = layers.download("beauty")[..., :3], ref["beauty"][..., :3]
    assert np.isfinite(got).all()
    err = rel_mse(got, want)
    assert err < 1e-3, err
    # per material: mean radiance over the pixels whose first hit is that material
    rays = oracle.primary_rays(cam, 0).reshape(-1, 6)
    ids, _ = oracle.trace_closest(rays)
    hit = ids[:, 0] != 0xffffffff
    face = np.zeros(len(ids), np.int64)
    face[hit] = s.submesh_offsets[ids[hit, 0]].astype(np.int64) + ids[hit, 1]
    mat = np.where(hit, s.material_ids[face].astype(np.int64), -1).reshape(H, W)
    seen = 0
    for m in range(len(MATERIAL_CLASSES)):
        sel = mat == m
        if sel.sum() < 30:
            continue
        seen += 1
        a, b = got[sel].mean(), want[sel].mean()
        assert abs(a - b) <= 0.02 * max(b, 0.05), (list(MATERIAL_CLASSES)[m], a, b)
    assert seen >= 10
    # the path counts agree: same roulette decisions, same continuation rays
    assert renderer.statistics()["rays_radiance"] == oracle.ray_counts()["rays_radiance"]
