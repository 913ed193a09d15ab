"""Textured shading against the reference integrator: every texture slot of
fill_shading_params (pt.cu:181-280), height-map bump and normal map (pt.cu:709-742), emission textures
on area lights (pt.cu:125-139, 282-322, 859-889), sRGB decode of COLOR textures, wrap addressing and
bilinear filtering (cwl/texture.h:35-47) -- image and first-hit AOVs of a small scene on the same inputs."""
import os

import numpy as np
import pytest

from conftest import rel_mse
from fredholm_b200 import Camera, DeviceLayers, api, scenes
from fredholm_b200.scenes import _assemble, _box, _quad, procedural_textures
from fredholm_b200.types import make_material

pytestmark = pytest.mark.gpu

W, H = 96, 96


def gray(img):
    g = img[..., :1].repeat(4, axis=-1).copy()
    g[..., 3] = 255
    return g


def build_scene(lights=True):
    base, rough, normal = [t[0] for t in procedural_textures(64, seed=0xA7)]
    base2 = procedural_textures(32, seed=0x11)[0][0]
    stripes = np.zeros((32, 32, 4), np.uint8)
    stripes[..., :3] = np.where((np.arange(32) // 4 % 2 == 0)[None, :, None], 255, 40)
    stripes[..., 3] = 255
    glow = np.zeros((16, 16, 4), np.uint8)
    glow[..., 0] = 255
    glow[..., 1] = np.linspace(40, 255, 16).astype(np.uint8)[None, :]
    glow[..., 2] = 30
    glow[..., 3] = 255
    mr = np.zeros((32, 32, 4), np.uint8)                     # glTF packing: g = roughness, b = metalness
    mr[..., 1] = rough[::2, ::2, 0]
    mr[..., 2] = stripes[..., 0]
    mr[..., 3] = 255
    textures = [(base, True), (gray(rough), False), (normal, False), (gray(base2), False), (stripes, False),
                (glow, True), (mr, False), (base2, True)]
    mats = [
        make_material(base_color_texture_id=0, specular_roughness_texture_id=1, normalmap_texture_id=2),   # floor
        make_material(base_color=(0.7, 0.7, 0.75), heightmap_texture_id=3, specular_roughness=0.35),       # back
        make_material(base_color_texture_id=7, metallic_roughness_texture_id=6),                           # left
        make_material(base_color=(0.8, 0.3, 0.2), coat_texture_id=4, coat_roughness_texture_id=1,
                      specular_color_texture_id=0),                                                       # right
        make_material(base_color=(0.6, 0.6, 0.6), metalness_texture_id=4, specular_roughness=0.15),        # box
        make_material(base_color=(0.5, 0.5, 0.5), specular_color=(0, 0, 0), emission=1.0,
                      emission_color=(6, 6, 6), emission_texture_id=5),                                   # lamp
    ]
    room = []
    room += [(t, 0) for t in _quad((-1, 0, 1), (1, 0, 1), (1, 0, -1), (-1, 0, -1))]
    room += [(t, 1) for t in _quad((-1, 0, -1), (1, 0, -1), (1, 2, -1), (-1, 2, -1))]
    room += [(t, 2) for t in _quad((-1, 0, 1), (-1, 0, -1), (-1, 2, -1), (-1, 2, 1))]
    room += [(t, 3) for t in _quad((1, 0, -1), (1, 0, 1), (1, 2, 1), (1, 2, -1))]
    box = [(t, 4) for t in _box(0.2, 0.1, 0.6, 0.7, 0.6, 25.0)]
    shapes = [room, box]
    if lights:
        shapes.append([(t, 5) for t in _quad((-0.8, 1.9, -0.8), (0.8, 1.9, -0.8), (0.8, 1.9, 0.8), (-0.8, 1.9, 0.8))])
    s = _assemble(shapes, mats)
    # texcoords beyond [0,1] exercise wrap addressing; the loader default only covers half of the texture
    s.texcoords = (s.texcoords * np.float32(2.3) - np.float32(0.4)).astype(np.float32)
    s.textures = textures
    return s


def camera():
    c = scenes.CORNELL_CAMERA
    return Camera(api.camera_walk(c["origin"], 0.0, 0.0, 0, 0.0), c["fov"], c["F"], c["focus"])


@pytest.mark.parametrize("lights", [True, False])
def test_textured_scene_matches_oracle(renderer, oracle, lights):
    s, cam = build_scene(lights), camera()
    spp, depth = 16, 5
    renderer.set_scene(s)
    renderer.build_accel()
    renderer.set_resolution(W, H)
    oracle.set_scene(s)
    oracle.build_accel()
    oracle.set_resolution(W, H)
    bg = (0.0, 0.0, 0.0) if lights else (0.9, 1.0, 1.2)
    layers = DeviceLayers(W, H)
    renderer.render(cam, bg, layers, spp, depth)
    renderer.wait()
    ref, _ = oracle.render_canonical(cam, bg, spp, depth, n_threads=os.cpu_count() or 1)
    assert oracle.n_lights() == (2 if lights else 0)
    # first-hit AOVs are deterministic given the sampler: texture fetch + bump / normal-map frames
    for name, tol in (("albedo", 2e-4), ("normal", 2e-4), ("position", 1e-4), ("texcoord", 1e-4)):
        g, q = layers.download(name)[..., :3], ref[name][..., :3]
        close = np.isclose(g, q, rtol=0, atol=tol).all(axis=-1)
        assert close.mean() >= 0.999, (name, close.mean())
    got, want = layers.download("beauty")[..., :3], ref["beauty"][..., :3]
    assert np.isfinite(got).all()
    err = rel_mse(got, want)
    assert err < 1e-3, err
    assert abs(got.mean() - want.mean()) < 0.01 * want.mean()
    alb = layers.download("albedo")[..., :3]
    assert alb.std() > 0.05                                   # the textures are really being read
