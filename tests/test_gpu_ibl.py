"""IBL environment: Renderer::load_ibl / clear_ibl (renderer.h:574-586), the
lat-long lookup of __miss__radiance / __miss__light and the sky NEE strategy (pt.cu:344-350,
504-543, 796-857), and the priority IBL > Hosek > constant background (pt.cu:511-517) -- compared
with the reference integrator (host oracle) on the same inputs."""
import numpy as np
import pytest

from conftest import rel_mse
from fredholm_b200 import Camera, DeviceLayers, api, scenes
from test_image_codec import _write_hdr

pytestmark = pytest.mark.gpu

W, H = 96, 54
SPP, DEPTH = 8, 6


def make_ibl(w=64, h=32):
    """Procedural lat-long map: horizon gradient, a warm blob and a blue band (float RGBA)."""
    v, u = np.meshgrid((np.arange(h) + 0.5) / h, (np.arange(w) + 0.5) / w, indexing="ij")
    img = np.zeros((h, w, 4), np.float32)
    img[..., 0] = 0.3 + 0.5 * v + 6.0 * np.exp(-((u - 0.3) ** 2 + (v - 0.35) ** 2) / 0.004)
    img[..., 1] = 0.4 + 0.3 * np.cos(6.28318 * u) ** 2 + 4.0 * np.exp(-((u - 0.3) ** 2 + (v - 0.35) ** 2) / 0.004)
    img[..., 2] = 0.9 - 0.6 * v + 0.5 * (np.abs(u - 0.7) < 0.05)
    img[..., 3] = 1.0
    return img


def camera():
    c = scenes.STANDARD_CAMERA
    return Camera(api.camera_walk(c["origin"], 0.0, 150.0, 0, 0.0), c["fov"], c["F"], c["focus"])


def small_scene():
    return scenes.standard_surface_scene(64, 32, sphere_res=(16, 8))


def render(r, cam, bg=(0, 0, 0)):
    layers = DeviceLayers(W, H)
    r.init_render_states()
    r.render(cam, bg, layers, SPP, DEPTH)
    r.wait()
    out = layers.download("beauty")[..., :3].copy()
    depth = layers.download("depth").copy()
    layers.free()
    return out, depth


def test_ibl_image_matches_oracle(renderer, oracle):
    s, cam, ibl = small_scene(), camera(), make_ibl()
    renderer.set_scene(s)
    renderer.build_accel()
    renderer.set_resolution(W, H)
    renderer.set_ibl(ibl)
    got, depth = render(renderer, cam)
    oracle.set_scene(s)
    oracle.build_accel()
    oracle.set_resolution(W, H)
    oracle.set_ibl(ibl)
    ref, _ = oracle.render_canonical(cam, (0, 0, 0), SPP, DEPTH, n_threads=8)
    ref = ref["beauty"][..., :3]
    assert np.isfinite(got).all()
    err = rel_mse(got, ref)
    assert err < 1e-3, err
    # pixels that see the environment directly: pure texture lookups, no Monte-Carlo noise
    sky = depth.reshape(H, W) == 0
    assert 0.02 < sky.mean() < 0.9
    assert np.allclose(got[sky], ref[sky], rtol=1e-4, atol=1e-5)
    assert got[sky].max() > 1.0            # the map is HDR and is not clamped


def test_ibl_priority_and_clear(renderer):
    """IBL wins over the Hosek sky and over bg_color; clear_ibl gives the Hosek sky back."""
    s, cam, ibl = small_scene(), camera(), make_ibl()
    L = scenes.STANDARD_LIGHTING
    renderer.set_scene(s)
    renderer.build_accel()
    renderer.set_resolution(W, H)
    renderer.set_ibl(ibl)
    only_ibl, _ = render(renderer, cam)
    renderer.load_arhosek_sky(L["turbidity"], L["albedo"])
    both, _ = render(renderer, cam, bg=(0.2, 0.9, 0.1))
    assert np.array_equal(only_ibl, both)
    renderer.clear_ibl()
    hosek, _ = render(renderer, cam)
    assert not np.array_equal(hosek, both)
    renderer.clear_arhosek_sky()
    const, depth = render(renderer, cam, bg=(0.25, 0.5, 0.75))
    sky = depth.reshape(H, W) == 0
    assert np.allclose(const[sky], (0.25, 0.5, 0.75), atol=1e-6)


def test_load_ibl_file_equals_texels(renderer, tmp_path):
    """load_ibl(path) decodes the Radiance file to the texels set_ibl would be given (image_codec.cpp,
    texel-exact against stbi_loadf in tests/test_image_codec.py)."""
    s, cam = small_scene(), camera()
    rgb = make_ibl(32, 16)[..., :3].astype(np.float64)
    p = tmp_path / "env.hdr"
    _write_hdr(p, rgb, True)
    renderer.set_scene(s)
    renderer.build_accel()
    renderer.set_resolution(W, H)
    renderer.load_ibl(p)
    a, _ = render(renderer, cam)
    img = api.load_imagef(p)
    assert img.shape == (16, 32, 4)
    renderer.set_ibl(img)
    b, _ = render(renderer, cam)
    assert np.array_equal(a, b)
    with pytest.raises(Exception):
        renderer.load_ibl(tmp_path / "missing.hdr")
