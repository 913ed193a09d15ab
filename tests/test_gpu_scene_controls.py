"""Renderer controls around the path that the other suites do not touch: sky intensity,
clearing the sun, re-uploading instance transforms (what set_time does, renderer.h:614-640), and handing a
loaded fredholm::Scene object to the renderer (load_scene path, renderer.h:354-432) -- each against the
reference integrator on the same inputs."""
import os

import numpy as np
import pytest

from conftest import rel_mse
from fredholm_b200 import Camera, DeviceLayers, api, scenes

pytestmark = pytest.mark.gpu

W, H = 96, 54


def camera():
    c = scenes.STANDARD_CAMERA
    return Camera(api.camera_walk(c["origin"], 0.0, 150.0, 0, 0.0), c["fov"], c["F"], c["focus"])


def render(r, cam, spp=8, depth=5, bg=(0, 0, 0)):
    layers = DeviceLayers(W, H)
    r.init_render_states()
    r.render(cam, bg, layers, spp, depth)
    r.wait()
    out = layers.download("beauty")[..., :3].copy()
    layers.free()
    return out


def test_sky_intensity_and_clear_sun(renderer, oracle):
    s, cam = scenes.standard_surface_scene(48, 24, sphere_res=(12, 6)), camera()
    L = scenes.STANDARD_LIGHTING
    for x in (renderer, oracle):
        x.set_scene(s)
        x.build_accel()
        x.set_resolution(W, H)
        x.set_directional_light(L["sun_le"], L["sun_dir"], L["sun_angle"])
        x.load_arhosek_sky(L["turbidity"], L["albedo"])
        x.set_sky_intensity(2.5)
    got = render(renderer, cam)
    ref, _ = oracle.render_canonical(cam, (0, 0, 0), 8, 5, n_threads=os.cpu_count() or 1)
    assert rel_mse(got, ref["beauty"][..., :3]) < 1e-3
    renderer.set_sky_intensity(1.0)
    dimmer = render(renderer, cam)
    assert dimmer.mean() < got.mean()
    # without the sun: no directional NEE strategy at all (pt.cu:770-794), one 2-D draw fewer per bounce
    for x in (renderer, oracle):
        x.clear_directional_light()
        x.set_sky_intensity(1.0)
    no_sun = render(renderer, cam)
    oracle.init_render_states()
    ref2, _ = oracle.render_canonical(cam, (0, 0, 0), 8, 5, n_threads=os.cpu_count() or 1)
    assert rel_mse(no_sun, ref2["beauty"][..., :3]) < 1e-3
    assert no_sun.mean() < dimmer.mean()


def test_set_transforms_moves_instances(renderer, oracle):
    """New object-to-world matrices for the sub-meshes, then a rebuild: hits, ids and the image follow."""
    s, cam = scenes.standard_surface_scene(32, 16, sphere_res=(12, 6)), camera()
    for x in (renderer, oracle):
        x.set_scene(s)
        x.build_accel()
        x.set_resolution(W, H)
        x.load_arhosek_sky(3.0, 0.3)
    rng = np.random.default_rng(9)
    rays = np.concatenate([rng.uniform(-12, 12, (40000, 3)), rng.normal(size=(40000, 3))], 1).astype(np.float32)
    rays[:, 1] = np.abs(rays[:, 1]) * 0.3 + 0.3
    rays[:, 3:] /= np.linalg.norm(rays[:, 3:], axis=1, keepdims=True)
    before, _ = renderer.trace_closest(rays)
    tr = s.transforms.copy().reshape(-1, 4, 4)
    for i in range(1, len(tr)):                       # sub-mesh 0 (terrain) stays; the others hop and grow
        m = np.eye(4, dtype=np.float32)
        m[:3, :3] *= 1.0 + 0.25 * (i % 3)
        m[:3, 3] = (0.3 * (i % 5), 0.6 + 0.1 * (i % 4), -0.2 * (i % 7))
        tr[i] = (m @ tr[i].T).T                       # column-major storage
    for x in (renderer, oracle):
        x.set_transforms(tr.reshape(-1, 16))
        x.build_accel()
    ids_g, tuv_g = renderer.trace_closest(rays)
    ids_o, tuv_o = oracle.trace_closest(rays)
    assert np.array_equal(ids_g, ids_o)
    hit = ids_o[:, 0] != 0xffffffff
    assert np.array_equal(tuv_g[hit].view(np.uint32), tuv_o[hit].view(np.uint32))
    assert (ids_g != before).any()
    got = render(renderer, cam)
    ref, _ = oracle.render_canonical(cam, (0, 0, 0), 8, 5, n_threads=os.cpu_count() or 1)
    assert rel_mse(got, ref["beauty"][..., :3]) < 1e-3


def test_scene_object_path_equals_arrays(renderer, tmp_path):
    """fr_scene_load + fr_set_scene (what Renderer::load_scene does) renders the same image as the
    flat-array upload of the same scene."""
    s, cam = scenes.cornell_box(), None
    c = scenes.CORNELL_CAMERA
    cam = Camera(api.camera_walk(c["origin"], 0.0, 0.0, 0, 0.0), c["fov"], c["F"], c["focus"])
    path = scenes.write_obj(s, str(tmp_path), "cornell")
    sc = api.Scene()
    sc.load_model(path)
    renderer.set_scene_object(sc)
    renderer.build_accel()
    renderer.set_resolution(W, H)
    a = render(renderer, cam, spp=4, depth=4)
    renderer.set_scene(sc.arrays())
    renderer.build_accel()
    b = render(renderer, cam, spp=4, depth=4)
    sc.close()
    assert np.array_equal(a, b)
    assert a.mean() > 0.01


def test_invalid_input_is_rejected_and_the_renderer_stays_usable(renderer):
    """Round-2 hardening: a scene with an out-of-range index never reaches the device (Scene::validate in
    upload_scene), max_depth beyond the Sobol table is refused, and the renderer keeps working afterwards."""
    import copy
    good = scenes.cornell_box()
    bad = copy.deepcopy(good)
    bad.indices[5, 2] = len(bad.vertices) + 7
    with pytest.raises(api.FredholmError, match="invalid scene: vertex index out of range"):
        renderer.set_scene(bad)
    bad = copy.deepcopy(good)
    bad.materials["normalmap_texture_id"][0] = 3          # no textures in this scene
    with pytest.raises(api.FredholmError, match="invalid scene: texture id 3 out of range"):
        renderer.set_scene(bad)
    renderer.set_scene(good)
    renderer.build_accel()
    W = H = 32
    renderer.set_resolution(W, H)
    c = scenes.CORNELL_CAMERA
    cam = Camera(api.camera_walk(c["origin"], 0.0, 0.0, 0, 0.0), c["fov"], c["F"], c["focus"])
    lay = DeviceLayers(W, H, names=("beauty",))
    with pytest.raises(api.FredholmError, match="max_depth must be <= 255"):
        renderer.render(cam, (0, 0, 0), lay, 1, 256)
    renderer.render(cam, (0, 0, 0), lay, 4, 255)          # the deepest path the sampler table supports
    renderer.wait()
    img = lay.download("beauty")
    assert np.isfinite(img).all() and img[..., :3].mean() > 0.01


def test_wave_compaction_in_a_closed_scene(renderer, oracle):
    """Cornell box with an area light (k_trace_light, three NEE strategies): paths stay alive, so with waves of one
    sample most waves keep more paths than the straggler set holds and finish in place; all six layers are
    bit-identical to the one-wave render, and the image matches the reference integrator."""
    s = scenes.cornell_box()
    c = scenes.CORNELL_CAMERA
    cam = Camera(api.camera_walk(c["origin"], 0.0, 0.0, 0, 0.0), c["fov"], c["F"], c["focus"])
    names = ("beauty", "position", "normal", "depth", "texcoord", "albedo")
    renderer.set_scene(s)
    renderer.build_accel()
    renderer.clear_directional_light()
    renderer.clear_arhosek_sky()
    renderer.set_resolution(W, H)
    slots_per_sample = ((W + 7) // 8) * ((H + 3) // 4) * 32

    def run(wave, on, depth):
        renderer.set_wave_compaction(on, depth)
        renderer.set_max_wave_paths(wave)
        layers = DeviceLayers(W, H, names=names)
        renderer.init_render_states()
        renderer.render(cam, (0, 0, 0), layers, 12, 8)
        renderer.wait()
        out = {n: layers.download(n) for n in names}
        layers.free()
        return out

    try:
        whole = run(1 << 26, True, 0)
        for on, depth in ((False, 0), (True, 1), (True, 3), (True, 7)):
            got = run(slots_per_sample, on, depth)
            for n in names:
                assert np.array_equal(whole[n], got[n], equal_nan=True), (on, depth, n)
    finally:
        renderer.set_wave_compaction(True)
        renderer.set_max_wave_paths(1 << 26)
    oracle.set_scene(s)
    oracle.build_accel()
    oracle.set_resolution(W, H)
    ref, _ = oracle.render_canonical(cam, (0, 0, 0), 12, 8, n_threads=os.cpu_count() or 1)
    assert rel_mse(whole["beauty"][..., :3], ref["beauty"][..., :3]) < 1e-3
