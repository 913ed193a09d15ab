"""Denoise stage and multi-frame batch driver on the B200.

* fr_denoise vs its numpy restatement (oracle/denoise_np.py) on real AOV layers, and as a
  quality gate against a converged render;
* fr_batch_run (fredholm::FrameBatch, the rtcamp8 render/save loop as a pipelined library
  call) vs the same frames produced one call at a time through the reference-style API
  sequence set_time -> render -> denoise -> post_process -> read-back -> host RGBA8
  conversion (app/rtcamp8.cpp:165-282): bit-exact, also through the PNG files it writes,
  and when the frames are split over two "ranks"."""
import ctypes as C

import numpy as np
import pytest

from conftest import rel_mse
from fredholm_b200 import Camera, DeviceLayers, Renderer, api, scenes
from oracle import denoise_np as dn

pytestmark = pytest.mark.gpu

W, H = 96, 80   # multiples of 16: the post-process grid covers the whole image


def cornell_camera():
    c = scenes.CORNELL_CAMERA
    return Camera(api.camera_walk(c["origin"], 0.0, 0.0, 0, 0.0), c["fov"], c["F"], c["focus"])


def render_layers(r, cam, spp, depth=6):
    layers = DeviceLayers(W, H)
    r.init_render_states()
    r.render(cam, (0, 0, 0), layers, spp, depth)
    r.wait()
    return layers


def dev_alloc(nbytes):
    p = api.lib().fr_device_alloc(nbytes)
    assert p
    api._check(api.lib().fr_device_memset(p, 0, nbytes))
    return p


def host_f4(p, h, w):
    out = np.empty((h, w, 4), np.float32)
    api._check(api.lib().fr_copy_to_host(out.ctypes.data_as(C.c_void_p), p, out.nbytes))
    return out


@pytest.fixture()
def cornell(renderer):
    renderer.set_scene(scenes.cornell_box())
    renderer.build_accel()
    renderer.set_resolution(W, H)
    return renderer


@pytest.mark.parametrize("upscale", [False, True])
def test_denoise_matches_numpy_restatement(cornell, upscale):
    layers = render_layers(cornell, cornell_camera(), 4)
    ow, oh = (2 * W, 2 * H) if upscale else (W, H)
    out = dev_alloc(ow * oh * 16)
    api.denoise(layers.ptr["beauty"], layers.ptr["normal"], layers.ptr["albedo"], out, W, H, upscale=upscale)
    got = host_f4(out, oh, ow)
    ref = dn.atrous(layers.download("beauty"), layers.download("normal"), layers.download("albedo"))
    if upscale:
        ref = dn.upscale2x(ref)
    api.lib().fr_device_free(out)
    layers.free()
    assert np.isfinite(got).all()
    assert np.allclose(got, ref, rtol=1e-4, atol=1e-5), np.abs(got - ref).max()


def test_denoise_parameters_and_errors(cornell):
    layers = render_layers(cornell, cornell_camera(), 2)
    out = dev_alloc(W * H * 16)
    b, n, a = (layers.download(k) for k in ("beauty", "normal", "albedo"))
    for kw in (dict(iterations=1), dict(iterations=3, sigma_color=0.2, sigma_albedo=0.2, albedo_floor=0.05),
               dict(firefly_k=0.0), dict(iterations=2, firefly_k=1.0)):
        api.denoise(layers.ptr["beauty"], layers.ptr["normal"], layers.ptr["albedo"], out, W, H, **kw)
        ref = dn.atrous(b, n, a, **{k: v for k, v in kw.items()})
        assert np.allclose(host_f4(out, H, W), ref, rtol=1e-4, atol=1e-5)
    with pytest.raises(api.FredholmError):
        api.denoise(layers.ptr["beauty"], None, layers.ptr["albedo"], out, W, H)
    with pytest.raises(api.FredholmError):
        api.denoise(layers.ptr["beauty"], layers.ptr["normal"], layers.ptr["albedo"], out, 0, H)
    with pytest.raises(api.FredholmError):
        api.denoise(layers.ptr["beauty"], layers.ptr["normal"], layers.ptr["albedo"], out, W, H, iterations=99)
    api.lib().fr_device_free(out)
    layers.free()


def test_denoise_moves_a_noisy_render_towards_the_converged_one(cornell):
    cam = cornell_camera()
    conv = render_layers(cornell, cam, 2048)
    ref = conv.download("beauty")[..., :3]
    conv.free()
    noisy = render_layers(cornell, cam, 8)
    out = dev_alloc(W * H * 16)
    api.denoise(noisy.ptr["beauty"], noisy.ptr["normal"], noisy.ptr["albedo"], out, W, H)
    den = host_f4(out, H, W)[..., :3]
    e_noisy = rel_mse(noisy.download("beauty")[..., :3], ref)
    e_den = rel_mse(den, ref)
    api.lib().fr_device_free(out)
    noisy.free()
    assert e_den < 0.5 * e_noisy, (e_noisy, e_den)


# ---- batch driver ---------------------------------------------------------------------------
ANIM = [dict(node=0, translation=([0.0, 0.5], [(0, 0, 0), (0.3, 0.1, 0)])),
        dict(node=1, rotation=([0.0, 0.25, 0.5], [(0, 0, 0, 1), (0, 0.3826834, 0, 0.9238795), (0, 0.7071068, 0, 0.7071068)]))]
POST = dict(use_bloom=True, bloom_threshold=0.8, bloom_sigma=3.0, ISO=100.0, chromatic_aberration=1.0)
FPS = 24.0
SPP, DEPTH = 3, 4


def quantize_like_the_reference(img):
    """app/rtcamp8.cpp:268-280"""
    out = np.clip(np.float32(255.0) * img, np.float32(0), np.float32(255)).astype(np.uint8)
    out[..., 3] = 255
    return out


def sequential_frames(r, cam_path, frame_ids, denoise=True):
    """One frame at a time through the individual calls, reference order."""
    frames = []
    bufs = [dev_alloc(W * H * 16) for _ in range(4)]   # denoised, high, temp, out
    layers = DeviceLayers(W, H)
    for f in frame_ids:
        t = np.float32(0.0)
        for _ in range(f):
            t = np.float32(t + np.float32(1.0) / np.float32(FPS))
        layers.clear()
        r.init_render_states()
        r.set_time(float(t))
        cam = cornell_camera()
        cam.transform = cam_path[min(f, len(cam_path) - 1)]
        r.render(cam, (0, 0, 0), layers, SPP, DEPTH)
        r.wait()
        src = layers.ptr["beauty"]
        if denoise:
            api.denoise(layers.ptr["beauty"], layers.ptr["normal"], layers.ptr["albedo"], bufs[0], W, H)
            src = bufs[0]
        api.post_process(src, bufs[1], bufs[2], W, H, bufs[3], **POST)
        api._check(api.lib().fr_device_synchronize())
        frames.append(quantize_like_the_reference(host_f4(bufs[3], H, W)))
    for b in bufs:
        api.lib().fr_device_free(b)
    layers.free()
    return np.stack(frames)


@pytest.fixture()
def animated(renderer, tmp_path):
    p = scenes.write_gltf(scenes.cornell_box(), str(tmp_path), "anim", animations=ANIM)
    renderer.load_scene(p)
    renderer.build_accel()
    renderer.set_resolution(W, H)
    c = scenes.CORNELL_CAMERA
    path = np.stack([api.camera_walk(c["origin"], 2.0 * i, 0.0, 0, 0.0) for i in range(6)])
    return renderer, path, tmp_path


def test_batch_matches_frame_by_frame_calls(animated):
    r, path, tmp = animated
    out_dir = tmp / "frames"
    recs, frames, info = r.batch_run(cornell_camera(), W, H, SPP, DEPTH, 6, camera_path=path, output_dir=out_dir,
                                     fps=FPS, max_time=10.0, **POST)
    assert info["n_frames"] == 6 and not info["killed"] and (info["out_width"], info["out_height"]) == (W, H)
    assert [x["frame_idx"] for x in recs] == list(range(6))
    assert np.allclose([x["time"] for x in recs], np.arange(6) / FPS, atol=1e-6)
    ref = sequential_frames(r, path, range(6))
    assert np.array_equal(frames, ref)
    assert len({f.tobytes() for f in frames}) == 6            # animation + camera path: frames differ
    for i, x in enumerate(recs):
        assert x["render_ms"] > 0 and x["denoise_ms"] > 0 and x["post_ms"] > 0 and x["transfer_ms"] > 0
        assert x["png_bytes"] > 0
        png = api.load_image8(out_dir / ("%d.png" % i))[::-1]   # loader flips like fredholm::Texture
        assert np.array_equal(png, frames[i])


def test_batch_without_denoise_and_time_limits(animated):
    r, path, _ = animated
    recs, frames, info = r.batch_run(cornell_camera(), W, H, SPP, DEPTH, 8, camera_path=path, denoise=False,
                                     fps=FPS, max_time=2.5 / FPS, **POST)
    assert info["n_frames"] == 3                               # times 0, 1/24, 2/24 <= max_time < 3/24
    assert np.array_equal(frames, sequential_frames(r, path, range(3), denoise=False))
    assert all(x["png_bytes"] == 0 for x in recs)              # no output_dir: nothing written
    _, _, info = r.batch_run(cornell_camera(), W, H, SPP, DEPTH, 4, kill_time_s=-1.0, fps=FPS, **POST)
    assert info["n_frames"] == 0 and info["killed"]            # watchdog (rtcamp8.cpp:162-166)
    with pytest.raises(api.FredholmError):
        r.batch_run(cornell_camera(), W, H, SPP, DEPTH, 2, frame_stride=0, fps=FPS, **POST)
    with pytest.raises(api.FredholmError):
        r.batch_run(cornell_camera(), W, H, SPP, DEPTH, 2, denoise=False, upscale=True, fps=FPS, **POST)


def test_batch_frames_split_over_two_ranks(animated):
    """C5 sharding: frame f goes to rank f mod G; the union is the single-rank batch."""
    r, path, _ = animated
    _, whole, _ = r.batch_run(cornell_camera(), W, H, SPP, DEPTH, 6, camera_path=path, fps=FPS, max_time=10.0, **POST)
    for rank in range(2):
        recs, part, _ = r.batch_run(cornell_camera(), W, H, SPP, DEPTH, 3, camera_path=path, fps=FPS, max_time=10.0,
                                    first_frame=rank, frame_stride=2, n_slots=1, n_save_threads=1, **POST)
        assert [x["frame_idx"] for x in recs] == [rank, rank + 2, rank + 4]
        assert np.array_equal(part, whole[rank::2])


def test_batch_upscale_output_size(animated):
    r, path, _ = animated
    recs, frames, info = r.batch_run(cornell_camera(), W, H, SPP, DEPTH, 2, camera_path=path, upscale=True, fps=FPS,
                                     **POST)
    assert (info["out_width"], info["out_height"]) == (2 * W, 2 * H) and frames.shape == (2, 2 * H, 2 * W, 4)
    assert frames[..., :3].mean() > 5 and (frames[..., 3] == 255).all()
