"""The drop-in claim as a build: the reference's batch application app/rtcamp8.cpp, compiled UNCHANGED against
this repo's include/ (fredholm/renderer.h, fredholm/denoiser.h, cwl/buffer.h, cwl/util.h, optwl/optwl.h,
kernels/post-process.h) and linked with libfredholm_b200.so (oracle/Makefile: rtcamp8_b200).  The binaries are
built where /root/reference exists and travel to the GPU box in oracle/_ref/ like the oracle library.

  CPU   the unchanged application builds, links and -- with no CUDA device -- stops at its first CUDA call
        (rtcamp8.cpp:75) with the reference's CUDA_CHECK message.
  GPU   the same application with ONE edit (max_time 9.5 s -> 0.05 s: two frames instead of 229) runs its whole
        loop -- load .obj + camera .gltf, build_gas / build_ias, set_time, render 16 spp at 1080p, denoise,
        post-process x2, copy_from_device_to_host, PNG saver thread -- in a directory laid out like the
        reference's (../resources/rtcamp8/..., ./output/)."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from fredholm_b200 import api, scenes

APP = os.path.join(ROOT, "oracle", "_ref", "rtcamp8_b200")
APP2 = os.path.join(ROOT, "oracle", "_ref", "rtcamp8_b200_2frames")
REF_APP = "/root/reference/app/rtcamp8.cpp"


def build_apps():
    if os.path.exists(REF_APP):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "_ref/rtcamp8_b200", "_ref/rtcamp8_b200_2frames"],
                       check=True, capture_output=True)
    if not (os.path.exists(APP) and os.path.exists(APP2)):
        pytest.skip("reference application not built (needs /root/reference at build time)")


def test_reference_app_compiles_unchanged_and_fails_loudly_without_gpu(tmp_path):
    build_apps()
    # the edited copy differs from the reference source in exactly one line
    if os.path.exists(REF_APP):
        a = open(REF_APP).read().split("\n")
        b = open(os.path.join(ROOT, "oracle", "_ref", "patched", "rtcamp8_2frames.cpp")).read().split("\n")
        diff = [(x, y) for x, y in zip(a, b) if x != y]
        assert len(a) == len(b) and len(diff) == 1 and "max_time" in diff[0][0]
    if api.lib().fr_device_count() > 0:
        pytest.skip("CUDA device present")
    p = subprocess.run([APP], cwd=tmp_path, capture_output=True, text=True)
    assert p.returncode != 0
    assert "CUDA call (cudaFree(0) ) failed" in p.stderr and "rtcamp8.cpp:75" in p.stderr


def write_camera_gltf(path, translation):
    doc = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}],
           "nodes": [{"name": "camera", "camera": 0, "translation": list(translation)}],
           "cameras": [{"type": "perspective", "perspective": {"yfov": 1.0, "znear": 0.1, "aspectRatio": 1.7778}}]}
    with open(path, "w") as f:
        json.dump(doc, f)


def test_camera_only_gltf_appends_to_a_loaded_scene(tmp_path):
    """rtcamp8.cpp:120-121: load_scene(obj) then load_scene(camera.gltf, clear=false)."""
    obj = scenes.write_obj(scenes.cornell_box(), str(tmp_path), "rtcamp8")
    write_camera_gltf(tmp_path / "cam.gltf", (0.0, 1.0, 3.4))
    sc = api.Scene()
    sc.load_model(obj)
    n = sc.arrays().n_faces
    sc.load_model(str(tmp_path / "cam.gltf"), clear=False)
    sc.validate()
    assert sc.arrays().n_faces == n


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_reference_app_renders_two_frames(tmp_path):
    build_apps()
    res = tmp_path / "resources" / "rtcamp8"
    run = tmp_path / "build"
    os.makedirs(res)
    os.makedirs(run / "output")
    scenes.write_obj(scenes.cornell_box(), str(res), "rtcamp8")
    write_camera_gltf(res / "rtcamp8_camera.gltf", (0.0, 1.0, 3.4))
    p = subprocess.run([APP2], cwd=run, capture_output=True, text=True, timeout=500)
    assert p.returncode == 0, (p.stdout[-2000:], p.stderr[-2000:])
    frames = sorted(os.listdir(run / "output"))
    assert frames == ["0.png", "1.png"], frames
    for f in frames:
        img = api.load_image8(str(run / "output" / f))
        assert img.shape == (1080, 1920, 4)
        assert img[..., :3].mean() > 8 and img[..., :3].std() > 5 and (img[..., 3] == 255).all()
    assert "[Render] rendering frame: 1" in p.stdout + p.stderr
