"""The C++ face of the boundary: an application written against include/fredholm/renderer.h
(the reference's class API and call order) compiles, links against libfredholm_b200.so and
-- on the GPU box -- renders a .obj scene end to end (load, BVH build, render, post-process,
read-back)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from fredholm_b200 import api, scenes

EXE = os.path.join(ROOT, "examples", "render_obj")


def build_example():
    cmd = ["g++", "-std=c++17", "-O2", os.path.join(ROOT, "examples", "render_obj.cpp"),
           "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include",
           "-L" + os.path.join(ROOT, "fredholm_b200"), "-lfredholm_b200", "-L/usr/local/cuda/lib64", "-lcudart",
           "-Wl,-rpath," + os.path.join(ROOT, "fredholm_b200"), "-o", EXE]
    subprocess.run(cmd, check=True)


def test_cpp_application_compiles_and_fails_loudly_without_gpu(tmp_path):
    build_example()
    if api.lib().fr_device_count() > 0:
        pytest.skip("CUDA device present")
    p = subprocess.run([EXE, "none.obj", str(tmp_path / "o.ppm")], capture_output=True, text=True)
    assert p.returncode == 1 and "CUDA" in p.stderr


@pytest.mark.gpu
def test_cpp_application_renders(tmp_path):
    build_example()
    obj = scenes.write_obj(scenes.cornell_box(), str(tmp_path), "cornell")
    out = tmp_path / "o.ppm"
    p = subprocess.run([EXE, obj, str(out), "96", "96", "8", "5"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert "paths" in p.stdout and int(p.stdout.split()[0]) == 96 * 96 * 8
    data = out.read_bytes()
    header = b"P6\n96 96\n255\n"
    assert data.startswith(header)
    img = np.frombuffer(data[len(header):], np.uint8).reshape(96, 96, 3)
    assert img.mean() > 5 and img.std() > 1


CWL_EXE = os.path.join(ROOT, "examples", "cwl_check")


def build_cwl_check():
    cmd = ["g++", "-std=c++17", "-O1", os.path.join(ROOT, "examples", "cwl_check.cpp"),
           "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include", "-L/usr/local/cuda/lib64", "-lcudart",
           "-o", CWL_EXE]
    subprocess.run(cmd, check=True)


def test_cwl_headers_compile_and_fail_loudly_without_gpu():
    """include/cwl/{buffer,util,texture}.h (reference cwl/include/cwl/*.h) are header-only over the CUDA runtime."""
    build_cwl_check()
    if api.lib().fr_device_count() > 0:
        pytest.skip("CUDA device present")
    p = subprocess.run([CWL_EXE], capture_output=True, text=True)
    assert p.returncode == 1 and "CUDA call (cudaFree(0) ) failed" in p.stderr


@pytest.mark.gpu
def test_cwl_buffer_object_texture_on_the_gpu():
    build_cwl_check()
    p = subprocess.run([CWL_EXE], capture_output=True, text=True)
    assert p.returncode == 0 and p.stdout.strip() == "ok", p.stderr
