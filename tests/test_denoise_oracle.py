"""CPU checks of the denoise-stage oracle (oracle/denoise_np.py, the numpy restatement of
fredholm_b200/csrc/denoiser.cu) and of the batch structures on the C ABI.  The reference's
stage is the OptiX AI denoiser (denoiser.h:14-145): no arithmetic to pin against, so these are
the properties the replacement filter must have."""
import ctypes as C
import os
import subprocess

import numpy as np

from fredholm_b200 import api
from oracle import denoise_np as dn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def guides(h, w, normal=(0, 0, 1), albedo=(0.5, 0.5, 0.5)):
    n = np.zeros((h, w, 4), np.float32)
    n[..., :3] = normal
    a = np.zeros((h, w, 4), np.float32)
    a[..., :3] = albedo
    return n, a


def test_constant_image_is_a_fixed_point():
    n, a = guides(24, 40)
    b = np.full((24, 40, 4), 0.7, np.float32)
    out = dn.atrous(b, n, a)
    assert np.allclose(out, b, rtol=1e-6, atol=1e-7)


def test_noise_on_a_flat_wall_is_reduced():
    rng = np.random.default_rng(1)
    n, a = guides(64, 64)
    clean = np.full((64, 64, 4), 0.4, np.float32)
    noisy = clean.copy()
    noisy[..., :3] += rng.normal(0, 0.08, (64, 64, 3)).astype(np.float32)
    out = dn.atrous(noisy, n, a)
    e0 = np.mean((noisy[..., :3] - 0.4) ** 2)
    e1 = np.mean((out[..., :3] - 0.4) ** 2)
    assert e1 < 0.05 * e0, (e0, e1)
    assert np.array_equal(out[..., 3], noisy[..., 3])        # alpha travels untouched


def test_normal_and_albedo_edges_are_kept():
    rng = np.random.default_rng(2)
    h, w = 32, 64
    n, a = guides(h, w)
    n[:, w // 2:, :3] = (1, 0, 0)                            # a crease at x = w/2
    b = np.zeros((h, w, 4), np.float32)
    b[:, :w // 2, :3] = 0.2
    b[:, w // 2:, :3] = 0.9
    b[..., :3] += rng.normal(0, 0.02, (h, w, 3)).astype(np.float32)
    out = dn.atrous(b, n, a)
    assert abs(out[:, :w // 2, :3].mean() - 0.2) < 0.01 and abs(out[:, w // 2:, :3].mean() - 0.9) < 0.01
    assert out[:, w // 2 - 1, 0].max() < 0.3 and out[:, w // 2, 0].min() > 0.8   # no bleeding over the crease
    # texture detail lives in the albedo: demodulation keeps it sharp
    n2, a2 = guides(h, w)
    a2[:, ::2, :3] = 0.9
    a2[:, 1::2, :3] = 0.1
    b2 = np.zeros((h, w, 4), np.float32)
    b2[..., :3] = a2[..., :3] * 0.5
    out2 = dn.atrous(b2, n2, a2)
    assert np.allclose(out2[..., :3], b2[..., :3], rtol=1e-5, atol=1e-6)


def test_sky_pixels_pass_through():
    """Primary misses carry zero normal / albedo AOVs (pt.cu:745-751 only writes them on a
    hit): every neighbour weight is zero, the analytic sky is left alone."""
    rng = np.random.default_rng(3)
    n = np.zeros((16, 16, 4), np.float32)
    a = np.zeros((16, 16, 4), np.float32)
    b = rng.uniform(0, 5, (16, 16, 4)).astype(np.float32)
    out = dn.atrous(b, n, a)
    assert np.allclose(out, b, rtol=1e-5, atol=1e-6)


def test_upscale_of_constant_and_shape():
    img = np.full((5, 7, 4), 0.3, np.float32)
    up = dn.upscale2x(img)
    assert up.shape == (10, 14, 4) and np.allclose(up, 0.3, atol=1e-7)
    ramp = np.tile(np.arange(8, dtype=np.float32)[None, :, None], (4, 1, 4))
    up = dn.upscale2x(ramp)
    assert np.all(np.diff(up[0, :, 0]) >= 0) and up[0, 0, 0] == 0 and up[0, -1, 0] == 7


def test_batch_structs_match_the_c_header(tmp_path):
    """ctypes mirrors of fr_batch_config / fr_frame_record have the layout gcc gives the header."""
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "fredholm_b200.h"\n'
                   'int main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(fr_batch_config), sizeof(fr_frame_record),'
                   'offsetof(fr_batch_config, output_dir), offsetof(fr_batch_config, bg_color),'
                   'offsetof(fr_frame_record, png_bytes));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert got == [C.sizeof(api._BatchConfig), C.sizeof(api._FrameRecord), api._BatchConfig.output_dir.offset,
                   api._BatchConfig.bg_color.offset, api._FrameRecord.png_bytes.offset]
