"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by fredholm_b200/).

numpy restatement of the reference's post-process stage, one function per kernel of
fredholm/kernels/src/post-process.cu, including the launch-grid quirk (integer division of
the image size by the 16x16 block: rows/columns beyond 16*floor(n/16) are never written).
fp32 arithmetic throughout; the sRGB encode promotes to double like the reference
(post-process.h:18-29).  Validated on the GPU box against the reference's own kernels
compiled with nvcc (oracle/_ref/libpostprocess_ref.so, tests/test_gpu_post_process.py).
"""
import numpy as np

f32 = np.float32


def covered(width, height):
    """post-process.cu:9-11: blocks = (max(W/16,1), max(H/16,1)) of 16x16 threads."""
    cw = min(max(width // 16, 1) * 16, width)
    ch = min(max(height // 16, 1) * 16, height)
    return cw, ch


def luminance(rgb):
    """post-process.h:13-16"""
    return (rgb[..., 0] * f32(0.2126729) + rgb[..., 1] * f32(0.7151522)) + rgb[..., 2] * f32(0.0721750)


def bloom_kernel_0(beauty, threshold, out):
    """post-process.cu:60-74: keep the pixel if its luminance exceeds the threshold."""
    h, w = beauty.shape[:2]
    cw, ch = covered(w, h)
    b = beauty[:ch, :cw]
    keep = luminance(b[..., :3]) > f32(threshold)
    out[:ch, :cw] = np.where(keep[..., None], b, f32(0))
    return out


def bloom_kernel_1(beauty, high, sigma, out):
    """post-process.cu:76-109: b0 + sum(h * b1) / sum(h), h = exp(-(u^2+v^2) / (2 sigma)),
    33x33 taps, clamp-to-edge addressing."""
    h, w = beauty.shape[:2]
    cw, ch = covered(w, h)
    K = 16
    ys = np.clip(np.arange(ch)[:, None] + np.arange(-K, K + 1)[None, :], 0, h - 1)   # (ch, 33)
    xs = np.clip(np.arange(cw)[:, None] + np.arange(-K, K + 1)[None, :], 0, w - 1)   # (cw, 33)
    acc = np.zeros((ch, cw, 4), np.float64)
    wsum = 0.0
    for iv, v in enumerate(range(-K, K + 1)):
        rows = high[ys[:, iv]]                         # (ch, w, 4)
        for iu, u in enumerate(range(-K, K + 1)):
            hw = np.exp(f32(-(u * u + v * v)) / (f32(2.0) * f32(sigma)), dtype=f32)
            acc += np.float64(hw) * rows[:, xs[:, iu]]
            wsum += float(hw)
    out[:ch, :cw] = (beauty[:ch, :cw].astype(np.float64) + acc / wsum).astype(f32)
    return out


def copy_kernel(src, out):
    """post-process.cu:49-58"""
    h, w = src.shape[:2]
    cw, ch = covered(w, h)
    out[:ch, :cw] = src[:ch, :cw]
    return out


def uchimura(x):
    """post-process.h:76-113 with P=1, a=1, m=0.22, l=0.4, c=1.33, b=0"""
    P, a, m, l, c, b = f32(1.0), f32(1.0), f32(0.22), f32(0.4), f32(1.33), f32(0.0)
    l0 = ((P - m) * l) / a
    S0 = m + l0
    S1 = m + a * l0
    C2 = (a * P) / (P - S1)
    CP = -C2 / P
    t = np.clip((x - f32(0)) / (m - f32(0)), 0, 1).astype(f32)
    smooth = np.where(x < 0, f32(0), np.where(x > m, f32(1), t * t * (f32(3) - f32(2) * t)))
    w0 = f32(1) - smooth
    w2 = np.where(x < m + l0, f32(0), f32(1))
    w1 = f32(1) - w0 - w2
    with np.errstate(invalid="ignore"):
        T = m * np.power(x / m, c, dtype=f32) + b
    S = P - (P - S1) * np.exp(CP * (x - S0), dtype=f32)
    L = m + a * (x - m)
    return (T * w0 + L * w1 + S * w2).astype(f32)


def linear_to_srgb(x):
    """post-process.h:18-29 (double-precision pow, rounded to float on store)"""
    xd = x.astype(np.float64)
    with np.errstate(invalid="ignore"):
        hi = 1.055 * np.power(np.float64(1) * xd, np.float64(f32(1.0) / f32(2.4))) - 0.055
    return np.where(xd < 0.0031308, 12.92 * xd, hi).astype(f32)


def tone_mapping_kernel(src, ISO, chromatic_aberration, out):
    """post-process.cu:111-153: per-channel UV shift, float->int index arithmetic, exposure
    from EV100(aperture 1, shutter 1, ISO), Uchimura, sRGB."""
    h, w = src.shape[:2]
    cw, ch = covered(w, h)
    j, i = np.meshgrid(np.arange(ch), np.arange(cw), indexing="ij")
    uvx = i.astype(f32) / f32(w)
    uvy = j.astype(f32) / f32(h)
    n = f32(w * h)
    dx = (uvx - f32(0.5)) / n * f32(chromatic_aberration)
    dy = (uvy - f32(0.5)) / n * f32(chromatic_aberration)
    flat = src.reshape(-1, 4)
    color = np.zeros((ch, cw, 3), f32)
    for c, k in enumerate((f32(0), f32(1), f32(2))):
        ux = np.clip(uvx - k * dx, f32(0), f32(1)).astype(f32)
        uy = np.clip(uvy - k * dy, f32(0), f32(1)).astype(f32)
        # const int idx = uv.x * width + width * (uv.y * height): float arithmetic, truncation
        idx = (ux * f32(w) + f32(w) * (uy * f32(h))).astype(f32).astype(np.int64)
        color[..., c] = flat[np.clip(idx, 0, w * h - 1), c]
    ev100 = f32(np.log2(np.float64(f32(1.0) * f32(1.0) / f32(1.0)) * 100.0 / np.float64(f32(ISO))))
    exposure = f32(1.0) / f32(1.2 * np.float64(np.power(f32(2.0), ev100, dtype=f32)))
    color = color * exposure
    color = linear_to_srgb(uchimura(color))
    out[:ch, :cw, :3] = color
    out[:ch, :cw, 3] = f32(1.0)
    return out


def post_process(beauty, high, temp, use_bloom, bloom_threshold, bloom_sigma, ISO, chromatic_aberration, out):
    """post_process_kernel_launch, post-process.cu:5-35.  high / temp / out are the caller's
    scratch and output images (their uncovered border keeps whatever it held)."""
    if use_bloom:
        bloom_kernel_0(beauty, bloom_threshold, high)
        bloom_kernel_1(beauty, high, bloom_sigma, temp)
    else:
        copy_kernel(beauty, temp)
    return tone_mapping_kernel(temp, ISO, chromatic_aberration, out)
