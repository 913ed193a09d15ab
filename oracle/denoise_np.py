"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by fredholm_b200/).

numpy restatement of the denoise stage (fredholm_b200/csrc/denoiser.cu).  PARITY UNPINNED
against the reference: the reference's stage is the proprietary OptiX AI denoiser
(fredholm/include/fredholm/denoiser.h:14-145), which cannot run without OptiX and has no
published arithmetic, so this oracle pins the replacement filter only -- an edge-avoiding
a-trous wavelet filter (Dammertz, Sewtz, Hanika, Lensch 2010) guided by the same normal and
albedo AOV layers the reference hands to OptiX (denoiser.h:74-83).
fp32 throughout; differs from the CUDA kernel only by FMA contraction and expf ulps.
"""
import numpy as np

f32 = np.float32
KW = np.array([1 / 16, 1 / 4, 3 / 8, 1 / 4, 1 / 16], dtype=f32)


def _compress(c):
    """range domain of the colour weight: log(1 + c) per channel (relative differences, so that
    emitters and fireflies do not leak into their surroundings)"""
    return np.log1p(np.maximum(c, f32(0)), dtype=f32)


def _pow64(x):
    for _ in range(6):
        x = x * x
    return x


def firefly_clamp(c, n, a, k, inv_sa2):
    """Pass-0 outlier suppression: a pixel whose brightest channel exceeds k x the brightest
    channel among its (up to 8) direct neighbours on the same surface (normal x albedo weight
    >= 0.5) is scaled down to that limit (+1e-3)."""
    m = np.full(c.shape[:2], f32(-1), dtype=f32)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx == 0 and dy == 0:
                continue
            cq, valid = _shift(c, dx, dy)
            nq, _ = _shift(n, dx, dy)
            aq, _ = _shift(a, dx, dy)
            nd = np.minimum(np.maximum((n * nq).sum(-1, dtype=f32), f32(0)), f32(1))
            da = a - aq
            g = _pow64(nd) * np.exp(-((da * da).sum(-1, dtype=f32) * inv_sa2), dtype=f32)
            ok = valid & (g >= f32(0.5))
            m = np.where(ok, np.maximum(m, cq.max(-1)), m)
    top = c.max(-1)
    limit = f32(k) * m + f32(1e-3)
    clamp = (m >= 0) & (top > limit)
    scale = np.where(clamp, limit / np.where(clamp, top, f32(1)), f32(1)).astype(f32)
    return c * scale[..., None]


def _shift(img, dx, dy):
    """img[y+dy, x+dx] with validity mask (taps outside the image are skipped)."""
    h, w = img.shape[:2]
    out = np.zeros_like(img)
    ys0, ys1 = max(0, -dy), min(h, h - dy)
    xs0, xs1 = max(0, -dx), min(w, w - dx)
    valid = np.zeros((h, w), dtype=bool)
    if ys1 > ys0 and xs1 > xs0:
        out[ys0:ys1, xs0:xs1] = img[ys0 + dy:ys1 + dy, xs0 + dx:xs1 + dx]
        valid[ys0:ys1, xs0:xs1] = True
    return out, valid


def _unit_or_zero(n):
    d = (n * n).sum(-1, dtype=f32, keepdims=True)
    inv = np.where(d > 0, f32(1) / np.sqrt(np.where(d > 0, d, f32(1)), dtype=f32), f32(0)).astype(f32)
    return n * inv


def atrous(beauty, normal, albedo, iterations=5, sigma_color=0.5, sigma_albedo=0.1, albedo_floor=0.01,
           firefly_k=2.0):
    """beauty/normal/albedo: (H, W, 4) float32.  Returns the (H, W, 4) denoised image."""
    beauty = np.asarray(beauty, dtype=f32)
    n = _unit_or_zero(np.asarray(normal, dtype=f32)[..., :3])
    a = np.asarray(albedo, dtype=f32)[..., :3]
    mod = np.maximum(a, f32(albedo_floor))
    c = beauty[..., :3] / mod
    inv_sa2 = f32(1) / (f32(sigma_albedo) * f32(sigma_albedo))
    if firefly_k > 0:
        c = firefly_clamp(c, n, a, firefly_k, inv_sa2)
    for i in range(iterations):
        step = 1 << i
        sc = f32(sigma_color) / f32(step)
        inv_sc2 = f32(1) / (sc * sc)
        r = _compress(c)
        acc = np.zeros_like(c)
        wsum = np.zeros(c.shape[:2], dtype=f32)
        for dy in range(-2, 3):
            for dx in range(-2, 3):
                cq, valid = _shift(c, dx * step, dy * step)
                nq, _ = _shift(n, dx * step, dy * step)
                aq, _ = _shift(a, dx * step, dy * step)
                rq, _ = _shift(r, dx * step, dy * step)
                if dx == 0 and dy == 0:
                    wn = np.ones(c.shape[:2], dtype=f32)
                else:
                    nd = np.minimum(np.maximum((n * nq).sum(-1, dtype=f32), f32(0)), f32(1))
                    wn = _pow64(nd)
                da = a - aq
                dr = r - rq
                e = (da * da).sum(-1, dtype=f32) * inv_sa2 + (dr * dr).sum(-1, dtype=f32) * inv_sc2
                w = (KW[dx + 2] * KW[dy + 2]) * wn * np.exp(-e, dtype=f32)
                w = np.where(valid, w, f32(0)).astype(f32)
                acc += w[..., None] * cq
                wsum += w
        c = acc / wsum[..., None]
    out = np.empty_like(beauty)
    out[..., :3] = c * mod
    out[..., 3] = beauty[..., 3]
    return out


def upscale2x(img):
    """2x bilinear upscale at pixel centres with clamp-to-edge (k_upscale2x)."""
    img = np.asarray(img, dtype=f32)
    h, w = img.shape[:2]
    ys = (np.arange(2 * h, dtype=f32) + f32(0.5)) * f32(0.5) - f32(0.5)
    xs = (np.arange(2 * w, dtype=f32) + f32(0.5)) * f32(0.5) - f32(0.5)
    fy, fx = np.floor(ys), np.floor(xs)
    ty, tx = (ys - fy).astype(f32), (xs - fx).astype(f32)
    y0 = np.clip(fy.astype(int), 0, h - 1)
    y1 = np.clip(fy.astype(int) + 1, 0, h - 1)
    x0 = np.clip(fx.astype(int), 0, w - 1)
    x1 = np.clip(fx.astype(int) + 1, 0, w - 1)
    w00 = ((1 - tx)[None, :] * (1 - ty)[:, None])[..., None].astype(f32)
    w10 = (tx[None, :] * (1 - ty)[:, None])[..., None].astype(f32)
    w01 = ((1 - tx)[None, :] * ty[:, None])[..., None].astype(f32)
    w11 = (tx[None, :] * ty[:, None])[..., None].astype(f32)
    return (w00 * img[y0][:, x0] + w10 * img[y0][:, x1] + w01 * img[y1][:, x0] + w11 * img[y1][:, x1]).astype(f32)
