"""Image files on the boundary (fredholm::Texture / FloatTexture, scene.cpp:7-67): our PNG, JPEG and
Radiance decoders must hand the renderer exactly the texels the reference's stb_image calls produce
(same RGBA conversion, same flip, same JPEG IDCT / chroma up-sampling / colour conversion).  The test
images are written with Pillow, decoded by the reference loader compiled into the oracle and by
fr_image8_load / fr_imagef_load.  Host only -- no GPU needed."""
import io
import struct
import zlib

import numpy as np
import pytest

from fredholm_b200 import api

PIL = pytest.importorskip("PIL.Image")


def pattern(h, w, channels, seed=1, smooth=True):
    rng = np.random.default_rng(seed)
    if smooth:
        y, x = np.mgrid[0:h, 0:w].astype(np.float64)
        img = np.stack([127 + 120 * np.sin(0.11 * x * (c + 1) + 0.07 * y + c) * np.cos(0.05 * y * (c + 1))
                        for c in range(channels)], axis=-1)
        img += rng.normal(0, 6, img.shape)
    else:
        img = rng.integers(0, 256, (h, w, channels)).astype(np.float64)
    return np.clip(img, 0, 255).astype(np.uint8)


def both8(oracle_mod, path):
    return api.load_image8(path), oracle_mod.load_image8(path)


PNG_MODES = [("L", 1), ("LA", 2), ("RGB", 3), ("RGBA", 4), ("P", 3), ("1", 1)]


@pytest.mark.parametrize("mode,ch", PNG_MODES)
@pytest.mark.parametrize("size", [(1, 1), (7, 5), (64, 33)])
def test_png_modes(oracle_mod, tmp_path, mode, ch, size):
    w, h = size
    a = pattern(h, w, ch, seed=w * 31 + h, smooth=False)
    if mode == "P":
        im = PIL.fromarray(a, "RGB").quantize(colors=min(256, max(2, w * h)))
    elif mode == "1":
        im = PIL.fromarray((a[..., 0] > 127).astype(np.uint8) * 255, "L").convert("1")
    else:
        im = PIL.fromarray(a if ch > 1 else a[..., 0], mode)
    p = tmp_path / ("t_%s.png" % mode)
    im.save(p)
    ours, ref = both8(oracle_mod, p)
    assert ours.shape == ref.shape == (h, w, 4)
    assert np.array_equal(ours, ref)


def test_png_palette_with_alpha(oracle_mod, tmp_path):
    a = pattern(20, 24, 4, smooth=False)
    im = PIL.fromarray(a, "RGBA").quantize(colors=64)      # P mode + tRNS
    p = tmp_path / "pa.png"
    im.save(p)
    ours, ref = both8(oracle_mod, p)
    assert np.array_equal(ours, ref) and ours[..., 3].min() < 255


def _png_bytes(w, h, depth, ctype, rows, extra=b"", interlace=0):
    """Hand-assembled PNG: rows = list of already packed scanlines (filter 0)."""
    def chunk(tag, body):
        return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body))
    raw = b"".join(b"\x00" + r for r in rows)
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, interlace)) + extra +
            chunk(b"IDAT", zlib.compress(raw, 9)) + chunk(b"IEND", b""))


def _trns(body):
    return struct.pack(">I", len(body)) + b"tRNS" + body + struct.pack(">I", zlib.crc32(b"tRNS" + body))


@pytest.mark.parametrize("depth", [1, 2, 4, 8, 16])
def test_png_grey_depths_and_colour_key(oracle_mod, tmp_path, depth):
    w, h = 13, 6
    rng = np.random.default_rng(depth)
    vals = rng.integers(0, 1 << depth, (h, w))
    rows = []
    for y in range(h):
        if depth == 16:
            rows.append(b"".join(struct.pack(">H", int(v)) for v in vals[y]))
        elif depth == 8:
            rows.append(bytes(int(v) for v in vals[y]))
        else:
            bits = "".join(format(int(v), "0%db" % depth) for v in vals[y])
            bits += "0" * (-len(bits) % 8)
            rows.append(bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8)))
    key = int(vals[2, 3])
    for name, extra in (("plain", b""), ("key", _trns(struct.pack(">H", key)))):
        p = tmp_path / ("g%d_%s.png" % (depth, name))
        p.write_bytes(_png_bytes(w, h, depth, 0, rows, extra))
        ours, ref = both8(oracle_mod, p)
        assert np.array_equal(ours, ref), (depth, name)
        if name == "key":
            assert ours[..., 3].min() == 0


def test_png_rgb16_and_filters(oracle_mod, tmp_path):
    a = (pattern(31, 17, 3).astype(np.uint16) << 8) | pattern(31, 17, 3, seed=9, smooth=False)
    rows = [b"".join(struct.pack(">H", int(v)) for v in a[y].reshape(-1)) for y in range(a.shape[0])]
    p = tmp_path / "rgb16.png"
    p.write_bytes(_png_bytes(17, 31, 16, 2, rows, _trns(struct.pack(">HHH", *[int(v) for v in a[4, 5]]))))
    ours, ref = both8(oracle_mod, p)
    assert np.array_equal(ours, ref)
    # Pillow's encoder picks per-row filters adaptively on smooth content: all five filter types
    im = PIL.fromarray(pattern(97, 130, 4), "RGBA")
    p2 = tmp_path / "filters.png"
    im.save(p2, optimize=True)
    ours, ref = both8(oracle_mod, p2)
    assert np.array_equal(ours, ref)


def test_png_interlaced(oracle_mod, tmp_path):
    """Adam7: the seven passes are assembled by hand from an RGB image."""
    w, h = 19, 11
    a = pattern(h, w, 3, smooth=False)
    xs, ys, dx, dy = [0, 4, 0, 2, 0, 1, 0], [0, 0, 4, 0, 2, 0, 1], [8, 8, 4, 4, 2, 2, 1], [8, 8, 8, 4, 4, 2, 2]
    rows = []
    for k in range(7):
        sub = a[ys[k]::dy[k], xs[k]::dx[k]]
        if sub.size:
            rows += [sub[y].tobytes() for y in range(sub.shape[0])]
    p = tmp_path / "adam7.png"
    p.write_bytes(_png_bytes(w, h, 8, 2, rows, interlace=1))
    ours, ref = both8(oracle_mod, p)
    assert np.array_equal(ours, ref)
    assert np.array_equal(ours[::-1, :, :3], a)


JPEG_CASES = [
    dict(mode="RGB", subsampling=0, quality=90),
    dict(mode="RGB", subsampling=1, quality=85),       # 4:2:2 -> horizontal 2x
    dict(mode="RGB", subsampling=2, quality=75),       # 4:2:0 -> 2x2
    dict(mode="RGB", subsampling=2, quality=30, progressive=True),
    dict(mode="RGB", subsampling=0, quality=95, progressive=True),
    dict(mode="L", quality=80),
    dict(mode="L", quality=60, progressive=True),
    dict(mode="CMYK", quality=85),
    dict(mode="RGB", subsampling=2, quality=80, restart_marker_blocks=3),
]


@pytest.mark.parametrize("case", JPEG_CASES, ids=lambda c: "-".join("%s%s" % (k[:4], v) for k, v in c.items()))
@pytest.mark.parametrize("size", [(8, 8), (33, 17), (150, 101), (1, 1)])
def test_jpeg_matches_reference_decoder(oracle_mod, tmp_path, case, size):
    w, h = size
    kw = dict(case)
    mode = kw.pop("mode")
    ch = {"L": 1, "RGB": 3, "CMYK": 4}[mode]
    a = pattern(h, w, ch, seed=w + h)
    im = PIL.fromarray(a if ch > 1 else a[..., 0], mode)
    p = tmp_path / "t.jpg"
    im.save(p, format="JPEG", **kw)
    ours, ref = both8(oracle_mod, p)
    assert ours.shape == ref.shape == (h, w, 4)
    assert np.array_equal(ours, ref)


def test_jpeg_vertical_subsampling(oracle_mod, tmp_path):
    """4:4:0 (1x2) is not offered by Pillow's presets; cv2 can write it when available."""
    cv2 = pytest.importorskip("cv2")
    if not hasattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR"):
        pytest.skip("cv2 without sampling-factor control")
    a = pattern(70, 45, 3)
    p = tmp_path / "v.jpg"
    ok = cv2.imwrite(str(p), a, [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440])
    if not ok:
        pytest.skip("cv2 could not write 4:4:0")
    ours, ref = both8(oracle_mod, p)
    assert np.array_equal(ours, ref)


def _write_hdr(path, img, rle):
    """img: (H, W, 3) float -> Radiance RGBE file (optionally new-style RLE scanlines)."""
    h, w, _ = img.shape
    m = img.max(axis=-1)
    e = np.where(m > 1e-32, np.floor(np.log2(np.maximum(m, 1e-38))) + 1, 0).astype(np.int32)
    scale = np.where(m > 1e-32, np.ldexp(1.0, 8 - e), 0.0)
    rgbe = np.zeros((h, w, 4), np.uint8)
    rgbe[..., :3] = np.clip(img * scale[..., None], 0, 255).astype(np.uint8)
    rgbe[..., 3] = np.where(m > 1e-32, e + 128, 0).astype(np.uint8)
    out = io.BytesIO()
    out.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=1.0\n\n-Y %d +X %d\n" % (h, w))
    for y in range(h):
        if not rle:
            out.write(rgbe[y].tobytes())
            continue
        out.write(bytes([2, 2, w >> 8, w & 255]))
        for c in range(4):
            row = rgbe[y, :, c]
            i = 0
            while i < w:
                run = 1
                while i + run < w and run < 127 and row[i + run] == row[i]:
                    run += 1
                if run >= 4:
                    out.write(bytes([128 + run, int(row[i])]))
                    i += run
                else:
                    j = i
                    while j < w and j - i < 128:
                        if j + 3 < w and row[j] == row[j + 1] == row[j + 2] == row[j + 3]:
                            break
                        j += 1
                    j = max(j, i + 1)
                    out.write(bytes([j - i]) + row[i:j].tobytes())
                    i = j
    with open(path, "wb") as f:
        f.write(out.getvalue())


@pytest.mark.parametrize("rle,w", [(True, 64), (False, 64), (True, 5)])
def test_hdr_matches_reference(oracle_mod, tmp_path, rle, w):
    rng = np.random.default_rng(3)
    img = rng.gamma(0.6, 2.0, (24, w, 3)) * np.array([1.0, 0.8, 0.6])
    img[3:6, : w // 2] = 0.25          # constant runs
    img[10, :, :] = 0.0                 # zero exponent
    p = tmp_path / "env.hdr"
    _write_hdr(p, img, rle and w >= 8)
    ours, ref = api.load_imagef(p), oracle_mod.load_imagef(p)
    assert ours.shape == ref.shape == (24, w, 4)
    assert np.array_equal(ours, ref)
    # RGBE shares one exponent per pixel: error below one mantissa step of the largest channel
    assert (np.abs(ours[..., :3] - img) <= img.max(axis=-1, keepdims=True) / 128 + 1e-6).all()


def test_ldr_as_environment_map(oracle_mod, tmp_path):
    """stbi_loadf on an 8-bit file: gamma 2.2 on colour, linear alpha (scene.cpp:44-45)."""
    a = pattern(9, 12, 4, smooth=False)
    p = tmp_path / "ldr.png"
    PIL.fromarray(a, "RGBA").save(p)
    ours, ref = api.load_imagef(p), oracle_mod.load_imagef(p)
    assert np.array_equal(ours, ref)


def test_png_writer_round_trip(tmp_path):
    """fr_write_png (the frame savers' stbi_write_png): any PNG reader must get the pixels back."""
    for ch in (3, 4):
        for a in (pattern(37, 53, ch), pattern(16, 16, ch, smooth=False), np.zeros((5, 300, ch), np.uint8)):
            p = tmp_path / ("w%d.png" % ch)
            api.write_png(p, a)
            back = np.asarray(PIL.open(p))
            assert back.shape == a.shape and np.array_equal(back, a)
            assert np.array_equal(api.load_image8(p)[::-1, :, :ch], a)
    big = np.repeat(np.repeat(pattern(27, 48, 3), 10, axis=0), 10, axis=1)     # flat 10x10 blocks
    api.write_png(tmp_path / "big.png", big)
    assert (tmp_path / "big.png").stat().st_size < big.nbytes // 4      # it does compress


def test_bad_files_raise(tmp_path):
    (tmp_path / "a.png").write_bytes(b"\x89PNG\r\n\x1a\n" + b"\x00" * 20)
    (tmp_path / "b.jpg").write_bytes(b"\xff\xd8\xff\xe0" + b"\x00" * 8)
    (tmp_path / "c.bmp").write_bytes(b"BM" + b"\x00" * 64)
    for f in ("a.png", "b.jpg", "c.bmp", "missing.png"):
        with pytest.raises(api.FredholmError):
            api.load_image8(tmp_path / f)
    with pytest.raises(api.FredholmError):
        api.load_imagef(tmp_path / "c.bmp")
