"""ctypes binding of libfredholm_b200.so (C ABI: include/fredholm_b200.h).

Mirrors the reference's call sequence (fredholm::Renderer, renderer.h:29-846):
    r = Renderer(device); r.set_resolution(w, h); r.load_scene(path) / r.set_scene(arrays)
    r.build_accel(); r.render(camera, bg, layers, n_samples, max_depth); r.wait()
There is no CPU path: every call goes to the CUDA library and raises if it is missing.
"""
import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from .types import MATERIAL_DTYPE, SceneArrays

_HERE = os.path.dirname(os.path.abspath(__file__))
# FREDHOLM_B200_LIB: alternative build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("FREDHOLM_B200_LIB") or os.path.join(_HERE, "libfredholm_b200.so")


class LibraryNotBuilt(RuntimeError):
    pass


class FredholmError(RuntimeError):
    pass


class _Layers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("beauty", "position", "depth", "normal", "texcoord", "albedo")]


class _PostProcessParams(C.Structure):
    _fields_ = [("use_bloom", C.c_int), ("bloom_threshold", C.c_float), ("bloom_sigma", C.c_float),
                ("ISO", C.c_float), ("chromatic_aberration", C.c_float)]


class _BatchConfig(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("n_spp", C.c_uint32), ("max_depth", C.c_uint32),
                ("post", _PostProcessParams), ("denoise", C.c_int), ("upscale", C.c_int),
                ("dn_iterations", C.c_int), ("dn_sigma_color", C.c_float), ("dn_sigma_albedo", C.c_float),
                ("dn_albedo_floor", C.c_float), ("dn_firefly_k", C.c_float), ("fps", C.c_float), ("start_time", C.c_float),
                ("max_time", C.c_float), ("kill_time_s", C.c_float), ("first_frame", C.c_uint32),
                ("frame_stride", C.c_uint32), ("max_frames", C.c_uint32), ("bg_color", C.c_float * 3),
                ("animate", C.c_int), ("output_dir", C.c_char_p), ("n_save_threads", C.c_uint32),
                ("n_slots", C.c_uint32)]


class _FrameRecord(C.Structure):
    _fields_ = [("frame_idx", C.c_uint32)] + [(n, C.c_float) for n in (
        "time", "accel_ms", "render_ms", "denoise_ms", "post_ms", "transfer_ms", "encode_ms", "save_ms")] + [
        ("png_bytes", C.c_uint64)]


_fp = C.POINTER(C.c_float)
_up = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)
_u64p = C.POINTER(C.c_uint64)
_vp = C.c_void_p

# name -> (restype, argtypes); this table is also what the "exports every symbol" test walks
SIGNATURES = {
    "fr_last_error": (C.c_char_p, []),
    "fr_device_count": (C.c_int, []),
    "fr_version": (C.c_char_p, []),
    "fr_renderer_create": (_vp, [C.c_int]),
    "fr_renderer_destroy": (None, [_vp]),
    "fr_load_scene": (C.c_int, [_vp, C.c_char_p, C.c_int]),
    "fr_stage_texture": (C.c_int, [_vp, _u8p, C.c_uint32, C.c_uint32, C.c_int]),
    "fr_set_scene_arrays": (C.c_int, [_vp, _fp, _fp, _fp, C.c_uint32, _up, _up, _up, C.c_uint32, _vp, C.c_uint32,
                                      _up, _up, _fp, C.c_uint32]),
    "fr_scene_set_arrays": (C.c_int, [_vp, _fp, _fp, _fp, C.c_uint32, _up, _up, _up, C.c_uint32, _vp, C.c_uint32,
                                      _up, _up, _fp, C.c_uint32]),
    "fr_scene_validate": (C.c_int, [_vp]),
    "fr_get_scene_sizes": (C.c_int, [_vp, _up]),
    "fr_get_scene_arrays": (C.c_int, [_vp, _fp, _fp, _fp, _up, _up, _up, _vp, _up, _up, _fp, _fp]),
    "fr_get_texture_info": (C.c_int, [_vp, C.c_uint32, _up, _up, _up]),
    "fr_get_texture_data": (C.c_int, [_vp, C.c_uint32, _u8p]),
    "fr_scene_create": (_vp, []),
    "fr_scene_destroy": (None, [_vp]),
    "fr_scene_load": (C.c_int, [_vp, C.c_char_p, C.c_int]),
    "fr_scene_get_sizes": (C.c_int, [_vp, _up]),
    "fr_scene_get_arrays": (C.c_int, [_vp, _fp, _fp, _fp, _up, _up, _up, _vp, _up, _up, _fp, _fp]),
    "fr_scene_get_texture_info": (C.c_int, [_vp, C.c_uint32, _up, _up, _up]),
    "fr_scene_get_texture_data": (C.c_int, [_vp, C.c_uint32, _u8p]),
    "fr_scene_update_animation": (C.c_int, [_vp, C.c_float]),
    "fr_set_scene": (C.c_int, [_vp, _vp]),
    "fr_build_accel": (C.c_int, [_vp]),
    "fr_get_accel_info": (C.c_int, [_vp, _up, _fp, _u64p]),
    "fr_set_accel_mode": (C.c_int, [_vp, C.c_int]),
    "fr_get_accel_info2": (C.c_int, [_vp, _up, C.POINTER(C.c_float)]),
    "fr_get_accel_data": (C.c_int, [_vp, _vp, _vp]),
    "fr_set_time": (C.c_int, [_vp, C.c_float]),
    "fr_set_transforms": (C.c_int, [_vp, _fp, C.c_uint32]),
    "fr_set_directional_light": (C.c_int, [_vp, _fp, _fp, C.c_float]),
    "fr_clear_directional_light": (C.c_int, [_vp]),
    "fr_set_sky_intensity": (C.c_int, [_vp, C.c_float]),
    "fr_load_arhosek_sky": (C.c_int, [_vp, C.c_float, C.c_float]),
    "fr_clear_arhosek_sky": (C.c_int, [_vp]),
    "fr_set_ibl": (C.c_int, [_vp, _fp, C.c_uint32, C.c_uint32]),
    "fr_load_ibl": (C.c_int, [_vp, C.c_char_p]),
    "fr_clear_ibl": (C.c_int, [_vp]),
    "fr_set_resolution": (C.c_int, [_vp, C.c_uint32, C.c_uint32]),
    "fr_init_render_states": (C.c_int, [_vp]),
    "fr_set_sample_offset": (C.c_int, [_vp, C.c_uint32]),
    "fr_get_sample_count": (C.c_uint32, [_vp]),
    "fr_set_film_mode": (C.c_int, [_vp, C.c_int]),
    "fr_set_max_wave_paths": (C.c_int, [_vp, C.c_uint64]),
    "fr_set_wave_overlap": (C.c_int, [_vp, C.c_int]),
    "fr_set_wave_compaction": (C.c_int, [_vp, C.c_int, C.c_uint32]),
    "fr_get_wave_state_bytes": (C.c_uint64, [_vp]),
    "fr_set_single_launch": (C.c_int, [_vp, C.c_int]),
    "fr_render": (C.c_int, [_vp, _fp, C.c_float, C.c_float, C.c_float, _fp, C.POINTER(_Layers), C.c_uint32,
                            C.c_uint32]),
    "fr_wait": (C.c_int, [_vp]),
    "fr_render_frame_host": (C.c_int, [_vp, _fp, C.c_float, C.c_float, C.c_float, _fp, C.POINTER(_Layers),
                                       C.c_uint32, C.c_uint32]),
    "fr_scale_layers": (C.c_int, [_vp, C.POINTER(_Layers), C.c_float]),
    "fr_set_device": (C.c_int, [C.c_int]),
    "fr_get_device_attributes": (C.c_int, [C.c_int, _up, _u64p]),
    "fr_comm_get_unique_id": (C.c_int, [_u8p]),
    "fr_comm_init": (C.c_int, [_vp, _u8p, C.c_int, C.c_int]),
    "fr_comm_destroy": (C.c_int, [_vp]),
    "fr_sample_slice": (C.c_int, [C.c_uint32, C.c_int, C.c_int, _up, _up]),
    "fr_render_sharded": (C.c_int, [_vp, _fp, C.c_float, C.c_float, C.c_float, _fp, C.POINTER(_Layers), C.c_uint32,
                                    C.c_uint32, C.c_int]),
    "fr_reduce_layers": (C.c_int, [_vp, C.POINTER(_Layers), C.c_uint32, C.c_int]),
    "fr_multi_create": (_vp, [C.POINTER(C.c_int), C.c_int]),
    "fr_multi_destroy": (None, [_vp]),
    "fr_multi_size": (C.c_int, [_vp]),
    "fr_multi_renderer": (_vp, [_vp, C.c_int]),
    "fr_multi_render": (C.c_int, [_vp, _fp, C.c_float, C.c_float, C.c_float, _fp, C.POINTER(_Layers), C.c_uint32,
                                  C.c_uint32]),
    "fr_multi_wait": (C.c_int, [_vp]),
    "fr_get_statistics": (C.c_int, [_vp, _u64p]),
    "fr_reset_statistics": (C.c_int, [_vp]),
    "fr_set_traversal_counting": (C.c_int, [_vp, C.c_int]),
    "fr_get_traversal_counters": (C.c_int, [_vp, _u64p]),
    "fr_set_samples_per_warp": (C.c_int, [_vp, C.c_uint32]),
    "fr_set_stage_timing": (C.c_int, [_vp, C.c_int]),
    "fr_get_stage_times": (C.c_int, [_vp, C.POINTER(C.c_double), _u64p]),
    "fr_event_create": (_vp, []),
    "fr_event_destroy": (C.c_int, [_vp]),
    "fr_event_record": (C.c_int, [_vp, _vp]),
    "fr_event_elapsed_ms": (C.c_int, [_vp, _vp, _fp]),
    "fr_get_stream": (C.c_uint64, [_vp]),
    "fr_post_process": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.POINTER(_PostProcessParams), _vp]),
    "fr_tone_mapping": (C.c_int, [_vp, C.c_int, C.c_int, C.c_float, C.c_float, _vp]),
    "fr_denoise": (C.c_int, [_vp, _vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_float, C.c_float,
                             C.c_float, C.c_float]),
    "fr_batch_run": (C.c_int, [_vp, C.POINTER(_BatchConfig), _fp, C.c_uint32, C.c_float, C.c_float, C.c_float,
                               C.POINTER(_FrameRecord), C.c_uint32, _u8p, _up, C.POINTER(C.c_double)]),
    "fr_device_alloc": (_vp, [C.c_size_t]),
    "fr_device_free": (C.c_int, [_vp]),
    "fr_device_memset": (C.c_int, [_vp, C.c_int, C.c_size_t]),
    "fr_copy_to_device": (C.c_int, [_vp, _vp, C.c_size_t]),
    "fr_copy_to_host": (C.c_int, [_vp, _vp, C.c_size_t]),
    "fr_device_synchronize": (C.c_int, []),
    "fr_host_alloc_pinned": (_vp, [C.c_size_t]),
    "fr_host_free_pinned": (C.c_int, [_vp]),
    "fr_image8_load": (C.c_int, [C.c_char_p, _up, _up]),
    "fr_image8_copy": (C.c_int, [_u8p]),
    "fr_imagef_load": (C.c_int, [C.c_char_p, _up, _up]),
    "fr_imagef_copy": (C.c_int, [_fp]),
    "fr_write_png": (C.c_int, [C.c_char_p, _u8p, C.c_uint32, C.c_uint32, C.c_uint32]),
    "fr_trace_closest": (C.c_int, [_vp, _fp, C.c_uint32, C.c_float, C.c_float, _up, _fp, _u64p]),
    "fr_primary_rays": (C.c_int, [_vp, _fp, C.c_float, C.c_float, C.c_float, C.c_uint32, _fp]),
    "fr_sampler_sequence": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_char_p, _fp]),
    "fr_bsdf_eval_sample": (C.c_int, [_fp, C.c_uint32, _fp]),
    "fr_sky_radiance": (C.c_int, [_vp, _fp, C.c_uint32, _fp]),
    "fr_arhosek_cook": (C.c_int, [C.c_float, C.c_float, C.c_float, _fp]),
    "fr_camera_transform": (C.c_int, [_fp, _fp]),
    "fr_camera_walk": (C.c_int, [_fp, C.c_float, C.c_float, C.c_int, C.c_float, _fp]),
}

_lib = None


def lib():
    """Loads libfredholm_b200.so (once).  Raises LibraryNotBuilt if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryNotBuilt(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise FredholmError(lib().fr_last_error().decode("utf-8", "replace"))


def _f(a):
    return a.ctypes.data_as(_fp)


def _u(a):
    return a.ctypes.data_as(_up)


def _f32(x, n=None):
    a = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
    if n is not None:
        assert a.size == n, (a.size, n)
    return a


@dataclass
class Camera:
    """Camera parameter block handed to render(): camera-to-world 3x4 rows + thin lens
    (reference CameraParams, shared.h:59-64).  `from_origin` mirrors fredholm::Camera's
    constructor (camera.h:51-69): look down -z from `origin`."""
    transform: np.ndarray   # (12,) f32 row-major 3x4 camera-to-world
    fov: float = 0.5 * np.pi
    F: float = 8.0
    focus: float = 10000.0

    @staticmethod
    def from_origin(origin, fov=0.5 * np.pi, F=8.0, focus=10000.0):
        out = np.zeros(12, dtype=np.float32)
        _check(lib().fr_camera_transform(_f(_f32(origin, 3)), _f(out)))
        return Camera(out, float(fov), float(F), float(focus))


LAYER_NAMES = ("beauty", "position", "depth", "normal", "texcoord", "albedo")
STAGE_NAMES = ("generate", "trace_closest", "shade", "trace_shadow", "trace_light", "advance", "film")


def pinned_array(shape, dtype=np.float32):
    """numpy array backed by page-locked host memory (never freed: benchmark helper)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = lib().fr_host_alloc_pinned(n)
    if not p:
        raise FredholmError(lib().fr_last_error().decode())
    buf = (C.c_char * n).from_address(p)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def event_elapsed_ms(ev0, ev1, destroy=True):
    ms = C.c_float()
    _check(lib().fr_event_elapsed_ms(ev0, ev1, C.byref(ms)))
    if destroy:
        lib().fr_event_destroy(ev0)
        lib().fr_event_destroy(ev1)
    return ms.value


class DeviceLayers:
    """Caller-owned device AOV buffers (reference: the app owns six CUDABuffers,
    controller.cpp:80-124)."""

    def __init__(self, width, height, names=LAYER_NAMES):
        self.width, self.height = width, height
        self.ptr = {}
        self.names = tuple(names)
        n = width * height
        for name in self.names:
            nbytes = n * (4 if name == "depth" else 16)
            p = lib().fr_device_alloc(nbytes)
            if not p:
                raise FredholmError(lib().fr_last_error().decode())
            self.ptr[name] = p
        self.clear()

    def clear(self):
        n = self.width * self.height
        for name, p in self.ptr.items():
            _check(lib().fr_device_memset(p, 0, n * (4 if name == "depth" else 16)))

    def struct(self):
        s = _Layers()
        for name in LAYER_NAMES:
            setattr(s, name, self.ptr.get(name))
        return s

    def download(self, name):
        n = self.width * self.height
        c = 1 if name == "depth" else 4
        out = np.empty((self.height, self.width, c) if c == 4 else (self.height, self.width), dtype=np.float32)
        _check(lib().fr_copy_to_host(out.ctypes.data_as(_vp), self.ptr[name], n * 4 * c))
        return out

    def free(self):
        for p in self.ptr.values():
            lib().fr_device_free(p)
        self.ptr = {}

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _read_scene(handle, sizes_fn, arrays_fn, texinfo_fn, texdata_fn) -> SceneArrays:
    sz = np.zeros(6, dtype=np.uint32)
    _check(sizes_fn(handle, _u(sz)))
    nv, nf, nm, nt, ns, _ = [int(v) for v in sz]
    v = np.zeros((nv, 3), np.float32)
    n = np.zeros((nv, 3), np.float32)
    t = np.zeros((nv, 2), np.float32)
    idx = np.zeros((nf, 3), np.uint32)
    mid = np.zeros(nf, np.uint32)
    iid = np.zeros(nf, np.uint32)
    mats = np.zeros(nm, MATERIAL_DTYPE)
    so = np.zeros(ns, np.uint32)
    sn = np.zeros(ns, np.uint32)
    tr = np.zeros((ns, 16), np.float32)
    cam = np.zeros(16, np.float32)
    _check(arrays_fn(handle, _f(v), _f(n), _f(t), _u(idx), _u(mid), _u(iid), mats.ctypes.data_as(_vp), _u(so),
                     _u(sn), _f(tr), _f(cam)))
    textures = []
    for i in range(nt):
        w, h, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(texinfo_fn(handle, i, C.byref(w), C.byref(h), C.byref(c)))
        img = np.zeros((h.value, w.value, 4), np.uint8)
        _check(texdata_fn(handle, i, img.ctypes.data_as(_u8p)))
        textures.append((img, bool(c.value)))
    s = SceneArrays(v, n, t, idx, mid, mats, so, sn, iid, tr, textures)
    s.has_camera = bool(sz[5])
    s.camera_transform = cam
    return s


def load_image8(path):
    """(H, W, 4) uint8 as fredholm::Texture holds it: row 0 = bottom row of the file."""
    w, h = C.c_uint32(), C.c_uint32()
    _check(lib().fr_image8_load(os.fsencode(str(path)), C.byref(w), C.byref(h)))
    img = np.zeros((h.value, w.value, 4), np.uint8)
    _check(lib().fr_image8_copy(img.ctypes.data_as(_u8p)))
    return img


def load_imagef(path):
    """(H, W, 4) float32 as fredholm::FloatTexture holds it: row 0 = top row of the file."""
    w, h = C.c_uint32(), C.c_uint32()
    _check(lib().fr_imagef_load(os.fsencode(str(path)), C.byref(w), C.byref(h)))
    img = np.zeros((h.value, w.value, 4), np.float32)
    _check(lib().fr_imagef_copy(img.ctypes.data_as(_fp)))
    return img


def write_png(path, pixels):
    """pixels: (H, W, 3|4) uint8, row 0 = top row."""
    a = np.ascontiguousarray(pixels, dtype=np.uint8)
    assert a.ndim == 3 and a.shape[2] in (3, 4)
    _check(lib().fr_write_png(os.fsencode(str(path)), a.ctypes.data_as(_u8p), a.shape[1], a.shape[0], a.shape[2]))


def _scene_args(s):
    """Argument tuple of fr_set_scene_arrays / fr_scene_set_arrays.  The C ABI sees pointers and ONE vertex /
    face / sub-mesh count, so the per-array lengths are checked here."""
    def arr(x, dtype, shape, what):
        a = np.ascontiguousarray(x, dtype=dtype)
        if a.ndim != len(shape) or any(d is not None and a.shape[i] != d for i, d in enumerate(shape)):
            raise FredholmError("invalid scene: %s has shape %s" % (what, a.shape))
        return a
    v = arr(s.vertices, np.float32, (None, 3), "vertices")
    n = arr(s.normals, np.float32, (len(v), 3), "normals")
    t = arr(s.texcoords, np.float32, (len(v), 2), "texcoords")
    idx = arr(s.indices, np.uint32, (None, 3), "indices")
    mid = arr(s.material_ids, np.uint32, (len(idx),), "material_ids")
    iid = arr(s.instance_ids, np.uint32, (len(idx),), "instance_ids")
    so = arr(s.submesh_offsets, np.uint32, (None,), "submesh_offsets")
    sn = arr(s.submesh_n_faces, np.uint32, (len(so),), "submesh_n_faces")
    tr = arr(np.asarray(s.transforms, dtype=np.float32).reshape(-1, 16), np.float32, (len(so), 16), "transforms")
    mats = np.ascontiguousarray(s.materials)
    if mats.dtype.itemsize != 180:
        raise FredholmError("invalid scene: materials must be 180-byte records")
    keep = (v, n, t, idx, mid, iid, so, sn, tr, mats)
    args = (_f(v), _f(n), _f(t), len(v), _u(idx), _u(mid), _u(iid), len(idx), mats.ctypes.data_as(_vp), len(mats),
            _u(so), _u(sn), _f(tr), len(so))
    _scene_args.keepalive = keep  # the arrays must outlive the call that follows
    return args


class Scene:
    """Host-side fredholm::Scene (file loaders, animation); needs no GPU."""

    def __init__(self):
        self._h = lib().fr_scene_create()

    def load_model(self, path, clear=True):
        _check(lib().fr_scene_load(self._h, os.fsencode(str(path)), 1 if clear else 0))

    def arrays(self) -> SceneArrays:
        L = lib()
        return _read_scene(self._h, L.fr_scene_get_sizes, L.fr_scene_get_arrays, L.fr_scene_get_texture_info,
                           L.fr_scene_get_texture_data)

    def update_animation(self, t):
        _check(lib().fr_scene_update_animation(self._h, float(t)))

    def set_arrays(self, s: "SceneArrays"):
        _check(lib().fr_scene_set_arrays(self._h, *_scene_args(s)))

    def validate(self):
        """Scene::validate: raises FredholmError("invalid scene: ...") on any out-of-range index."""
        _check(lib().fr_scene_validate(self._h))

    def close(self):
        if getattr(self, "_h", None):
            lib().fr_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Renderer:
    def __init__(self, device=0):
        L = lib()
        if L.fr_device_count() <= device:
            raise FredholmError("no CUDA device %d (fredholm_b200 has no CPU fallback)" % device)
        self._h = L.fr_renderer_create(device)
        if not self._h:
            raise FredholmError(L.fr_last_error().decode())
        self._borrowed = False
        self.width = self.height = 0

    @classmethod
    def _view(cls, handle):
        """Wraps a handle owned by someone else (a rank of a MultiRenderer); close() leaves it alone."""
        r = cls.__new__(cls)
        r._h = handle
        r._borrowed = True
        r.width = r.height = 0
        return r

    def close(self):
        if getattr(self, "_h", None):
            if not getattr(self, "_borrowed", False):
                lib().fr_renderer_destroy(self._h)
            self._h = None

    # ---- multi-GPU: this renderer as one rank of a world (include/fredholm/multi_gpu.h) ----
    def comm_init(self, comm_id, rank, world):
        """Collective: ncclCommInitRank inside the C++ core.  comm_id: the 128 bytes of comm_unique_id() made on
        rank 0 and shipped to every rank."""
        ident = np.frombuffer(bytes(comm_id), dtype=np.uint8).copy()
        assert ident.size == 128
        _check(lib().fr_comm_init(self._h, ident.ctypes.data_as(_u8p), int(rank), int(world)))

    def comm_destroy(self):
        _check(lib().fr_comm_destroy(self._h))

    def _layers_struct(self, layers):
        if isinstance(layers, DeviceLayers):
            return layers.struct()
        st = _Layers()
        for name in LAYER_NAMES:
            setattr(st, name, layers.get(name))
        return st

    def render_sharded(self, camera, bg_color, layers, total_spp, max_depth, root=0):
        """Collective: this rank's sample slice of a total_spp frame (sums into the ZEROED layers), one ncclReduce of
        the bound layers onto root, division by total_spp there.  Asynchronous on the renderer's stream."""
        st = self._layers_struct(layers)
        _check(lib().fr_render_sharded(self._h, _f(_f32(camera.transform, 12)), camera.fov, camera.F, camera.focus,
                                       _f(_f32(bg_color, 3)), C.byref(st), int(total_spp), int(max_depth), int(root)))

    def reduce_layers(self, layers, total_spp, root=0):
        st = self._layers_struct(layers)
        _check(lib().fr_reduce_layers(self._h, C.byref(st), int(total_spp), int(root)))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- scene ----
    def load_scene(self, path, clear=True):
        _check(lib().fr_load_scene(self._h, os.fsencode(str(path)), 1 if clear else 0))

    def set_scene(self, s: SceneArrays):
        L = lib()
        for rgba8, is_color in s.textures:
            img = np.ascontiguousarray(rgba8, dtype=np.uint8)
            assert img.ndim == 3 and img.shape[2] == 4
            if L.fr_stage_texture(self._h, img.ctypes.data_as(_u8p), img.shape[1], img.shape[0],
                                  1 if is_color else 0) < 0:
                raise FredholmError(L.fr_last_error().decode())
        _check(L.fr_set_scene_arrays(self._h, *_scene_args(s)))

    def get_scene(self) -> SceneArrays:
        L = lib()
        return _read_scene(self._h, L.fr_get_scene_sizes, L.fr_get_scene_arrays, L.fr_get_texture_info,
                           L.fr_get_texture_data)

    def set_scene_object(self, scene: "Scene"):
        _check(lib().fr_set_scene(self._h, scene._h))

    def build_accel(self):
        _check(lib().fr_build_accel(self._h))

    def accel_info(self):
        out = np.zeros(3, np.uint32)
        ms = C.c_float()
        nbytes = C.c_uint64()
        _check(lib().fr_get_accel_info(self._h, _u(out), C.byref(ms), C.byref(nbytes)))
        out5 = np.zeros(5, np.uint32)
        tlas_ms = C.c_float()
        _check(lib().fr_get_accel_info2(self._h, _u(out5), C.byref(tlas_ms)))
        return dict(n_faces=int(out[0]), n_nodes=int(out[1]), depth=int(out[2]), build_ms=ms.value,
                    bytes=int(nbytes.value), two_level=bool(out5[0] & 1), tlas_refitted=bool(out5[0] & 2),
                    n_instances=int(out5[1]), n_meshes=int(out5[2]), n_stored_faces=int(out5[3]),
                    tlas_update_ms=tlas_ms.value)

    ACCEL_MODES = {"auto": 0, "flat": 1, "two_level": 2}

    def set_accel_mode(self, mode):
        """"auto" | "flat" | "two_level" (include/fredholm/renderer.h AccelMode); applies at the next build_accel."""
        _check(lib().fr_set_accel_mode(self._h, self.ACCEL_MODES[mode]))

    NODE_DTYPE = np.dtype([("p", "<f4", 3), ("e", "u1", 3), ("imask", "u1"), ("child_base", "<u4"),
                           ("tri_base", "<u4"), ("meta", "u1", 8), ("qlo", "u1", (3, 8)), ("qhi", "u1", (3, 8))])

    def accel_data(self):
        """(nodes, tris): CWBVH nodes as NODE_DTYPE records (80 B) and leaf triangles (n,3,4) f32
        (xyz + face id / flags bits in w)."""
        info = self.accel_info()
        nodes = np.zeros(info["n_nodes"], self.NODE_DTYPE)
        n_tris = info["n_instances"] + info["n_stored_faces"] if info["two_level"] else info["n_faces"]
        tris = np.zeros((n_tris, 3, 4), np.float32)
        _check(lib().fr_get_accel_data(self._h, nodes.ctypes.data_as(_vp), tris.ctypes.data_as(_vp)))
        return nodes, tris

    def set_time(self, t):
        _check(lib().fr_set_time(self._h, float(t)))

    def set_transforms(self, transforms):
        tr = _f32(transforms).reshape(-1, 16)
        _check(lib().fr_set_transforms(self._h, _f(tr), len(tr)))

    # ---- lights / sky ----
    def set_directional_light(self, le, direction, angle):
        _check(lib().fr_set_directional_light(self._h, _f(_f32(le, 3)), _f(_f32(direction, 3)), float(angle)))

    def clear_directional_light(self):
        _check(lib().fr_clear_directional_light(self._h))

    def set_sky_intensity(self, v):
        _check(lib().fr_set_sky_intensity(self._h, float(v)))

    def load_arhosek_sky(self, turbidity, albedo):
        _check(lib().fr_load_arhosek_sky(self._h, float(turbidity), float(albedo)))

    def clear_arhosek_sky(self):
        _check(lib().fr_clear_arhosek_sky(self._h))

    def set_ibl(self, rgba32f):
        img = np.ascontiguousarray(rgba32f, dtype=np.float32)
        assert img.ndim == 3 and img.shape[2] == 4
        _check(lib().fr_set_ibl(self._h, _f(img), img.shape[1], img.shape[0]))

    def load_ibl(self, path):
        """Renderer::load_ibl (renderer.h:574-583): Radiance .hdr file, not flipped."""
        _check(lib().fr_load_ibl(self._h, os.fsencode(str(path))))

    def clear_ibl(self):
        _check(lib().fr_clear_ibl(self._h))

    # ---- film ----
    def set_resolution(self, width, height):
        self.width, self.height = int(width), int(height)
        _check(lib().fr_set_resolution(self._h, self.width, self.height))

    def init_render_states(self):
        _check(lib().fr_init_render_states(self._h))

    def set_sample_offset(self, first):
        _check(lib().fr_set_sample_offset(self._h, int(first)))

    def sample_count(self):
        return int(lib().fr_get_sample_count(self._h))

    def set_film_mode(self, mode):
        _check(lib().fr_set_film_mode(self._h, {"mean": 0, "sum": 1}.get(mode, mode)))

    def set_max_wave_paths(self, n):
        _check(lib().fr_set_max_wave_paths(self._h, int(n)))

    def set_single_launch(self, on):
        """render(n_samples) as ONE reference launch (payload.firsthit outlives the sample loop, pt.cu:432-433)."""
        _check(lib().fr_set_single_launch(self._h, 1 if on else 0))

    # ---- render ----
    def render(self, camera: Camera, bg_color, layers, n_samples, max_depth):
        """layers: DeviceLayers, or a dict name -> device pointer (e.g. torch tensor .data_ptr())."""
        if isinstance(layers, DeviceLayers):
            st = layers.struct()
        else:
            st = _Layers()
            for name in LAYER_NAMES:
                setattr(st, name, layers.get(name))
        _check(lib().fr_render(self._h, _f(_f32(camera.transform, 12)), camera.fov, camera.F, camera.focus,
                               _f(_f32(bg_color, 3)), C.byref(st), int(n_samples), int(max_depth)))

    def wait(self):
        _check(lib().fr_wait(self._h))

    def render_frame_host(self, camera: Camera, bg_color, n_samples, max_depth, names=("beauty",), out=None):
        """One frame through host buffers (clear, render, read back).  Returns dict of arrays."""
        n = self.width * self.height
        res = out if out is not None else {}
        st = _Layers()
        for name in names:
            if name not in res:
                res[name] = np.empty((self.height, self.width, 4) if name != "depth" else (self.height, self.width),
                                     dtype=np.float32)
            assert res[name].size == n * (1 if name == "depth" else 4)
            setattr(st, name, res[name].ctypes.data_as(_vp))
        _check(lib().fr_render_frame_host(self._h, _f(_f32(camera.transform, 12)), camera.fov, camera.F,
                                          camera.focus, _f(_f32(bg_color, 3)), C.byref(st), int(n_samples),
                                          int(max_depth)))
        return res

    def batch_run(self, camera, width, height, n_spp, max_depth, n_frames, camera_path=None, output_dir=None,
                  keep_frames=True, denoise=True, upscale=False, use_bloom=True, bloom_threshold=2.0,
                  bloom_sigma=5.0, ISO=80.0, chromatic_aberration=1.0, fps=24.0, start_time=0.0, max_time=9.5,
                  kill_time_s=590.0, first_frame=0, frame_stride=1, bg_color=(0, 0, 0), animate=True,
                  n_save_threads=2, n_slots=3, dn_iterations=0, dn_sigma_color=0.0, dn_sigma_albedo=0.0,
                  dn_albedo_floor=0.0, dn_firefly_k=-1.0):
        """fredholm::FrameBatch::run (app/rtcamp8.cpp render + save loop).  Returns
        (records: list of dict, frames: (n, H, W, 4) uint8 or None, info dict)."""
        cfg = _BatchConfig()
        cfg.width, cfg.height, cfg.n_spp, cfg.max_depth = int(width), int(height), int(n_spp), int(max_depth)
        cfg.post = _PostProcessParams(1 if use_bloom else 0, bloom_threshold, bloom_sigma, ISO, chromatic_aberration)
        cfg.denoise, cfg.upscale = (1 if denoise else 0), (1 if upscale else 0)
        cfg.dn_iterations, cfg.dn_sigma_color = int(dn_iterations), float(dn_sigma_color)
        cfg.dn_sigma_albedo, cfg.dn_albedo_floor = float(dn_sigma_albedo), float(dn_albedo_floor)
        cfg.dn_firefly_k = float(dn_firefly_k)
        cfg.fps, cfg.start_time, cfg.max_time, cfg.kill_time_s = fps, start_time, max_time, kill_time_s
        cfg.first_frame, cfg.frame_stride, cfg.max_frames = int(first_frame), int(frame_stride), int(n_frames)
        cfg.bg_color = (C.c_float * 3)(*[float(v) for v in bg_color])
        cfg.animate = 1 if animate else 0
        cfg.output_dir = str(output_dir).encode() if output_dir else None
        cfg.n_save_threads, cfg.n_slots = int(n_save_threads), int(n_slots)
        if camera_path is not None:
            path = np.ascontiguousarray(camera_path, dtype=np.float32).reshape(-1, 12)
            n_path = len(path)
        else:
            path = _f32(camera.transform, 12).reshape(1, 12)
            n_path = 0
        ow, oh = (2 * width, 2 * height) if upscale else (width, height)
        recs = (_FrameRecord * int(n_frames))()
        frames = np.zeros((int(n_frames), oh, ow, 4), np.uint8) if keep_frames else None
        out5 = np.zeros(5, np.uint32)
        wall = C.c_double(0.0)
        _check(lib().fr_batch_run(self._h, C.byref(cfg), _f(path), n_path, camera.fov, camera.F, camera.focus, recs,
                                  int(n_frames), frames.ctypes.data_as(_u8p) if keep_frames else None, _u(out5),
                                  C.byref(wall)))
        n = int(out5[0])
        records = [{name: getattr(recs[i], name) for name, _ in _FrameRecord._fields_} for i in range(n)]
        info = dict(n_frames=n, out_width=int(out5[1]), out_height=int(out5[2]), killed=bool(out5[3]),
                    wall_s=wall.value)
        return records, (frames[:n] if keep_frames else None), info

    def scale_layers(self, layers, scale):
        st = layers.struct() if isinstance(layers, DeviceLayers) else None
        if st is None:
            st = _Layers()
            for name in LAYER_NAMES:
                setattr(st, name, layers.get(name))
        _check(lib().fr_scale_layers(self._h, C.byref(st), float(scale)))

    def statistics(self):
        out = np.zeros(6, np.uint64)
        _check(lib().fr_get_statistics(self._h, out.ctypes.data_as(_u64p)))
        d = dict(paths=int(out[0]), rays_radiance=int(out[1]), rays_shadow=int(out[2]), rays_light=int(out[3]),
                 kernel_launches=int(out[4]), rays_skipped=int(out[5]))
        d["rays"] = d["rays_radiance"] + d["rays_shadow"] + d["rays_light"]
        return d

    def reset_statistics(self):
        _check(lib().fr_reset_statistics(self._h))

    def set_traversal_counting(self, on=True):
        """Measurement: counting instantiations of the traversal kernels (process-wide)."""
        _check(lib().fr_set_traversal_counting(self._h, 1 if on else 0))

    def traversal_counters(self):
        """{ray type: (nodes visited, triangles tested)} since the last reset (counting mode only)."""
        out = np.zeros(6, np.uint64)
        _check(lib().fr_get_traversal_counters(self._h, out.ctypes.data_as(_u64p)))
        return {k: (int(out[i]), int(out[3 + i])) for i, k in enumerate(("radiance", "shadow", "light"))}

    def set_wave_overlap(self, on=True):
        _check(lib().fr_set_wave_overlap(self._h, 1 if on else 0))

    def set_wave_compaction(self, on=True, depth=0):
        """Stragglers of several waves finish their late bounces together (default on; depth 0 = after 3 bounces)."""
        _check(lib().fr_set_wave_compaction(self._h, 1 if on else 0, int(depth)))

    def wave_state_bytes(self):
        return int(lib().fr_get_wave_state_bytes(self._h))

    def set_samples_per_warp(self, spw):
        _check(lib().fr_set_samples_per_warp(self._h, int(spw)))

    def stream(self):
        return int(lib().fr_get_stream(self._h))

    def set_stage_timing(self, on=True):
        _check(lib().fr_set_stage_timing(self._h, 1 if on else 0))

    def stage_times(self):
        """{stage: (total ms, launches)} since the last call (device time, CUDA events)."""
        ms = (C.c_double * 7)()
        n = (C.c_uint64 * 7)()
        _check(lib().fr_get_stage_times(self._h, ms, n))
        return {name: (ms[i], int(n[i])) for i, name in enumerate(STAGE_NAMES)}

    def record_event(self):
        ev = lib().fr_event_create()
        if not ev:
            raise FredholmError(lib().fr_last_error().decode())
        _check(lib().fr_event_record(self._h, ev))
        return ev

    # ---- stage-level queries (parity tests) ----
    def trace_closest(self, rays, tmin=0.0, tmax=1e9, counters=False):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
        n = len(rays)
        ids = np.zeros((n, 2), np.uint32)
        tuv = np.zeros((n, 3), np.float32)
        cnt = np.zeros(2, np.uint64)
        _check(lib().fr_trace_closest(self._h, _f(rays), n, float(tmin), float(tmax), _u(ids), _f(tuv),
                                      cnt.ctypes.data_as(_u64p) if counters else None))
        return (ids, tuv, cnt) if counters else (ids, tuv)

    def primary_rays(self, camera: Camera, n_spp=0):
        out = np.zeros((self.height, self.width, 6), np.float32)
        _check(lib().fr_primary_rays(self._h, _f(_f32(camera.transform, 12)), camera.fov, camera.F, camera.focus,
                                     int(n_spp), _f(out)))
        return out

    def sky_radiance(self, dirs):
        d = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3)
        out = np.zeros_like(d)
        _check(lib().fr_sky_radiance(self._h, _f(d), len(d), _f(out)))
        return out


def set_device(device):
    """cudaSetDevice for this thread (fr_device_alloc / DeviceLayers allocate on the current device)."""
    _check(lib().fr_set_device(int(device)))


def device_attributes(device=0):
    out = np.zeros(4, np.uint32)
    mem = np.zeros(1, np.uint64)
    _check(lib().fr_get_device_attributes(int(device), _u(out), mem.ctypes.data_as(_u64p)))
    return dict(sm_count=int(out[0]), clock_khz=int(out[1]), cc=int(out[2]), l2_bytes=int(out[3]), total_mem=int(mem[0]))


def comm_unique_id():
    """ncclGetUniqueId through the C ABI: 128 bytes, made on rank 0."""
    out = np.zeros(128, np.uint8)
    _check(lib().fr_comm_get_unique_id(out.ctypes.data_as(_u8p)))
    return out.tobytes()


def sample_slice(total_spp, rank, world):
    """(first, count) of rank's sample slice -- the C++ core's rule (whole 16-sample CMJ patterns)."""
    first, count = C.c_uint32(), C.c_uint32()
    _check(lib().fr_sample_slice(int(total_spp), int(rank), int(world), C.byref(first), C.byref(count)))
    return first.value, count.value


class MultiRenderer:
    """fredholm::MultiGpuRenderer: one process, one renderer + host thread per device, NCCL inside the core."""

    def __init__(self, devices=None):
        L = lib()
        if L.fr_device_count() <= 0:
            raise FredholmError("no CUDA device (fredholm_b200 has no CPU fallback)")
        if devices:
            arr = (C.c_int * len(devices))(*devices)
            self._h = L.fr_multi_create(arr, len(devices))
        else:
            self._h = L.fr_multi_create(None, 0)
        if not self._h:
            raise FredholmError(L.fr_last_error().decode())
        self.ranks = [Renderer._view(L.fr_multi_renderer(self._h, i)) for i in range(L.fr_multi_size(self._h))]

    def __len__(self):
        return len(self.ranks)

    def for_each(self, f):
        for r in self.ranks:
            f(r)

    def render(self, camera, bg_color, layers_rank0, total_spp, max_depth):
        st = self.ranks[0]._layers_struct(layers_rank0)
        _check(lib().fr_multi_render(self._h, _f(_f32(camera.transform, 12)), camera.fov, camera.F, camera.focus,
                                     _f(_f32(bg_color, 3)), C.byref(st), int(total_spp), int(max_depth)))

    def wait(self):
        _check(lib().fr_multi_wait(self._h))

    def close(self):
        if getattr(self, "_h", None):
            for r in self.ranks:
                r.close()
            lib().fr_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sampler_sequence(width, height, seed, image_idx, n_spp, kinds):
    n_out = sum(1 if k == "1" else 2 for k in kinds)
    out = np.zeros(n_out, np.float32)
    _check(lib().fr_sampler_sequence(width, height, seed, image_idx, n_spp, kinds.encode(), _f(out)))
    return out


def bsdf_eval_sample(cases):
    """cases: (n,40) f32 -> (n,11): eval f[3], pdf, sample wi[3], f[3], pdf."""
    c = np.ascontiguousarray(cases, dtype=np.float32).reshape(-1, 40)
    out = np.zeros((len(c), 11), np.float32)
    _check(lib().fr_bsdf_eval_sample(_f(c), len(c), _f(out)))
    return out


def arhosek_cook(turbidity, albedo, elevation):
    out = np.zeros(30, np.float32)
    _check(lib().fr_arhosek_cook(float(turbidity), float(albedo), float(elevation), _f(out)))
    return out


def camera_walk(origin, d_phi, d_theta, movement, dt):
    out = np.zeros(12, np.float32)
    _check(lib().fr_camera_walk(_f(_f32(origin, 3)), float(d_phi), float(d_theta), int(movement), float(dt), _f(out)))
    return out


def post_process(beauty_in, high, temp, width, height, out, use_bloom=True, bloom_threshold=2.0, bloom_sigma=5.0,
                 ISO=80.0, chromatic_aberration=1.0):
    p = _PostProcessParams(1 if use_bloom else 0, bloom_threshold, bloom_sigma, ISO, chromatic_aberration)
    _check(lib().fr_post_process(beauty_in, high, temp, width, height, C.byref(p), out))


def denoise(beauty, normal, albedo, out, width, height, upscale=False, iterations=0, sigma_color=0.0,
            sigma_albedo=0.0, albedo_floor=0.0, firefly_k=-1.0):
    """fredholm::Denoiser one-shot on device pointers (defaults when a parameter is 0; firefly_k < 0)."""
    _check(lib().fr_denoise(beauty, normal, albedo, out, int(width), int(height), 1 if upscale else 0,
                            int(iterations), float(sigma_color), float(sigma_albedo), float(albedo_floor),
                            float(firefly_k)))


def tone_mapping(beauty_in, width, height, out, ISO=80.0, chromatic_aberration=1.0):
    _check(lib().fr_tone_mapping(beauty_in, width, height, ISO, chromatic_aberration, out))
