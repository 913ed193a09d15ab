"""ctypes binding of oracle/_ref/libfredholm_oracle.so -- TEST INFRASTRUCTURE ONLY.

The library is the reference's own integrator (pt.cu and everything it includes)
compiled for the host plus the oracle's OptiX shim (oracle_host.cpp); see the
header of oracle_host.cpp.  Built by oracle/Makefile (needs /root/reference); the
prebuilt .so travels to the GPU box.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libfredholm_oracle.so")
REFERENCE_ROOT = "/root/reference"

_fp = C.POINTER(C.c_float)
_up = C.POINTER(C.c_uint32)
_vp = C.c_void_p


def build(force=False):
    """Compiles the oracle from the reference sources (only possible where
    /root/reference exists).  Returns True if the library is available afterwards."""
    if os.path.isdir(REFERENCE_ROOT) and (force or not os.path.exists(LIB_PATH)):
        subprocess.run(["make", "-C", _HERE, "-j8"], check=True, stdout=subprocess.DEVNULL)
    return os.path.exists(LIB_PATH)


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("oracle library missing: run `make -C oracle` where /root/reference exists")
        L = C.CDLL(LIB_PATH)
        L.orc_render.restype = C.c_double
        if hasattr(L, "orc_render_tiles"):
            L.orc_render_tiles.restype = C.c_double
        L.orc_n_lights.restype = C.c_uint32
        L.orc_last_error.restype = C.c_char_p
        for n in ("orc_xxhash32_1", "orc_xxhash32_4", "orc_cmj_permute", "orc_sobol", "orc_owen",
                  "orc_sizeof_material", "orc_sizeof_shading_params", "orc_sizeof_launch_params"):
            getattr(L, n).restype = C.c_uint32
        L.orc_sobol.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
        _lib = L
    return _lib


def _f(a):
    return a.ctypes.data_as(_fp)


def _u(a):
    return a.ctypes.data_as(_up)


def _f32(x, n=None):
    a = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
    if n is not None:
        assert a.size == n
    return a


LAYER_NAMES = ("beauty", "position", "depth", "normal", "texcoord", "albedo")


class Oracle:
    """Mirrors the Renderer call sequence against the host oracle (one global scene)."""

    def __init__(self):
        self.L = lib()
        self.L.orc_reset()
        self.width = self.height = 0

    # ---- scene ----
    def set_scene(self, s):
        self.L.orc_reset()
        self.L.orc_set_scene(_f(s.vertices), _f(s.normals), _f(s.texcoords), C.c_uint32(len(s.vertices)),
                             _u(s.indices), _u(s.material_ids), _u(s.instance_ids), C.c_uint32(len(s.indices)),
                             s.materials.ctypes.data_as(_vp), C.c_uint32(len(s.materials)), _u(s.submesh_offsets),
                             _u(s.submesh_n_faces), _f(s.transforms), C.c_uint32(len(s.submesh_offsets)))
        for rgba8, is_color in s.textures:
            img = np.ascontiguousarray(rgba8, dtype=np.uint8)
            self.L.orc_add_texture(img.ctypes.data_as(_vp), C.c_int(img.shape[1]), C.c_int(img.shape[0]),
                                   C.c_int(1 if is_color else 0))

    def load_scene(self, path, clear=True):
        rc = self.L.orc_load_scene(os.fsencode(str(path)), C.c_int(1 if clear else 0))
        if rc != 0:
            raise RuntimeError(self.L.orc_last_error().decode())

    def get_loaded_scene(self):
        """Flat arrays produced by the reference's own loader (after load_scene)."""
        from fredholm_b200.types import MATERIAL_DTYPE, SceneArrays
        sz = np.zeros(6, np.uint32)
        self.L.orc_scene_sizes(_u(sz))
        nv, nf, nm, nt, ns, has_cam = [int(v) for v in sz]
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
        self.L.orc_scene_copy(_f(v), _f(n), _f(t), _u(idx), _u(mid), _u(iid), mats.ctypes.data_as(_vp), _u(so),
                              _u(sn), _f(tr), _f(cam))
        textures = []
        for i in range(nt):
            w, h, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
            self.L.orc_scene_texture_info(C.c_uint32(i), C.byref(w), C.byref(h), C.byref(c))
            img = np.zeros((h.value, w.value, 4), np.uint8)
            self.L.orc_scene_texture_copy(C.c_uint32(i), img.ctypes.data_as(_vp))
            textures.append((img, bool(c.value)))
        s = SceneArrays(v, n, t, idx, mid, mats, so, sn, iid, tr, textures)
        s.has_camera = bool(has_cam)
        s.camera_transform = cam
        return s

    def set_time(self, t):
        self.L.orc_set_time(C.c_float(t))

    def set_transforms(self, transforms):
        tr = _f32(transforms).reshape(-1, 16)
        self.L.orc_set_transforms(_f(tr), C.c_uint32(len(tr)))

    def build_accel(self):
        self.L.orc_build_accel()

    def n_lights(self):
        return int(self.L.orc_n_lights())

    # ---- lights / sky ----
    def set_directional_light(self, le, direction, angle):
        self.L.orc_set_directional_light(_f(_f32(le, 3)), _f(_f32(direction, 3)), C.c_float(angle))

    def clear_directional_light(self):
        self.L.orc_clear_directional_light()

    def set_sky_intensity(self, v):
        self.L.orc_set_sky_intensity(C.c_float(v))

    def load_arhosek_sky(self, turbidity, albedo):
        self.L.orc_load_arhosek_sky(C.c_float(turbidity), C.c_float(albedo))

    def clear_arhosek_sky(self):
        self.L.orc_clear_arhosek_sky()

    def set_ibl(self, rgba32f):
        img = np.ascontiguousarray(rgba32f, dtype=np.float32)
        self.L.orc_set_ibl(_f(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]))

    # ---- film ----
    def set_resolution(self, w, h):
        self.width, self.height = int(w), int(h)
        self.L.orc_set_resolution(C.c_uint32(w), C.c_uint32(h))

    def init_render_states(self):
        self.L.orc_init_render_states()

    def set_sample_count(self, v):
        self.L.orc_set_sample_count(C.c_uint32(v))

    def new_layers(self):
        h, w = self.height, self.width
        return {n: np.zeros((h, w, 4) if n != "depth" else (h, w), np.float32) for n in LAYER_NAMES}

    def render(self, camera, bg_color, layers, n_samples, max_depth, window=None, n_threads=1):
        """One launch of the reference raygen over `window` = (x0, y0, x1, y1).
        Returns seconds spent in the launch loop."""
        x0, y0, x1, y1 = window if window is not None else (0, 0, self.width, self.height)
        return float(self.L.orc_render(
            _f(_f32(camera.transform, 12)), C.c_float(camera.fov), C.c_float(camera.F), C.c_float(camera.focus),
            _f(_f32(bg_color, 3)), _f(layers["beauty"]), _f(layers["position"]), _f(layers["depth"]),
            _f(layers["normal"]), _f(layers["texcoord"]), _f(layers["albedo"]), C.c_uint32(n_samples),
            C.c_uint32(max_depth), C.c_uint32(x0), C.c_uint32(y0), C.c_uint32(x1), C.c_uint32(y1),
            C.c_int(n_threads)))

    def render_tiles(self, camera, bg_color, layers, n_samples, max_depth, tiles, n_threads=1):
        """One launch over a list of windows [(x0, y0, x1, y1), ...]; returns seconds in the launch loop."""
        t = np.ascontiguousarray(tiles, dtype=np.uint32).reshape(-1, 4)
        return float(self.L.orc_render_tiles(
            _f(_f32(camera.transform, 12)), C.c_float(camera.fov), C.c_float(camera.F), C.c_float(camera.focus),
            _f(_f32(bg_color, 3)), _f(layers["beauty"]), _f(layers["position"]), _f(layers["depth"]),
            _f(layers["normal"]), _f(layers["texcoord"]), _f(layers["albedo"]), C.c_uint32(n_samples),
            C.c_uint32(max_depth), _u(t), C.c_uint32(len(t)), C.c_int(n_threads)))

    def render_canonical(self, camera, bg_color, spp, max_depth, window=None, n_threads=1, layers=None, tiles=None):
        """`spp` launches of one sample each (what the reference GUI does,
        controller.cpp:221-224).  Returns (layers, seconds)."""
        layers = layers if layers is not None else self.new_layers()
        secs = 0.0
        for _ in range(spp):
            if tiles is not None:
                secs += self.render_tiles(camera, bg_color, layers, 1, max_depth, tiles, n_threads)
            else:
                secs += self.render(camera, bg_color, layers, 1, max_depth, window, n_threads)
        return layers, secs

    def ray_counts(self):
        out = np.zeros(3, np.uint64)
        self.L.orc_get_ray_counts(out.ctypes.data_as(_vp))
        return dict(rays_radiance=int(out[0]), rays_shadow=int(out[1]), rays_light=int(out[2]),
                    rays=int(out.sum()))

    def reset_ray_counts(self):
        self.L.orc_reset_ray_counts()

    # ---- stage-level queries ----
    def trace_closest(self, rays, tmin=0.0, tmax=1e9, bruteforce=False):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
        o = np.ascontiguousarray(rays[:, :3])
        d = np.ascontiguousarray(rays[:, 3:])
        ids = np.zeros((len(rays), 2), np.uint32)
        tuv = np.zeros((len(rays), 3), np.float32)
        fn = self.L.orc_trace_closest_bruteforce if bruteforce else self.L.orc_trace_closest
        fn(_f(o), _f(d), C.c_uint32(len(rays)), C.c_float(tmin), C.c_float(tmax), _u(ids), _f(tuv))
        return ids, tuv

    def primary_rays(self, camera, n_spp=0):
        out = np.zeros((self.height, self.width, 6), np.float32)
        self.L.orc_primary_rays(_f(_f32(camera.transform, 12)), C.c_float(camera.fov), C.c_float(camera.F),
                                C.c_float(camera.focus), C.c_uint32(n_spp), _f(out))
        return out

    def sky_radiance(self, dirs):
        d = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3)
        out = np.zeros_like(d)
        for i in range(len(d)):
            self.L.orc_sky_radiance(_f(d[i]), _f(out[i]))
        return out


def load_image8(path):
    """fredholm::Texture(path) of the reference (stb_image, flipped): (H, W, 4) uint8."""
    L = lib()
    w, h = C.c_uint32(), C.c_uint32()
    if L.orc_image8_load(os.fsencode(str(path)), C.byref(w), C.byref(h)) != 0:
        raise RuntimeError(L.orc_last_error().decode())
    img = np.zeros((h.value, w.value, 4), np.uint8)
    L.orc_image8_copy(img.ctypes.data_as(_vp))
    return img


def load_imagef(path):
    """fredholm::FloatTexture(path) of the reference (stbi_loadf, not flipped): (H, W, 4) float32."""
    L = lib()
    w, h = C.c_uint32(), C.c_uint32()
    if L.orc_imagef_load(os.fsencode(str(path)), C.byref(w), C.byref(h)) != 0:
        raise RuntimeError(L.orc_last_error().decode())
    img = np.zeros((h.value, w.value, 4), np.float32)
    L.orc_imagef_copy(_f(img))
    return img


def sampler_sequence(width, height, seed, image_idx, n_spp, kinds):
    n_out = sum(1 if k == "1" else 2 for k in kinds)
    out = np.zeros(n_out, np.float32)
    lib().orc_sampler_sequence(C.c_uint32(width), C.c_uint32(height), C.c_uint32(seed), C.c_uint32(image_idx),
                               C.c_uint32(n_spp), kinds.encode(), _f(out))
    return out


def bsdf_eval_sample(cases):
    """cases (n,40) as in fredholm_b200.api.bsdf_eval_sample -> (n,11)."""
    c = np.ascontiguousarray(cases, dtype=np.float32).reshape(-1, 40)
    out = np.zeros((len(c), 11), np.float32)
    L = lib()
    for i in range(len(c)):
        sp = np.ascontiguousarray(c[i, :30])
        wo = np.ascontiguousarray(c[i, 30:33])
        wi = np.ascontiguousarray(c[i, 34:37])
        v2 = np.ascontiguousarray(c[i, 38:40])
        ev = np.zeros(4, np.float32)
        sm = np.zeros(7, np.float32)
        L.orc_bsdf_eval(sp.ctypes.data_as(_vp), _f(wo), C.c_int(int(c[i, 33] != 0)), _f(wi), _f(ev))
        L.orc_bsdf_sample(sp.ctypes.data_as(_vp), _f(wo), C.c_int(int(c[i, 33] != 0)), C.c_float(c[i, 37]),
                          _f(v2), _f(sm))
        out[i, :4] = ev
        out[i, 4:] = sm
    return out


def arhosek_cook(turbidity, albedo, elevation):
    out = np.zeros(30, np.float32)
    lib().orc_arhosek_cook(C.c_float(turbidity), C.c_float(albedo), C.c_float(elevation), _f(out))
    return out


def camera_transform(origin):
    out = np.zeros(12, np.float32)
    lib().orc_camera_transform(_f(_f32(origin, 3)), _f(out))
    return out


def camera_walk(origin, d_phi, d_theta, movement, dt):
    out = np.zeros(12, np.float32)
    lib().orc_camera_walk(_f(_f32(origin, 3)), C.c_float(d_phi), C.c_float(d_theta), C.c_int(movement),
                          C.c_float(dt), _f(out))
    return out


# ---- the reference's own post-process kernels compiled with nvcc (GPU box only) ----------
PP_LIB_PATH = os.path.join(_HERE, "_ref", "libpostprocess_ref.so")
_pp = None


def post_process_ref_available():
    return os.path.exists(PP_LIB_PATH)


def post_process_ref_lib():
    global _pp
    if _pp is None:
        _pp = C.CDLL(PP_LIB_PATH)
        _pp.ppr_last_error.restype = C.c_char_p
        _pp.ppr_post_process.argtypes = [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                         C.c_float, _vp]
        _pp.ppr_tone_mapping.argtypes = [_vp, C.c_int, C.c_int, C.c_float, C.c_float, _vp]
    return _pp
