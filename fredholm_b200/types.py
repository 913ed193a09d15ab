"""Plain-data types shared with the C ABI (numpy views of the C structs)."""
from dataclasses import dataclass, field

import numpy as np

# fredholm::Material, include/fredholm/shared.h (reference shared.h:100-142), 180 bytes
MATERIAL_DTYPE = np.dtype([
    ("diffuse", "<f4"), ("base_color", "<f4", 3), ("base_color_texture_id", "<i4"),
    ("diffuse_roughness", "<f4"),
    ("specular", "<f4"), ("specular_color", "<f4", 3), ("specular_color_texture_id", "<i4"),
    ("specular_roughness", "<f4"), ("specular_roughness_texture_id", "<i4"),
    ("metalness", "<f4"), ("metalness_texture_id", "<i4"),
    ("metallic_roughness_texture_id", "<i4"),
    ("coat", "<f4"), ("coat_texture_id", "<i4"), ("coat_color", "<f4", 3),
    ("coat_roughness", "<f4"), ("coat_roughness_texture_id", "<i4"),
    ("transmission", "<f4"), ("transmission_color", "<f4", 3),
    ("sheen", "<f4"), ("sheen_color", "<f4", 3), ("sheen_roughness", "<f4"),
    ("subsurface", "<f4"), ("subsurface_color", "<f4", 3),
    ("thin_walled", "<f4"),
    ("emission", "<f4"), ("emission_color", "<f4", 3), ("emission_texture_id", "<i4"),
    ("heightmap_texture_id", "<i4"), ("normalmap_texture_id", "<i4"), ("alpha_texture_id", "<i4"),
])
assert MATERIAL_DTYPE.itemsize == 180

_MATERIAL_DEFAULTS = dict(
    diffuse=1.0, base_color=(1, 1, 1), base_color_texture_id=-1, diffuse_roughness=0.0,
    specular=1.0, specular_color=(1, 1, 1), specular_color_texture_id=-1,
    specular_roughness=0.2, specular_roughness_texture_id=-1,
    metalness=0.0, metalness_texture_id=-1, metallic_roughness_texture_id=-1,
    coat=0.0, coat_texture_id=-1, coat_color=(1, 1, 1), coat_roughness=0.1,
    coat_roughness_texture_id=-1,
    transmission=0.0, transmission_color=(1, 1, 1),
    sheen=0.0, sheen_color=(1, 1, 1), sheen_roughness=0.3,
    subsurface=0.0, subsurface_color=(1, 1, 1), thin_walled=0.0,
    emission=0.0, emission_color=(0, 0, 0), emission_texture_id=-1,
    heightmap_texture_id=-1, normalmap_texture_id=-1, alpha_texture_id=-1,
)


def make_material(**kw):
    """One Material record with the reference's defaults (shared.h:100-142)."""
    m = np.zeros((), dtype=MATERIAL_DTYPE)
    vals = dict(_MATERIAL_DEFAULTS)
    for k in kw:
        if k not in vals:
            raise KeyError(k)
    vals.update(kw)
    for k, v in vals.items():
        m[k] = v
    return m


# reference ShadingParams (shared.h:173-199) as 30 floats, for the BSDF unit vectors
SHADING_PARAM_FIELDS = [
    ("diffuse", 1), ("base_color", 3), ("diffuse_roughness", 1), ("specular", 1),
    ("specular_color", 3), ("specular_roughness", 1), ("metalness", 1), ("coat", 1),
    ("coat_color", 3), ("coat_roughness", 1), ("transmission", 1), ("transmission_color", 3),
    ("sheen", 1), ("sheen_color", 3), ("sheen_roughness", 1), ("subsurface", 1),
    ("subsurface_color", 3), ("thin_walled", 1),
]


def shading_params(**kw):
    d = dict(diffuse=1.0, base_color=(0, 0, 0), diffuse_roughness=0.0, specular=1.0,
             specular_color=(0, 0, 0), specular_roughness=0.2, metalness=0.0, coat=0.0,
             coat_color=(1, 1, 1), coat_roughness=0.1, transmission=0.0,
             transmission_color=(1, 1, 1), sheen=0.0, sheen_color=(1, 1, 1), sheen_roughness=0.3,
             subsurface=0.0, subsurface_color=(1, 1, 1), thin_walled=0.0)
    for k in kw:
        if k not in d:
            raise KeyError(k)
    d.update(kw)
    out = []
    for name, n in SHADING_PARAM_FIELDS:
        v = np.atleast_1d(np.asarray(d[name], dtype=np.float32))
        assert v.size == n, name
        out.extend(v.tolist())
    return np.asarray(out, dtype=np.float32)


@dataclass
class SceneArrays:
    """The flat arrays of fredholm::Scene the renderer consumes (scene.h:107-130)."""
    vertices: np.ndarray          # (V,3) f32
    normals: np.ndarray           # (V,3) f32
    texcoords: np.ndarray         # (V,2) f32
    indices: np.ndarray           # (F,3) u32, global vertex ids
    material_ids: np.ndarray      # (F,) u32
    materials: np.ndarray         # (M,) MATERIAL_DTYPE
    submesh_offsets: np.ndarray   # (S,) u32
    submesh_n_faces: np.ndarray   # (S,) u32
    instance_ids: np.ndarray = None   # (F,) u32 (defaults to 0, like the .obj loader)
    transforms: np.ndarray = None     # (S,16) f32 column-major mat4 (defaults to identity)
    textures: list = field(default_factory=list)  # [(rgba8 (H,W,4) u8, is_color)]

    def __post_init__(self):
        self.vertices = np.ascontiguousarray(self.vertices, dtype=np.float32).reshape(-1, 3)
        self.normals = np.ascontiguousarray(self.normals, dtype=np.float32).reshape(-1, 3)
        self.texcoords = np.ascontiguousarray(self.texcoords, dtype=np.float32).reshape(-1, 2)
        self.indices = np.ascontiguousarray(self.indices, dtype=np.uint32).reshape(-1, 3)
        self.material_ids = np.ascontiguousarray(self.material_ids, dtype=np.uint32)
        self.materials = np.ascontiguousarray(self.materials, dtype=MATERIAL_DTYPE)
        self.submesh_offsets = np.ascontiguousarray(self.submesh_offsets, dtype=np.uint32)
        self.submesh_n_faces = np.ascontiguousarray(self.submesh_n_faces, dtype=np.uint32)
        if self.instance_ids is None:
            self.instance_ids = np.zeros(len(self.indices), dtype=np.uint32)
        self.instance_ids = np.ascontiguousarray(self.instance_ids, dtype=np.uint32)
        if self.transforms is None:
            self.transforms = np.tile(np.eye(4, dtype=np.float32).reshape(1, 16),
                                      (len(self.submesh_offsets), 1))
        self.transforms = np.ascontiguousarray(self.transforms, dtype=np.float32).reshape(-1, 16)
        assert len(self.normals) == len(self.vertices) == len(self.texcoords)
        assert len(self.material_ids) == len(self.indices) == len(self.instance_ids)
        assert len(self.transforms) == len(self.submesh_offsets) == len(self.submesh_n_faces)

    @property
    def n_faces(self):
        return len(self.indices)
