"""Scene::validate (round-2 hardening): everything the device code indexes with scene data is checked on
the host before an upload, so a malformed scene fails with "invalid scene: ..." instead of an
out-of-bounds device read.  Also the glTF loader's handling of primitives without NORMAL / TEXCOORD_0
and of cyclic node graphs.  Host only -- no GPU needed."""
import copy
import json

import numpy as np
import pytest

from fredholm_b200 import api, scenes
from fredholm_b200.api import FredholmError


def _scene(s):
    sc = api.Scene()
    sc.set_arrays(s)
    return sc


def test_well_formed_scenes_validate():
    for s in (scenes.cornell_box(), scenes.standard_surface_scene(16, 8, sphere_res=(8, 4))):
        _scene(s).validate()


def _broken(mutate):
    s = copy.deepcopy(scenes.cornell_box())
    mutate(s)
    return s


@pytest.mark.parametrize("name,mutate,what", [
    ("vertex index", lambda s: s.indices.__setitem__((3, 1), len(s.vertices)), "vertex index out of range"),
    ("material id", lambda s: s.material_ids.__setitem__(0, len(s.materials)), "valid material"),
    ("instance id", lambda s: s.instance_ids.__setitem__(2, len(s.submesh_offsets)), "instance id out of range"),
    ("submesh overrun", lambda s: s.submesh_n_faces.__setitem__(-1, s.submesh_n_faces[-1] + 1), "exceeds the face array"),
    ("submesh overlap", lambda s: s.submesh_offsets.__setitem__(-1, s.submesh_offsets[-1] - 1), "overlap"),
    ("submesh gap", lambda s: s.submesh_n_faces.__setitem__(0, s.submesh_n_faces[0] - 1), "belongs to no sub-mesh"),
    ("texture id", lambda s: s.materials["base_color_texture_id"].__setitem__(0, 0), "texture id 0 out of range"),
    ("texture id below -1", lambda s: s.materials["alpha_texture_id"].__setitem__(0, -2), "texture id -2 out of range"),
])
def test_malformed_scene_is_rejected(name, mutate, what):
    sc = _scene(_broken(mutate))
    with pytest.raises(FredholmError, match="invalid scene: .*" + what):
        sc.validate()


def test_array_length_mismatch_is_rejected_before_the_abi():
    s = copy.deepcopy(scenes.cornell_box())
    s.normals = s.normals[:-1]
    with pytest.raises(FredholmError, match="invalid scene: normals"):
        api.Scene().set_arrays(s)
    s = copy.deepcopy(scenes.cornell_box())
    s.transforms = s.transforms[:-1]
    with pytest.raises(FredholmError, match="invalid scene: transforms"):
        api.Scene().set_arrays(s)


def test_gltf_primitive_without_normals_or_texcoords_is_padded(tmp_path):
    s = scenes.cornell_box()
    p = scenes.write_gltf(s, str(tmp_path), "nonormal")
    doc = json.load(open(p))
    for mesh in doc["meshes"]:
        for prim in mesh["primitives"]:
            prim["attributes"].pop("NORMAL", None)
            prim["attributes"].pop("TEXCOORD_0", None)
    json.dump(doc, open(p, "w"))
    sc = api.Scene()
    sc.load_model(p)
    sc.validate()
    a = sc.arrays()
    assert len(a.normals) == len(a.vertices) == len(a.texcoords)
    assert np.allclose(np.linalg.norm(a.normals, axis=1), 1.0, atol=1e-5)
    assert (a.texcoords == 0).all()
    # the generated normals are the area-weighted face normals: for the box's planar walls, the wall normal
    f = a.indices[0]
    e1, e2 = a.vertices[f[1]] - a.vertices[f[0]], a.vertices[f[2]] - a.vertices[f[0]]
    n = np.cross(e1, e2)
    assert abs(abs(np.dot(n / np.linalg.norm(n), a.normals[f[0]])) - 1.0) < 1e-5


def test_gltf_cyclic_node_graph_fails_instead_of_recursing(tmp_path):
    s = scenes.cornell_box()
    p = scenes.write_gltf(s, str(tmp_path), "cycle")
    doc = json.load(open(p))
    doc["nodes"][0]["children"] = [0]
    json.dump(doc, open(p, "w"))
    with pytest.raises(FredholmError, match="cycle"):
        api.Scene().load_model(p)
