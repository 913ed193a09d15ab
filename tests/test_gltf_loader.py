"""glTF ingestion + animation on the boundary (Scene::load_gltf / update_animation,
fredholm/src/scene.cpp:445-898): our own JSON / accessor / TRS code must fill the flat arrays
exactly like the reference's tinygltf + glm based loader, quirks included (per-node mesh copies,
(u, 1-v) texcoords, emission = 1, clearcoat texture ids read as 0, un-normalised keyframe weight,
root-only animation targets).  Host only -- no GPU needed."""
import json
import os

import numpy as np
import pytest

from fredholm_b200 import api, scenes
from test_scene_loader import FIELDS, assert_same, load_both

PIL = pytest.importorskip("PIL.Image")


def small_scene():
    return scenes.standard_surface_scene(16, 8, sphere_res=(8, 4))


def test_gltf_static_scene_matches_reference(oracle_mod, tmp_path):
    s = small_scene()
    for embed in (False, True):
        p = scenes.write_gltf(s, str(tmp_path / ("e%d" % embed)), "std", embed=embed, clearcoat=True)
        ours, ref = load_both(oracle_mod, p)
        assert_same(ours, ref)
        assert ours.n_faces == s.n_faces
        assert (ours.materials["emission"] == 1.0).all()          # tinygltf always has emissiveFactor
        assert np.allclose(ours.materials["coat"], s.materials["coat"])


def test_gltf_node_hierarchy_and_transforms(oracle_mod, tmp_path):
    s = scenes.cornell_box()
    n = len(s.submesh_offsets)
    c, si = np.cos(0.3), np.sin(0.3)
    xf = [dict(translation=(0.5, -0.25, 2.0), rotation=(0.0, np.sin(0.4), 0.0, np.cos(0.4)), scale=(1.0, 2.0, 0.5)),
          dict(matrix=(c, 0, -si, 0, 0, 1, 0, 0, si, 0, c, 0, 1.5, 0.25, -3.0, 1)),
          dict(scale=(0.5, 0.5, 0.5)),
          dict(rotation=(0.1825742, 0.3651484, 0.5477226, 0.7302967))][:n]
    parents = {1: 0, 2: 1} if n >= 3 else {}
    p = scenes.write_gltf(s, str(tmp_path), "tree", node_transforms=xf, parents=parents)
    ours, ref = load_both(oracle_mod, p)
    assert_same(ours, ref)
    # the grandchild's world matrix is the product down the chain
    assert not np.allclose(ours.transforms[2].reshape(4, 4), np.eye(4))


def test_gltf_textures_and_material_slots(oracle_mod, tmp_path):
    s = small_scene()
    rng = np.random.default_rng(5)
    files = []
    for i, (w, h, mode) in enumerate([(16, 8, "RGBA"), (5, 9, "RGB"), (12, 12, "L")]):
        a = rng.integers(0, 256, (h, w, {"RGBA": 4, "RGB": 3, "L": 1}[mode]), dtype=np.uint8)
        name = "tex%d.%s" % (i, "jpg" if mode == "RGB" else "png")
        PIL.fromarray(a if a.shape[2] > 1 else a[..., 0], mode).save(tmp_path / name)
        files.append(name)
    s.materials["base_color_texture_id"][0] = 0
    s.materials["metallic_roughness_texture_id"][0] = 2
    s.materials["normalmap_texture_id"][1] = 1
    s.materials["emission_texture_id"][2] = 0
    p = scenes.write_gltf(s, str(tmp_path), "tex", image_files=files)
    ours, ref = load_both(oracle_mod, p)
    assert_same(ours, ref)
    assert len(ours.textures) == 3
    for (a, ca), (b, cb) in zip(ours.textures, ref.textures):
        assert ca == cb is False                 # all glTF textures are NONCOLOR (scene.cpp:564-566)
        assert np.array_equal(a, b)


def test_gltf_clearcoat_texture_quirk(oracle_mod, tmp_path):
    s = scenes.cornell_box()
    p = scenes.write_gltf(s, str(tmp_path), "cc", clearcoat=True)
    doc = json.load(open(p))
    doc["materials"][0]["extensions"]["KHR_materials_clearcoat"]["clearcoatTexture"] = {"index": 3}
    doc["materials"][0]["extensions"]["KHR_materials_clearcoat"]["clearcoatRoughnessTexture"] = {"index": 2}
    del doc["materials"][1]["pbrMetallicRoughness"]        # defaults: white, rough 1, metal 1
    json.dump(doc, open(p, "w"))
    ours, ref = load_both(oracle_mod, p)
    assert_same(ours, ref)
    assert ours.materials["coat_texture_id"][0] == 0 and ours.materials["coat_roughness_texture_id"][0] == 0
    assert ours.materials["metalness"][1] == 1.0 and ours.materials["specular_roughness"][1] == 1.0
    # a primitive without a material gets id -1 = 0xffffffff (scene.cpp:807-809; the reference then reads
    # materials[-1] while building its light list, so only our loader is exercised on this one)
    del doc["meshes"][0]["primitives"][0]["material"]
    json.dump(doc, open(p, "w"))
    sc = api.Scene()
    sc.load_model(p)
    assert sc.arrays().material_ids.max() == 0xffffffff
    sc.close()


def test_gltf_append_to_obj(oracle_mod, tmp_path):
    """rtcamp8 loads an .obj and then a .gltf with clear = false (rtcamp8.cpp:120-121)."""
    obj = scenes.write_obj(scenes.cornell_box(), str(tmp_path), "c")
    g = scenes.write_gltf(small_scene(), str(tmp_path), "g", node_transforms=[dict(translation=(0, 1, 0))])
    sc = api.Scene()
    sc.load_model(obj)
    sc.load_model(g, clear=False)
    a = sc.arrays()
    sc.close()
    o = oracle_mod.Oracle()
    o.load_scene(obj)
    o.load_scene(g, clear=False)
    assert_same(a, o.get_loaded_scene())


ANIM = [dict(node=0, translation=([0.0, 0.5, 1.25, 2.0], [(0, 0, 0), (1, 0, 0), (1, 2, 0), (0, 0, 3)]),
             rotation=([0.0, 1.0, 2.0], [(0, 0, 0, 1), (0, 0.7071068, 0, 0.7071068), (0, 1, 0, 0)]),
             scale=([0.0, 2.0], [(1, 1, 1), (2, 0.5, 1.5)])),
        dict(node=1, rotation=([0.0, 0.75, 1.5], [(0, 0, 0, 1), (0, 0, 0, 1), (0.5, 0.5, 0.5, 0.5)]))]


@pytest.mark.parametrize("time", [0.0, 0.25, 0.5, 0.9, 1.25, 1.99, 2.0, 3.7, 11.3])
def test_gltf_animation_matches_reference(oracle_mod, tmp_path, time):
    s = scenes.cornell_box()
    p = scenes.write_gltf(s, str(tmp_path), "anim", animations=ANIM, parents={2: 0},
                          node_transforms=[None, None, dict(translation=(0.0, 0.5, 0.0))])
    sc = api.Scene()
    sc.load_model(p)
    sc.update_animation(time)
    ours = sc.arrays()
    sc.close()
    o = oracle_mod.Oracle()
    o.load_scene(p)
    o.set_time(time)
    ref = o.get_loaded_scene()
    assert np.array_equal(ours.transforms, ref.transforms), (ours.transforms - ref.transforms)
    assert_same(ours, ref)


def test_gltf_animation_of_child_node_is_rejected(oracle_mod, tmp_path):
    """find_node only returns root nodes (scene.cpp:876-886): both loaders refuse the file."""
    p = scenes.write_gltf(scenes.cornell_box(), str(tmp_path), "bad", parents={1: 0},
                          animations=[dict(node=1, translation=([0.0, 1.0], [(0, 0, 0), (1, 1, 1)]))])
    sc = api.Scene()
    with pytest.raises(api.FredholmError):
        sc.load_model(p)
    sc.close()
    with pytest.raises(RuntimeError):
        oracle_mod.Oracle().load_scene(p)


def test_gltf_errors(tmp_path):
    s = scenes.cornell_box()
    p = scenes.write_gltf(s, str(tmp_path), "e")
    doc = json.load(open(p))
    sc = api.Scene()
    # 32-bit indices are refused ("indices stride is not ushort", scene.cpp:743-746)
    bad = json.loads(json.dumps(doc))
    bad["accessors"][bad["meshes"][0]["primitives"][0]["indices"]]["componentType"] = 5125
    json.dump(bad, open(tmp_path / "e32.gltf", "w"))
    with pytest.raises(api.FredholmError, match="ushort"):
        sc.load_model(str(tmp_path / "e32.gltf"))
    # broken JSON, missing buffer file
    (tmp_path / "broken.gltf").write_text("{ \"asset\": ")
    with pytest.raises(api.FredholmError):
        sc.load_model(str(tmp_path / "broken.gltf"))
    os.remove(tmp_path / "e.bin")
    with pytest.raises(api.FredholmError):
        sc.load_model(p)
    sc.close()
