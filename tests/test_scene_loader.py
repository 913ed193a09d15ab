"""Scene ingestion on the boundary (Scene::load_model, scene.cpp:103-443): our own .obj/.mtl
parser must produce exactly the flat arrays the reference's tinyobjloader-based loader
produces (vertex de-duplication order, default normals / texcoords, MTL -> Material incl. the
custom keys and the reference's quirks).  Host only -- no GPU needed."""
import os

import numpy as np
import pytest

from fredholm_b200 import api, scenes

FIELDS = ("vertices", "normals", "texcoords", "indices", "material_ids", "instance_ids", "submesh_offsets",
          "submesh_n_faces", "transforms")


def load_both(oracle_mod, path):
    sc = api.Scene()
    sc.load_model(path)
    ours = sc.arrays()
    sc.close()
    o = oracle_mod.Oracle()
    o.load_scene(path)
    return ours, o.get_loaded_scene()


def assert_same(a, b):
    for k in FIELDS:
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert a.materials.tobytes() == b.materials.tobytes()
    assert len(a.textures) == len(b.textures)


@pytest.mark.parametrize("with_attributes", [True, False])
def test_obj_cornell(oracle_mod, tmp_path, with_attributes):
    p = scenes.write_obj(scenes.cornell_box(), str(tmp_path), "cornell", with_attributes=with_attributes)
    ours, ref = load_both(oracle_mod, p)
    assert_same(ours, ref)
    assert ours.n_faces == 32 and len(ours.submesh_offsets) == 4
    # emissive material came through Ke (light list is every emissive face, renderer.h:388-402)
    assert (ours.materials["emission_color"].sum(axis=1) > 0).sum() == 1


def test_obj_standard_surface(oracle_mod, tmp_path):
    s = scenes.standard_surface_scene(16, 8, sphere_res=(8, 4))
    p = scenes.write_obj(s, str(tmp_path), "std")
    ours, ref = load_both(oracle_mod, p)
    assert_same(ours, ref)
    # custom MTL keys round-trip (scene.cpp:251-312)
    for k in ("sheen", "coat", "metalness", "transmission", "specular_roughness", "diffuse_roughness"):
        assert np.allclose(ours.materials[k], s.materials[k], atol=1e-6), k


def test_obj_quirks(oracle_mod, tmp_path):
    """Hand-written file: negative indices, polygons (fan triangulation), missing vt/vn,
    comments, `Pcr`, `d`, several usemtl in one object, object without a material."""
    (tmp_path / "q.mtl").write_text(
        "# comment\nnewmtl a\nKd 0.1 0.2 0.3\nKs 0.5 0.5 0.5\nPr 0.35\nPm 0.25\nPc 0.8\nPcr 0.15\nd 0.25\n"
        "Tf 0.9 0.8 0.7\nKe 0 0 0\n\nnewmtl b\nKd 1 1 1\nKe 2 3 4\nsheen 0.5\nsheen_color 0.1 0.2 0.3\n"
        "sheen_roughness 0.7\nsubsurface 0.2\nsubsurface_color 0.3 0.4 0.5\nthin_walled 1\ndiffuse 0.75\n"
        "diffuse_roughness 0.4\n")
    (tmp_path / "q.obj").write_text(
        "mtllib q.mtl\n# a quad, a pentagon and a triangle\n"
        "v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0.5 1.5 0\nv 0 0 1\nv 1 0 1\nv 0 1 1\n"
        "vn 0 0 1\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\n"
        "o first\nusemtl a\nf 1/1/1 2/2/1 3/3/1 4/4/1\nusemtl b\nf 1 2 3 5 4\n"
        "o second\nusemtl a\nf -3//1 -2//1 -1//1\nf 6/1 7/2 8/3\n")
    ours, ref = load_both(oracle_mod, str(tmp_path / "q.obj"))
    assert_same(ours, ref)
    assert ours.n_faces == 2 + 3 + 2
    a = ours.materials[0]
    assert np.isclose(a["transmission"], 0.75) and np.isclose(a["coat"], 0.8)


def test_invalid_scene(tmp_path):
    sc = api.Scene()
    with pytest.raises(api.FredholmError):
        sc.load_model(str(tmp_path / "missing.obj"))
    (tmp_path / "x.ply").write_text("ply\n")
    with pytest.raises(api.FredholmError):
        sc.load_model(str(tmp_path / "x.ply"))   # reference: std::runtime_error("invalid scene")
    sc.close()


def test_append_without_clear(oracle_mod, tmp_path):
    p = scenes.write_obj(scenes.cornell_box(), str(tmp_path), "c")
    sc = api.Scene()
    sc.load_model(p)
    sc.load_model(p, clear=False)
    a = sc.arrays()
    sc.close()
    o = oracle_mod.Oracle()
    o.load_scene(p)
    o.load_scene(p, clear=False)
    assert_same(a, o.get_loaded_scene())
    assert a.n_faces == 64


@pytest.mark.gpu
def test_gpu_vertex_dedup_and_face_normals_equal_the_host_loop(oracle_mod, tmp_path, monkeypatch):
    """csrc/mesh_prep.cu (expansion, face-normal generation, vertex de-duplication as kernels) against the host loop
    of csrc/scene.cpp and the reference's loader: identical arrays, same vertex order (first occurrence), for a
    mesh WITH normals / texcoords and for one WITHOUT (face normals, default texcoords, many duplicate corners)."""
    import time
    s = scenes.standard_surface_scene(96, 48, sphere_res=(24, 12))
    for with_attributes in (True, False):
        p = scenes.write_obj(s, str(tmp_path / ("a%d" % with_attributes)), "mesh", with_attributes=with_attributes)
        out = {}
        for mode in ("0", "1"):
            monkeypatch.setenv("FRD_GPU_MESH_PREP", mode)
            t0 = time.time()
            sc = api.Scene()
            sc.load_model(p)
            out[mode] = (sc.arrays(), time.time() - t0)
        a, b = out["0"][0], out["1"][0]
        for k in FIELDS:
            x, y = getattr(a, k), getattr(b, k)
            assert x.shape == y.shape and x.tobytes() == y.tobytes(), (with_attributes, k)
        ours, ref = load_both(oracle_mod, p)          # the reference's loader (tinyobjloader + unordered_map)
        for k in FIELDS:                              # bytes, not ==: degenerate triangles have NaN face normals
            assert getattr(ours, k).tobytes() == getattr(ref, k).tobytes(), (with_attributes, k)
        print("obj with_attributes=%s: %d faces, %d unique vertices, host loop %.2f s, kernels %.2f s (incl. parse)"
              % (with_attributes, a.n_faces, len(a.vertices), out["0"][1], out["1"][1]))


def test_obj_degenerate_triangle_without_normals_maps_to_vertex_zero(oracle_mod, tmp_path):
    """Reference quirk (scene.cpp:363-387): a zero-area triangle without `vn` gets a NaN face normal; a NaN vertex
    equals nothing, so the reference appends it but `unique_vertices[vertex]` hands its corner index 0."""
    (tmp_path / "deg.mtl").write_text("newmtl m\nKd 0.5 0.5 0.5\n")
    p = tmp_path / "deg.obj"
    p.write_text("mtllib deg.mtl\no s\nv 0 0 0\nv 1 0 0\nv 0 1 0\nv 2 2 2\nv 3 3 3\nv 4 4 4\nusemtl m\nf 1 2 3\nf 4 5 6\nf 1 3 2\n")
    ours, ref = load_both(oracle_mod, str(p))
    for k in FIELDS:
        assert getattr(ours, k).tobytes() == getattr(ref, k).tobytes(), k
    assert ours.indices[1].tolist() == [0, 0, 0] and np.isnan(ours.normals).any()
    assert len(ours.vertices) == 9          # 3 + 3 NaN vertices (appended, unreferenced) + 3
