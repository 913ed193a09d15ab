"""Procedural scenes for the tests and the benchmark (no network, no assets: the
reference ships no scenes either).

Every generator is deterministic (fixed integer hashes, no RNG state) and returns
SceneArrays -- exactly the flat arrays fredholm::Scene holds -- so the same data can
be handed to the CUDA core (fr_set_scene_arrays) and to the host oracle
(orc_set_scene).  `write_obj` serialises a scene as .obj/.mtl text to exercise the
file loaders on both sides.

  cornell_box()            BASELINE.json config 1: ~32 triangles, Lambert-like + quad light
  standard_surface_scene() BASELINE.json config 2/3: terrain + spheres, 1 048 576
                           triangles by default, 8 Standard-Surface materials
                           (metal / coat / rough glass / sheen / diffuse)
  instanced_scene()        BASELINE.json config 4: 52 M triangles, terrain + 3072 transformed
                           placements of a 16 K-triangle mesh, all diffuse
  textured_scene()         BASELINE.json config 5: config 2's scene with procedural base-colour
                           / roughness / normal-map textures
"""
import os

import numpy as np

from .types import MATERIAL_DTYPE, SceneArrays, make_material


# --------------------------------------------------------------------------------------
def _quad(p0, p1, p2, p3):
    """Two triangles (p0,p1,p2), (p0,p2,p3)."""
    return [(p0, p1, p2), (p0, p2, p3)]


def _assemble(tris_by_shape, materials):
    """tris_by_shape: list of shapes; each shape = list of (tri (3x3), material id).
    Vertices are de-duplicated on (position, normal, texcoord) in order of first use,
    like the reference's .obj path (scene.cpp:317-393); normals are face normals and
    texcoords the loader's defaults for faces without vt (scene.cpp:363-379)."""
    verts, lookup = [], {}
    indices, mat_ids, offsets, counts = [], [], [], []
    for shape in tris_by_shape:
        offsets.append(len(indices))
        for tri, mid in shape:
            p = np.asarray(tri, dtype=np.float32)
            e1 = p[1] - p[0]
            e2 = p[2] - p[0]
            e1 = e1 * (np.float32(1) / np.sqrt(np.dot(e1, e1)))
            e2 = e2 * (np.float32(1) / np.sqrt(np.dot(e2, e2)))
            n = np.cross(e1, e2).astype(np.float32)
            n = n * (np.float32(1) / np.sqrt(np.dot(n, n)))
            uv = [(0.0, 0.0), (1.0, 0.0), (0.0, 1.0)]
            ids = []
            for k in range(3):
                key = (p[k].tobytes(), n.astype(np.float32).tobytes(), uv[k])
                if key not in lookup:
                    lookup[key] = len(verts)
                    verts.append((p[k], n, uv[k]))
                ids.append(lookup[key])
            indices.append(ids)
            mat_ids.append(mid)
        counts.append(len(indices) - offsets[-1])
    return SceneArrays(
        vertices=np.array([v[0] for v in verts], np.float32),
        normals=np.array([v[1] for v in verts], np.float32),
        texcoords=np.array([v[2] for v in verts], np.float32),
        indices=np.array(indices, np.uint32), material_ids=np.array(mat_ids, np.uint32),
        materials=np.array(materials, dtype=MATERIAL_DTYPE),
        submesh_offsets=np.array(offsets, np.uint32), submesh_n_faces=np.array(counts, np.uint32))


def _box(cx, cz, sx, sy, sz, angle_deg):
    """Axis box of size (sx,sy,sz) standing on y=0, rotated about y; bottom face omitted."""
    a = np.deg2rad(angle_deg)
    c, s = np.cos(a), np.sin(a)

    def P(x, y, z):
        return (cx + c * x + s * z, y, cz - s * x + c * z)

    hx, hz = sx / 2, sz / 2
    b = [P(-hx, 0, -hz), P(hx, 0, -hz), P(hx, 0, hz), P(-hx, 0, hz)]
    t = [P(-hx, sy, -hz), P(hx, sy, -hz), P(hx, sy, hz), P(-hx, sy, hz)]
    tris = []
    tris += _quad(t[0], t[3], t[2], t[1])  # top (normal +y)
    tris += _quad(b[3], b[2], t[2], t[3])  # +z side
    tris += _quad(b[2], b[1], t[1], t[2])  # +x side
    tris += _quad(b[1], b[0], t[0], t[1])  # -z side
    tris += _quad(b[0], b[3], t[3], t[0])  # -x side
    return tris


def cornell_box():
    """Cornell box in x in [-1,1], y in [0,2], z in [-1,1], open towards +z.
    Materials are Lambert-like: Kd set, specular colour 0 (specular lobe off),
    metalness 0, opaque; one quad light with emission (17, 12, 4)."""
    white = make_material(base_color=(0.725, 0.71, 0.68), specular_color=(0, 0, 0))
    red = make_material(base_color=(0.63, 0.065, 0.05), specular_color=(0, 0, 0))
    green = make_material(base_color=(0.14, 0.45, 0.091), specular_color=(0, 0, 0))
    light = make_material(base_color=(0.78, 0.78, 0.78), specular_color=(0, 0, 0), emission=1.0,
                          emission_color=(17, 12, 4))
    materials = [white, red, green, light]
    W, R, G, Lm = 0, 1, 2, 3
    room = []
    room += [(t, W) for t in _quad((-1, 0, 1), (1, 0, 1), (1, 0, -1), (-1, 0, -1))]    # floor (+y)
    room += [(t, W) for t in _quad((-1, 2, -1), (1, 2, -1), (1, 2, 1), (-1, 2, 1))]    # ceiling (-y)
    room += [(t, W) for t in _quad((-1, 0, -1), (1, 0, -1), (1, 2, -1), (-1, 2, -1))]  # back (+z)
    room += [(t, R) for t in _quad((-1, 0, 1), (-1, 0, -1), (-1, 2, -1), (-1, 2, 1))]  # left (+x)
    room += [(t, G) for t in _quad((1, 0, -1), (1, 0, 1), (1, 2, 1), (1, 2, -1))]      # right (-x)
    lamp = [(t, Lm) for t in _quad((-0.25, 1.98, -0.25), (0.25, 1.98, -0.25), (0.25, 1.98, 0.25),
                                   (-0.25, 1.98, 0.25))]                               # light (-y)
    short = [(t, W) for t in _box(0.33, 0.35, 0.6, 0.6, 0.6, -18.0)]
    tall = [(t, W) for t in _box(-0.35, -0.3, 0.6, 1.2, 0.6, 17.0)]
    return _assemble([room, lamp, short, tall], materials)


# The thin-lens model puts the lens 1/tan(fov/2) behind `origin` (camera.cu:24-53): with
# fov 45 deg the eye sits at z = 1.15 + 2.414 and the box opening just fills the frame.
CORNELL_CAMERA = dict(origin=(0.0, 1.0, 1.15), fov=np.deg2rad(45.0), F=100.0, focus=10000.0)


# --------------------------------------------------------------------------------------
def _hash_u32(x):
    """Integer hash (lowbias32) on uint32 arrays."""
    x = np.asarray(x, dtype=np.uint64) & 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x.astype(np.uint32)


def _hash01(x):
    return (_hash_u32(x).astype(np.float64) / 4294967296.0).astype(np.float32)


def standard_surface_materials():
    """8 materials cycling metal / coated dielectric / rough glass / sheen / diffuse."""
    return [
        make_material(base_color=(0.35, 0.37, 0.30), specular_color=(0.04, 0.04, 0.04),
                      specular_roughness=0.5, diffuse_roughness=0.3),                          # 0 ground
        make_material(base_color=(0.95, 0.64, 0.54), specular_color=(1, 1, 1), metalness=1.0,
                      specular_roughness=0.05),                                                # 1 polished copper
        make_material(base_color=(0.91, 0.92, 0.92), specular_color=(1, 1, 1), metalness=1.0,
                      specular_roughness=0.5),                                                 # 2 rough aluminium
        make_material(base_color=(0.7, 0.05, 0.05), specular_color=(1, 1, 1), coat=1.0,
                      coat_roughness=0.05, specular_roughness=0.3),                            # 3 car paint
        make_material(base_color=(1, 1, 1), specular_color=(1, 1, 1), transmission=0.9,
                      transmission_color=(0.9, 1.0, 0.95), specular_roughness=0.15),           # 4 rough glass
        make_material(base_color=(0.1, 0.15, 0.5), specular_color=(0, 0, 0), sheen=1.0,
                      sheen_color=(0.8, 0.8, 1.0), sheen_roughness=0.3),                       # 5 sheen cloth
        make_material(base_color=(0.8, 0.8, 0.2), specular_color=(0, 0, 0)),                   # 6 plain diffuse
        make_material(base_color=(0.2, 0.6, 0.3), specular_color=(1, 1, 1), specular_roughness=0.1,
                      coat=0.5, coat_color=(1.0, 0.9, 0.8)),                                   # 7 glossy plastic
    ]


def standard_surface_scene(terrain_res=512, n_spheres=512, sphere_res=(32, 16), seed=0xF2ED401):
    """Terrain (terrain_res^2 quads) + n_spheres UV spheres of 2*sphere_res[0]*sphere_res[1]
    triangles.  Defaults: 524 288 + 512 * 1024 = 1 048 576 triangles.  One sub-mesh for the
    terrain and one per sphere, identity transforms."""
    R = terrain_res
    ext = 40.0
    # ---- terrain: value-noise heightfield on a (R+1)^2 grid ----
    gx, gz = np.meshgrid(np.arange(R + 1), np.arange(R + 1), indexing="xy")
    x = (gx.astype(np.float32) / R - 0.5) * ext
    z = (gz.astype(np.float32) / R - 0.5) * ext

    def value_noise(cells):
        cx = gx.astype(np.float64) / R * cells
        cz = gz.astype(np.float64) / R * cells
        ix, iz = np.floor(cx).astype(np.int64), np.floor(cz).astype(np.int64)
        fx, fz = cx - ix, cz - iz
        fx, fz = fx * fx * (3 - 2 * fx), fz * fz * (3 - 2 * fz)

        def lat(a, b):
            return _hash01((a * 73856093 ^ b * 19349663 ^ (seed + cells)) & 0xFFFFFFFF).astype(np.float64)
        v = (lat(ix, iz) * (1 - fx) + lat(ix + 1, iz) * fx) * (1 - fz) + \
            (lat(ix, iz + 1) * (1 - fx) + lat(ix + 1, iz + 1) * fx) * fz
        return v

    h = (1.6 * value_noise(4) + 0.8 * value_noise(8) + 0.35 * value_noise(16) + 0.12 * value_noise(48))
    y = (h - h.mean()).astype(np.float32)
    tv = np.stack([x, y, z], axis=-1).reshape(-1, 3)
    # normals by central differences
    dx = np.zeros_like(y)
    dz = np.zeros_like(y)
    dx[:, 1:-1] = (y[:, 2:] - y[:, :-2]) / (2 * ext / R)
    dx[:, 0], dx[:, -1] = (y[:, 1] - y[:, 0]) / (ext / R), (y[:, -1] - y[:, -2]) / (ext / R)
    dz[1:-1, :] = (y[2:, :] - y[:-2, :]) / (2 * ext / R)
    dz[0, :], dz[-1, :] = (y[1, :] - y[0, :]) / (ext / R), (y[-1, :] - y[-2, :]) / (ext / R)
    tn = np.stack([-dx, np.ones_like(dx), -dz], axis=-1).reshape(-1, 3)
    tn /= np.linalg.norm(tn, axis=1, keepdims=True)
    tt = np.stack([gx.astype(np.float32) / R * 8, gz.astype(np.float32) / R * 8], axis=-1).reshape(-1, 2)
    i0 = (gz[:-1, :-1] * (R + 1) + gx[:-1, :-1]).reshape(-1)
    i1, i2, i3 = i0 + 1, i0 + (R + 1) + 1, i0 + (R + 1)
    # counter-clockwise seen from +y
    tf = np.concatenate([np.stack([i0, i3, i2], 1), np.stack([i0, i2, i1], 1)], axis=1).reshape(-1, 3)

    verts, norms, texs, faces, mats = [tv], [tn.astype(np.float32)], [tt], [tf], [np.zeros(len(tf), np.uint32)]
    offsets, counts = [0], [len(tf)]
    vbase, fbase = len(tv), len(tf)

    # ---- spheres on a jittered grid, resting on the terrain ----
    nu, nv = sphere_res
    uu, vv = np.meshgrid(np.arange(nu + 1), np.arange(nv + 1), indexing="xy")
    phi = uu.astype(np.float64) / nu * 2 * np.pi
    theta = vv.astype(np.float64) / nv * np.pi
    unit = np.stack([np.sin(theta) * np.cos(phi), np.cos(theta), np.sin(theta) * np.sin(phi)], -1).reshape(-1, 3)
    st = np.stack([uu / nu, vv / nv], -1).reshape(-1, 2).astype(np.float32)
    a = (vv[:-1, :-1] * (nu + 1) + uu[:-1, :-1]).reshape(-1)
    b, c, d = a + 1, a + (nu + 1) + 1, a + (nu + 1)
    sf = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)], axis=1).reshape(-1, 3)
    side = int(np.ceil(np.sqrt(n_spheres)))
    ids = np.arange(n_spheres)
    jx = _hash01(ids * 2 + 1 + seed) - 0.5
    jz = _hash01(ids * 2 + 2 + seed) - 0.5
    rad = 0.28 + 0.42 * _hash01(ids + 977 + seed)
    cxs = ((ids % side + 0.5 + 0.6 * jx) / side - 0.5) * (ext * 0.9)
    czs = ((ids // side + 0.5 + 0.6 * jz) / side - 0.5) * (ext * 0.9)
    # terrain height under the centre (nearest grid vertex)
    ixs = np.clip(np.round((cxs / ext + 0.5) * R).astype(int), 0, R)
    izs = np.clip(np.round((czs / ext + 0.5) * R).astype(int), 0, R)
    cys = y[izs, ixs] + rad * (0.9 + 1.5 * _hash01(ids + 4242 + seed))
    for s in range(n_spheres):
        centre = np.array([cxs[s], cys[s], czs[s]], dtype=np.float64)
        verts.append((centre + rad[s] * unit).astype(np.float32))
        norms.append(unit.astype(np.float32))
        texs.append(st)
        faces.append(sf + vbase)
        mats.append(np.full(len(sf), 1 + (s % 7), np.uint32))
        offsets.append(fbase)
        counts.append(len(sf))
        vbase += len(unit)
        fbase += len(sf)
    return SceneArrays(vertices=np.concatenate(verts), normals=np.concatenate(norms),
                       texcoords=np.concatenate(texs), indices=np.concatenate(faces).astype(np.uint32),
                       material_ids=np.concatenate(mats),
                       materials=np.array(standard_surface_materials(), dtype=MATERIAL_DTYPE),
                       submesh_offsets=np.array(offsets, np.uint32), submesh_n_faces=np.array(counts, np.uint32))


def procedural_textures(res=1024, seed=0xC5):
    """Hash-noise RGBA8 textures for BASELINE.json config 5: [0] base colour (COLOR: sRGB decoded
    on fetch), [1] specular roughness (NONCOLOR, .x), [2] tangent-space normal map (NONCOLOR).
    Multi-octave value noise so that bilinear fetches see structure at every scale."""
    yy, xx = np.meshgrid(np.arange(res), np.arange(res), indexing="ij")

    def octave(cells, salt):
        cx, cy = xx.astype(np.float64) / res * cells, yy.astype(np.float64) / res * cells
        ix, iy = np.floor(cx).astype(np.int64), np.floor(cy).astype(np.int64)
        fx, fy = cx - ix, cy - iy
        fx, fy = fx * fx * (3 - 2 * fx), fy * fy * (3 - 2 * fy)

        def lat(a, b):   # periodic lattice: wrap addressing stays seamless
            return _hash01(((a % cells) * 73856093 ^ (b % cells) * 19349663 ^ (seed * 977 + salt)) & 0xFFFFFFFF
                           ).astype(np.float64)
        return (lat(ix, iy) * (1 - fx) + lat(ix + 1, iy) * fx) * (1 - fy) + \
               (lat(ix, iy + 1) * (1 - fx) + lat(ix + 1, iy + 1) * fx) * fy

    def fbm(salt):
        return (octave(8, salt) + 0.5 * octave(16, salt + 1) + 0.25 * octave(64, salt + 2) +
                0.125 * octave(256, salt + 3)) / 1.875

    def rgba(r, g, b, a=1.0):
        img = np.stack([r, g, b, np.broadcast_to(a, r.shape)], -1)
        return np.ascontiguousarray(np.clip(np.round(img * 255.0), 0, 255).astype(np.uint8))

    n0, n1, n2 = fbm(1), fbm(11), fbm(21)
    checker = (((xx * 16 // res) + (yy * 16 // res)) & 1).astype(np.float64)
    base = rgba(0.25 + 0.6 * n0 * (0.6 + 0.4 * checker), 0.2 + 0.6 * n1, 0.15 + 0.5 * n2 * (1.0 - 0.5 * checker))
    rough = rgba(0.08 + 0.6 * n1, 0.08 + 0.6 * n1, 0.08 + 0.6 * n1)
    hgt = fbm(31)
    dx = np.roll(hgt, -1, axis=1) - np.roll(hgt, 1, axis=1)
    dy = np.roll(hgt, -1, axis=0) - np.roll(hgt, 1, axis=0)
    nrm = np.stack([-dx * res / 64.0, -dy * res / 64.0, np.ones_like(dx)], -1)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    normal = rgba(0.5 + 0.5 * nrm[..., 0], 0.5 + 0.5 * nrm[..., 1], 0.5 + 0.5 * nrm[..., 2])
    return [(base, True), (rough, False), (normal, False)]


def textured_scene(tex_res=1024, **kw):
    """BASELINE.json config 5 content: the Standard-Surface scene with procedurally generated
    textures on the terrain (base colour + roughness + normal map), the car paint (base
    colour) and the plastic (roughness + normal map): exercises CLS_GENERIC_TEX shading."""
    s = standard_surface_scene(**kw)
    s.textures = procedural_textures(tex_res)
    m = s.materials
    m[0]["base_color_texture_id"], m[0]["specular_roughness_texture_id"], m[0]["normalmap_texture_id"] = 0, 1, 2
    m[3]["base_color_texture_id"] = 0
    m[7]["specular_roughness_texture_id"], m[7]["normalmap_texture_id"] = 1, 2
    return s


def instanced_scene(n_instances=3072, mesh_res=(128, 64), terrain_res=1024, seed=0x50C4):
    """BASELINE.json config 4: a unique terrain (2 * terrain_res^2 triangles) plus n_instances
    placements of one UV-sphere-like blob mesh (2 * mesh_res[0] * mesh_res[1] triangles each),
    all diffuse.  Defaults: 2 097 152 + 3072 * 16 384 = 52 428 800 triangles.  Like the
    reference's glTF path (scene.cpp:730-822) every placement is its own
    sub-mesh with its own copy of the mesh and its own transform (instance i = sub-mesh i)."""
    R = terrain_res
    ext = 400.0
    gx, gz = np.meshgrid(np.arange(R + 1), np.arange(R + 1), indexing="xy")
    x = (gx.astype(np.float32) / R - 0.5) * ext
    z = (gz.astype(np.float32) / R - 0.5) * ext
    y = (6.0 * np.sin(x * 0.031) * np.cos(z * 0.027) + 2.0 * np.sin(x * 0.11 + 1.3) * np.sin(z * 0.093)).astype(np.float32)
    tv = np.stack([x, y, z], -1).reshape(-1, 3)
    ddx = np.gradient(y, ext / R, axis=1)
    ddz = np.gradient(y, ext / R, axis=0)
    tn = np.stack([-ddx, np.ones_like(ddx), -ddz], -1).reshape(-1, 3)
    tn /= np.linalg.norm(tn, axis=1, keepdims=True)
    tt = np.stack([gx / R, gz / R], -1).reshape(-1, 2).astype(np.float32)
    i0 = (gz[:-1, :-1] * (R + 1) + gx[:-1, :-1]).reshape(-1)
    i1, i2, i3 = i0 + 1, i0 + (R + 1) + 1, i0 + (R + 1)
    tf = np.concatenate([np.stack([i0, i3, i2], 1), np.stack([i0, i2, i1], 1)], axis=1).reshape(-1, 3)

    nu, nv = mesh_res
    uu, vv = np.meshgrid(np.arange(nu + 1), np.arange(nv + 1), indexing="xy")
    phi = uu.astype(np.float64) / nu * 2 * np.pi
    theta = vv.astype(np.float64) / nv * np.pi
    unit = np.stack([np.sin(theta) * np.cos(phi), np.cos(theta), np.sin(theta) * np.sin(phi)], -1).reshape(-1, 3)
    bump = 1.0 + 0.18 * np.sin(5 * phi).reshape(-1) * np.sin(4 * theta).reshape(-1) ** 2
    mv = (unit * bump[:, None]).astype(np.float32)
    mn = unit.astype(np.float32)
    mt = np.stack([uu / nu, vv / nv], -1).reshape(-1, 2).astype(np.float32)
    a = (vv[:-1, :-1] * (nu + 1) + uu[:-1, :-1]).reshape(-1)
    b, c, d = a + 1, a + (nu + 1) + 1, a + (nu + 1)
    mf = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)], axis=1).reshape(-1, 3)

    n = n_instances
    nmv, nmf = len(mv), len(mf)
    verts = np.concatenate([tv, np.tile(mv, (n, 1))])
    norms = np.concatenate([tn.astype(np.float32), np.tile(mn, (n, 1))])
    texs = np.concatenate([tt, np.tile(mt, (n, 1))])
    inst_faces = (np.tile(mf, (n, 1)).reshape(n, nmf, 3) + (len(tv) + np.arange(n) * nmv)[:, None, None])
    faces = np.concatenate([tf, inst_faces.reshape(-1, 3)]).astype(np.uint32)
    mats = np.concatenate([np.zeros(len(tf), np.uint32), 1 + (np.repeat(np.arange(n), nmf) % 3).astype(np.uint32)])
    offsets = np.concatenate([[0], len(tf) + np.arange(n) * nmf]).astype(np.uint32)
    counts = np.concatenate([[len(tf)], np.full(n, nmf)]).astype(np.uint32)
    # placements: jittered grid over the terrain, random scale and rotation about y
    ids = np.arange(n)
    side = int(np.ceil(np.sqrt(n)))
    px = ((ids % side + 0.5 + 0.7 * (_hash01(ids * 3 + 1 + seed) - 0.5)) / side - 0.5) * ext * 0.95
    pz = ((ids // side + 0.5 + 0.7 * (_hash01(ids * 3 + 2 + seed) - 0.5)) / side - 0.5) * ext * 0.95
    sc = 1.2 + 1.6 * _hash01(ids * 3 + 3 + seed)
    ang = 2 * np.pi * _hash01(ids + 7919 + seed)
    py = 6.0 * np.sin(px * 0.031) * np.cos(pz * 0.027) + 2.0 * np.sin(px * 0.11 + 1.3) * np.sin(pz * 0.093) + sc * 0.8
    tr = np.zeros((n + 1, 4, 4), np.float32)   # column-major: tr[i, col, row]
    tr[0] = np.eye(4)
    ca, sa = np.cos(ang) * sc, np.sin(ang) * sc
    tr[1:, 0, 0], tr[1:, 0, 2] = ca, -sa
    tr[1:, 1, 1] = sc
    tr[1:, 2, 0], tr[1:, 2, 2] = sa, ca
    tr[1:, 3, 0], tr[1:, 3, 1], tr[1:, 3, 2], tr[1:, 3, 3] = px, py, pz, 1.0
    materials = [make_material(base_color=(0.45, 0.42, 0.38), specular_color=(0, 0, 0)),
                 make_material(base_color=(0.75, 0.3, 0.25), specular_color=(0, 0, 0)),
                 make_material(base_color=(0.3, 0.65, 0.35), specular_color=(0, 0, 0)),
                 make_material(base_color=(0.3, 0.4, 0.8), specular_color=(0, 0, 0))]
    return SceneArrays(vertices=verts, normals=norms, texcoords=texs, indices=faces, material_ids=mats,
                       materials=np.array(materials, dtype=MATERIAL_DTYPE), submesh_offsets=offsets,
                       submesh_n_faces=counts, transforms=tr.reshape(n + 1, 16))


INSTANCED_CAMERA = dict(origin=(0.0, 45.0, 230.0), fov=np.deg2rad(55.0), F=100.0, focus=10000.0)


STANDARD_CAMERA = dict(origin=(0.0, 6.0, 22.0), fov=np.deg2rad(50.0), F=16.0, focus=20.0)
# rtcamp8 lighting (rtcamp8.cpp:142-146)
STANDARD_LIGHTING = dict(sun_le=(20.0, 20.0, 20.0), sun_dir=(-0.1, 1.0, 0.1), sun_angle=1.0,
                         turbidity=3.0, albedo=0.3)


# --------------------------------------------------------------------------------------
_MTL_KEYS = ("diffuse", "diffuse_roughness", "sheen", "sheen_color", "sheen_roughness", "subsurface",
             "subsurface_color", "thin_walled")


def write_obj(scene: SceneArrays, directory, name="scene", with_attributes=True):
    """Serialises `scene` as name.obj + name.mtl (one `o` per sub-mesh).  With
    with_attributes=False only positions are written, so the loaders have to supply
    face normals and default texcoords.  Returns the .obj path."""
    os.makedirs(directory, exist_ok=True)
    obj_path = os.path.join(directory, name + ".obj")
    with open(os.path.join(directory, name + ".mtl"), "w") as f:
        for i, m in enumerate(scene.materials):
            f.write("newmtl m%d\n" % i)
            f.write("Kd %.9g %.9g %.9g\n" % tuple(m["base_color"]))
            f.write("Ks %.9g %.9g %.9g\n" % tuple(m["specular_color"]))
            if any(m["emission_color"] > 0):
                f.write("Ke %.9g %.9g %.9g\n" % tuple(m["emission_color"]))
            f.write("Pr %.9g\nPm %.9g\n" % (m["specular_roughness"], m["metalness"]))
            if m["coat"] > 0:
                f.write("Pc %.9g\nPcr %.9g\n" % (m["coat"], m["coat_roughness"]))
            f.write("d %.9g\n" % (1.0 - m["transmission"]))
            f.write("Tf %.9g %.9g %.9g\n" % tuple(m["transmission_color"]))
            f.write("diffuse %.9g\ndiffuse_roughness %.9g\n" % (m["diffuse"], m["diffuse_roughness"]))
            f.write("sheen %.9g\nsheen_color %.9g %.9g %.9g\nsheen_roughness %.9g\n" %
                    (m["sheen"], *m["sheen_color"], m["sheen_roughness"]))
            f.write("subsurface %.9g\nsubsurface_color %.9g %.9g %.9g\nthin_walled %.9g\n" %
                    (m["subsurface"], *m["subsurface_color"], m["thin_walled"]))
    with open(obj_path, "w") as f:
        f.write("mtllib %s.mtl\n" % name)
        f.write("".join("v %.9g %.9g %.9g\n" % tuple(v) for v in scene.vertices))
        if with_attributes:
            f.write("".join("vn %.9g %.9g %.9g\n" % tuple(v) for v in scene.normals))
            f.write("".join("vt %.9g %.9g\n" % tuple(v) for v in scene.texcoords))
        for s, (off, cnt) in enumerate(zip(scene.submesh_offsets, scene.submesh_n_faces)):
            f.write("o shape%d\n" % s)
            cur = -1
            lines = []
            for fi in range(int(off), int(off + cnt)):
                mid = int(scene.material_ids[fi])
                if mid != cur:
                    lines.append("usemtl m%d\n" % mid)
                    cur = mid
                a, b, c = (int(v) + 1 for v in scene.indices[fi])
                if with_attributes:
                    lines.append("f %d/%d/%d %d/%d/%d %d/%d/%d\n" % (a, a, a, b, b, b, c, c, c))
                else:
                    lines.append("f %d %d %d\n" % (a, b, c))
            f.write("".join(lines))
    return obj_path


def write_gltf(scene: SceneArrays, directory, name="scene", embed=False, node_transforms=None, parents=None,
               animations=None, image_files=None, clearcoat=False):
    """Serialises `scene` as name.gltf (+ name.bin unless embed=True, which uses a base64 data URI).

    One mesh + one node per sub-mesh (16-bit indices, so every sub-mesh may reference at most
    65536 distinct vertices); node i gets node_transforms[i] = {"translation": .., "rotation":
    (x, y, z, w), "scale": ..} or {"matrix": 16 column-major floats}; parents[i] = parent node of
    node i (roots otherwise); animations = [{"node": i, "translation": (times, values), "rotation":
    (times, xyzw), "scale": (times, values)}]; image_files = file names (relative to `directory`)
    of the images behind texture ids 0..n-1 (the caller writes the files).  Returns the path."""
    import base64
    import json
    os.makedirs(directory, exist_ok=True)
    blob = bytearray()
    views, accessors = [], []

    def add(data, comp, typ, minmax=False):
        while len(blob) % 4:
            blob.append(0)
        a = np.ascontiguousarray(data)
        views.append({"buffer": 0, "byteOffset": len(blob), "byteLength": a.nbytes})
        blob.extend(a.tobytes())
        acc = {"bufferView": len(views) - 1, "componentType": comp, "count": int(a.shape[0]), "type": typ}
        if minmax:
            acc["min"] = [float(v) for v in np.atleast_1d(a.min(axis=0))]
            acc["max"] = [float(v) for v in np.atleast_1d(a.max(axis=0))]
        accessors.append(acc)
        return len(accessors) - 1

    meshes, nodes = [], []
    n_sub = len(scene.submesh_offsets)
    for s in range(n_sub):
        off, cnt = int(scene.submesh_offsets[s]), int(scene.submesh_n_faces[s])
        prims = []
        faces = np.arange(off, off + cnt)
        mids = scene.material_ids[faces]
        for mid in np.unique(mids):          # one primitive per material
            f = faces[mids == mid]
            idx = scene.indices[f].reshape(-1)
            used, local = np.unique(idx, return_inverse=True)
            assert len(used) <= 65536, "sub-mesh too large for 16-bit indices"
            uv = scene.texcoords[used].astype(np.float32).copy()
            uv[:, 1] = 1.0 - uv[:, 1]        # the loader stores (u, 1 - v)
            prims.append({"attributes": {"POSITION": add(scene.vertices[used].astype(np.float32), 5126, "VEC3", True),
                                         "NORMAL": add(scene.normals[used].astype(np.float32), 5126, "VEC3"),
                                         "TEXCOORD_0": add(uv, 5126, "VEC2")},
                          "indices": add(local.astype(np.uint16), 5123, "SCALAR"), "material": int(mid)})
        meshes.append({"primitives": prims})
        node = {"mesh": s, "name": "node%d" % s}
        if node_transforms and s < len(node_transforms) and node_transforms[s]:
            for k, v in node_transforms[s].items():
                node[k] = [float(x) for x in v]
        nodes.append(node)
    roots = list(range(n_sub))
    if parents:
        for child, parent in parents.items():
            nodes[parent].setdefault("children", []).append(int(child))
            roots.remove(child)

    def tex(i):
        return {"index": int(i)}

    materials = []
    for m in scene.materials:
        pbr = {"baseColorFactor": [float(v) for v in m["base_color"]] + [1.0],
               "roughnessFactor": float(m["specular_roughness"]), "metallicFactor": float(m["metalness"])}
        if m["base_color_texture_id"] >= 0:
            pbr["baseColorTexture"] = tex(m["base_color_texture_id"])
        if m["metallic_roughness_texture_id"] >= 0:
            pbr["metallicRoughnessTexture"] = tex(m["metallic_roughness_texture_id"])
        jm = {"pbrMetallicRoughness": pbr, "emissiveFactor": [float(v) for v in m["emission_color"]]}
        if m["emission_texture_id"] >= 0:
            jm["emissiveTexture"] = tex(m["emission_texture_id"])
        if m["normalmap_texture_id"] >= 0:
            jm["normalTexture"] = tex(m["normalmap_texture_id"])
        if clearcoat:
            jm["extensions"] = {"KHR_materials_clearcoat": {"clearcoatFactor": float(m["coat"]),
                                                            "clearcoatRoughnessFactor": float(m["coat_roughness"])}}
        materials.append(jm)

    janims = []
    for a in animations or []:
        samplers, channels = [], []
        for path, typ in (("translation", "VEC3"), ("rotation", "VEC4"), ("scale", "VEC3")):
            if path in a:
                times, values = a[path]
                samplers.append({"input": add(np.asarray(times, np.float32), 5126, "SCALAR", True),
                                 "output": add(np.asarray(values, np.float32), 5126, typ), "interpolation": "LINEAR"})
                channels.append({"sampler": len(samplers) - 1, "target": {"node": int(a["node"]), "path": path}})
        janims.append({"samplers": samplers, "channels": channels})

    doc = {"asset": {"version": "2.0", "generator": "fredholm_b200.scenes.write_gltf"}, "scene": 0,
           "scenes": [{"nodes": roots}], "nodes": nodes, "meshes": meshes, "materials": materials,
           "accessors": accessors, "bufferViews": views}
    if clearcoat:
        doc["extensionsUsed"] = ["KHR_materials_clearcoat"]
    if janims:
        doc["animations"] = janims
    if image_files:
        doc["images"] = [{"uri": f} for f in image_files]
        doc["textures"] = [{"source": i} for i in range(len(image_files))]
    if embed:
        doc["buffers"] = [{"byteLength": len(blob),
                           "uri": "data:application/octet-stream;base64," + base64.b64encode(bytes(blob)).decode()}]
    else:
        with open(os.path.join(directory, name + ".bin"), "wb") as f:
            f.write(bytes(blob))
        doc["buffers"] = [{"byteLength": len(blob), "uri": name + ".bin"}]
    path = os.path.join(directory, name + ".gltf")
    with open(path, "w") as f:
        json.dump(doc, f, indent=1)
    return path
