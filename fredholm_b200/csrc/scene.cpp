// Scene ingestion: Wavefront .obj/.mtl parser filling the flat arrays of
// fredholm::Scene.
//
// Behavioural spec = what the reference obtains from tinyobjloader (pinned at
// c44dde5.. in the reference's .SUBMODULES.json) followed by Scene::load_obj
// (fredholm/src/scene.cpp:119-443):
//   * shapes start at `o` / `g` lines; a shape without faces is dropped; the
//     active material survives shape boundaries; every shape becomes one sub-mesh
//     with an identity transform and instance id 0 for all of its faces;
//   * polygons are triangulated (quads along the shorter diagonal, like
//     tinyobjloader; larger polygons as a fan -- tinyobjloader ear-clips those);
//   * vertices are de-duplicated over the whole file on exact (position, normal,
//     texcoord) equality, in order of first use; faces without normals get the
//     face normal, faces without texcoords get (0,0),(1,0),(0,1);
//   * MTL keys -> Material exactly as scene.cpp:170-312 maps them, including the
//     custom keys (diffuse, diffuse_roughness, sheen, sheen_color, ...) and the
//     `Pcr` quirk (coat_roughness takes the clearcoat THICKNESS, scene.cpp:240-242).
// Decimal numbers are converted with the same digit-accumulation scheme
// tinyobjloader uses (not strtod), so vertex data is bit-identical to what the
// reference loader produces from the same file.
#include "fredholm/scene.h"

#include "mesh_prep.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

namespace fredholm
{

// ---- small linear algebra ----------------------------------------------------------
mat4 operator*(const mat4& a, const mat4& b)
{
  mat4 r;
  for (int c = 0; c < 4; ++c)
    for (int row = 0; row < 4; ++row) {
      float s = 0.0f;
      for (int k = 0; k < 4; ++k) s += a[k][row] * b[c][k];
      r[c][row] = s;
    }
  return r;
}

mat4 inverse(const mat4& m)
{
  // cofactor expansion, same formulation glm uses for mat4
  const float c00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
  const float c02 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
  const float c03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
  const float c04 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
  const float c06 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
  const float c07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
  const float c08 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
  const float c10 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
  const float c11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
  const float c12 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
  const float c14 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
  const float c15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
  const float c16 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
  const float c18 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
  const float c19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
  const float c20 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
  const float c22 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
  const float c23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];

  const float f0[4] = {c00, c00, c02, c03};
  const float f1[4] = {c04, c04, c06, c07};
  const float f2[4] = {c08, c08, c10, c11};
  const float f3[4] = {c12, c12, c14, c15};
  const float f4[4] = {c16, c16, c18, c19};
  const float f5[4] = {c20, c20, c22, c23};
  const float v0[4] = {m[1][0], m[0][0], m[0][0], m[0][0]};
  const float v1[4] = {m[1][1], m[0][1], m[0][1], m[0][1]};
  const float v2[4] = {m[1][2], m[0][2], m[0][2], m[0][2]};
  const float v3[4] = {m[1][3], m[0][3], m[0][3], m[0][3]};
  const float sa[4] = {+1, -1, +1, -1}, sb[4] = {-1, +1, -1, +1};
  mat4 inv;
  for (int i = 0; i < 4; ++i) {
    inv[0][i] = (v1[i] * f0[i] - v2[i] * f1[i] + v3[i] * f2[i]) * sa[i];
    inv[1][i] = (v0[i] * f0[i] - v2[i] * f3[i] + v3[i] * f4[i]) * sb[i];
    inv[2][i] = (v0[i] * f1[i] - v1[i] * f3[i] + v3[i] * f5[i]) * sa[i];
    inv[3][i] = (v0[i] * f2[i] - v1[i] * f4[i] + v2[i] * f5[i]) * sb[i];
  }
  const float det = m[0][0] * inv[0][0] + m[0][1] * inv[1][0] + m[0][2] * inv[2][0] + m[0][3] * inv[3][0];
  const float one_over_det = 1.0f / det;
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) inv[c][r] *= one_over_det;
  return inv;
}

// ---- text scanning -------------------------------------------------------------------
namespace
{

inline bool is_space(char c) { return c == ' ' || c == '\t'; }
inline bool is_digit(char c) { return c >= '0' && c <= '9'; }
inline bool is_eol(char c) { return c == '\r' || c == '\n' || c == '\0'; }

// Decimal -> double by digit accumulation (integer digits: m = 10 m + d; fraction
// digit k: m += d * 10^-k; exponent e: ldexp(m * 5^e, e)), i.e. the conversion the
// reference's loader library performs; then narrowed to float by the caller.
bool parse_decimal(const char* s, const char* end, double& result)
{
  if (s >= end) return false;
  double mant = 0.0;
  int exponent = 0;
  char sign = '+', exp_sign = '+';
  const char* c = s;
  int read = 0;
  bool more = false, leading_dot = false;
  if (*c == '+' || *c == '-') {
    sign = *c;
    c++;
    if (c != end && *c == '.') leading_dot = true;
  } else if (is_digit(*c)) {
  } else if (*c == '.') {
    leading_dot = true;
  } else {
    return false;
  }
  more = (c != end);
  if (!leading_dot) {
    while (more && is_digit(*c)) {
      mant *= 10;
      mant += static_cast<int>(*c - '0');
      c++;
      read++;
      more = (c != end);
    }
    if (read == 0) return false;
  }
  bool have_exp = false;
  if (more) {
    if (*c == '.') {
      c++;
      read = 1;
      more = (c != end);
      static const double lut[] = {1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001};
      while (more && is_digit(*c)) {
        mant += static_cast<int>(*c - '0') * (read < 8 ? lut[read] : std::pow(10.0, -read));
        read++;
        c++;
        more = (c != end);
      }
      have_exp = more && (*c == 'e' || *c == 'E');
    } else if (*c == 'e' || *c == 'E') {
      have_exp = true;
    }
  }
  if (have_exp) {
    c++;
    more = (c != end);
    if (more && (*c == '+' || *c == '-')) {
      exp_sign = *c;
      c++;
    } else if (more && is_digit(*c)) {
    } else {
      return false;
    }
    read = 0;
    more = (c != end);
    while (more && is_digit(*c)) {
      if (exponent > 214748364) return false;
      exponent = exponent * 10 + static_cast<int>(*c - '0');
      c++;
      read++;
      more = (c != end);
    }
    exponent *= (exp_sign == '+' ? 1 : -1);
    if (read == 0) return false;
  }
  result = (sign == '+' ? 1 : -1) * (exponent ? std::ldexp(mant * std::pow(5.0, exponent), exponent) : mant);
  return true;
}

struct Cursor {
  const char* p;
  void skip_space()
  {
    while (is_space(*p)) p++;
  }
  // next whitespace-delimited token on this line; [b, e)
  bool token(const char*& b, const char*& e)
  {
    skip_space();
    if (is_eol(*p)) return false;
    b = p;
    while (!is_space(*p) && !is_eol(*p)) p++;
    e = p;
    return true;
  }
  float real(double fallback = 0.0)
  {
    const char *b, *e;
    double v = fallback;
    if (token(b, e)) {
      double t;
      if (parse_decimal(b, e, t)) v = t;
    }
    return static_cast<float>(v);
  }
  std::string rest_of_line()
  {
    skip_space();
    const char* b = p;
    while (!is_eol(*p)) p++;
    const char* e = p;
    while (e > b && is_space(e[-1])) e--;
    return std::string(b, e);
  }
};

std::string read_file(const std::filesystem::path& path)
{
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f.is_open()) throw std::runtime_error("failed to load " + path.generic_string());
  const std::streamsize n = f.tellg();
  std::string s(static_cast<size_t>(n), '\0');
  f.seekg(0);
  f.read(s.data(), n);
  return s;
}

// ---- MTL ---------------------------------------------------------------------------------
struct MtlRecord {
  std::string name;
  float Kd[3] = {0, 0, 0}, Ks[3] = {0, 0, 0}, Ke[3] = {0, 0, 0}, Tf[3] = {0, 0, 0};
  float dissolve = 1.0f, roughness = 0.0f, metallic = 0.0f, clearcoat_thickness = 0.0f,
        clearcoat_roughness = 0.0f;
  std::string map_Kd, map_Ks, map_Pr, map_Pm, map_bump, map_norm, map_d;
  std::map<std::string, std::string> extra;  // first occurrence wins
};

// texture statements may carry options (-bm 1.0, -o u v w, ...); the file name is
// the last token
std::string texture_name(Cursor& cur)
{
  const std::string rest = cur.rest_of_line();
  std::vector<std::string> toks;
  std::stringstream ss(rest);
  std::string t;
  while (ss >> t) toks.push_back(t);
  if (toks.empty()) return "";
  // options: -name followed by 1..3 values; everything after the last option run is the name
  size_t i = 0;
  while (i < toks.size() && toks[i].size() > 1 && toks[i][0] == '-' && !is_digit(toks[i][1]) && toks[i][1] != '.') {
    const std::string& o = toks[i];
    size_t nargs = 1;
    if (o == "-o" || o == "-s" || o == "-t") nargs = 3;
    if (o == "-mm") nargs = 2;
    i += 1 + nargs;
  }
  std::string name;
  for (size_t k = std::min(i, toks.size() - 1); k < toks.size(); ++k) {
    if (!name.empty()) name += " ";
    name += toks[k];
  }
  return name;
}

void parse_mtl(const std::string& text, std::vector<MtlRecord>& out, std::map<std::string, int>& by_name)
{
  MtlRecord cur_m;
  bool open = false, has_d = false;
  auto flush = [&]() {
    if (!open) return;
    by_name.insert({cur_m.name, (int)out.size()});  // duplicate names: the first one is kept
    out.push_back(cur_m);
  };
  const char* p = text.c_str();
  while (*p) {
    Cursor cur{p};
    cur.skip_space();
    const char* line = cur.p;
    // find end of line for the next iteration
    const char* eol = line;
    while (*eol && *eol != '\n') eol++;
    p = *eol ? eol + 1 : eol;
    if (is_eol(*line) || *line == '#') continue;
    const char *kb, *ke;
    if (!cur.token(kb, ke)) continue;
    const std::string key(kb, ke);
    auto rgb = [&](float* dst) {
      dst[0] = cur.real();
      dst[1] = cur.real();
      dst[2] = cur.real();
    };
    if (key == "newmtl") {
      flush();
      cur_m = MtlRecord();
      cur_m.name = cur.rest_of_line();
      open = true;
      has_d = false;
    } else if (!open) {
      continue;
    } else if (key == "Kd") {
      rgb(cur_m.Kd);
    } else if (key == "Ks") {
      rgb(cur_m.Ks);
    } else if (key == "Ke") {
      rgb(cur_m.Ke);
    } else if (key == "Kt" || key == "Tf") {
      rgb(cur_m.Tf);
    } else if (key == "d") {
      cur_m.dissolve = cur.real();
      has_d = true;
    } else if (key == "Tr") {
      if (!has_d) cur_m.dissolve = 1.0f - cur.real();
    } else if (key == "Pr") {
      cur_m.roughness = cur.real();
    } else if (key == "Pm") {
      cur_m.metallic = cur.real();
    } else if (key == "Pc") {
      cur_m.clearcoat_thickness = cur.real();
    } else if (key == "Pcr") {
      cur_m.clearcoat_roughness = cur.real();
    } else if (key == "map_Kd") {
      cur_m.map_Kd = texture_name(cur);
    } else if (key == "map_Ks") {
      cur_m.map_Ks = texture_name(cur);
    } else if (key == "map_Pr") {
      cur_m.map_Pr = texture_name(cur);
    } else if (key == "map_Pm") {
      cur_m.map_Pm = texture_name(cur);
    } else if (key == "map_bump" || key == "map_Bump" || key == "bump") {
      cur_m.map_bump = texture_name(cur);
    } else if (key == "norm") {
      cur_m.map_norm = texture_name(cur);
    } else if (key == "map_d") {
      cur_m.map_d = texture_name(cur);
    } else if (key == "Ka" || key == "Ni" || key == "Ns" || key == "illum" || key == "Ps" || key == "aniso" ||
               key == "anisor" || key == "map_Ka" || key == "map_Ns" || key == "map_Ke" || key == "map_Ps" ||
               key == "disp" || key == "refl") {
      // known to the .mtl format but not consumed by the renderer
    } else {
      // custom parameter: value = rest of the line; the first occurrence is kept
      const std::string value = cur.rest_of_line();
      cur_m.extra.insert({key, value});
    }
  }
  flush();
}

float3 parse_float3_string(const std::string& s)
{
  std::vector<std::string> toks;
  std::stringstream ss(s);
  std::string t;
  while (std::getline(ss, t, ' '))
    if (!t.empty()) toks.push_back(t);
  if (toks.size() != 3) throw std::runtime_error("invalid vec3 string in .mtl: '" + s + "'");
  return make_float3(std::stof(toks[0]), std::stof(toks[1]), std::stof(toks[2]));
}

// ---- vertex de-duplication -------------------------------------------------------------
struct VKey {
  float v[8];
  bool operator==(const VKey& o) const
  {
    for (int i = 0; i < 8; ++i)
      if (!(v[i] == o.v[i])) return false;
    return true;
  }
};
struct VKeyHash {
  size_t operator()(const VKey& k) const
  {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 8; ++i) {
      float f = k.v[i] == 0.0f ? 0.0f : k.v[i];  // -0 == +0 must hash alike
      uint32_t b;
      std::memcpy(&b, &f, 4);
      h = (h ^ b) * 1099511628211ull;
    }
    return static_cast<size_t>(h);
  }
};

struct ObjIndex {
  int v = -1, vt = -1, vn = -1;
};

// "a", "a/b", "a//c", "a/b/c"; indices are 1-based, negative = relative to the end
bool parse_face_vertex(const char* b, const char* e, int nv, int nvt, int nvn, ObjIndex& out)
{
  auto fix = [](int idx, int n, int& dst) {
    if (idx > 0) dst = idx - 1;
    else if (idx < 0) dst = n + idx;
    else return false;
    return true;
  };
  auto read_int = [&](const char*& p, int& val) {
    bool neg = false;
    if (p < e && (*p == '-' || *p == '+')) {
      neg = *p == '-';
      p++;
    }
    if (p >= e || !is_digit(*p)) return false;
    long v = 0;
    while (p < e && is_digit(*p)) v = v * 10 + (*p++ - '0');
    val = neg ? -(int)v : (int)v;
    return true;
  };
  const char* p = b;
  int i;
  if (!read_int(p, i) || !fix(i, nv, out.v)) return false;
  if (p >= e || *p != '/') return true;
  p++;
  if (p < e && *p == '/') {
    p++;
    if (read_int(p, i)) fix(i, nvn, out.vn);
    return true;
  }
  if (read_int(p, i)) fix(i, nvt, out.vt);
  if (p < e && *p == '/') {
    p++;
    if (read_int(p, i)) fix(i, nvn, out.vn);
  }
  return true;
}

}  // namespace

// ---- Scene ------------------------------------------------------------------------------
bool Scene::is_valid() const
{
  return m_submesh_offsets.size() > 0 && m_vertices.size() > 0 && m_indices.size() > 0;
}

// Everything the kernels index with scene data is checked here once, on the host, so that a malformed scene
// (hand-filled arrays through fr_set_scene_arrays / set_scene, a glTF primitive without NORMAL or TEXCOORD_0)
// throws "invalid scene: ..." instead of reading outside a device buffer.
void Scene::validate() const
{
  const Scene& s = *this;
  auto fail = [](const std::string& what) { throw std::runtime_error("invalid scene: " + what); };
  const size_t nv = s.m_vertices.size(), nf = s.m_indices.size(), nsm = s.m_submesh_offsets.size();
  if (nv == 0 || nf == 0) fail("no geometry");
  if (s.m_normals.size() != nv) fail("normals.size() != vertices.size()");
  if (s.m_texcoords.size() != nv) fail("texcoords.size() != vertices.size()");
  if (s.m_material_ids.size() != nf) fail("material_ids.size() != number of faces");
  if (s.m_instance_ids.size() != nf) fail("instance_ids.size() != number of faces");
  if (s.m_submesh_n_faces.size() != nsm) fail("submesh_offsets / submesh_n_faces differ in length");
  if (s.m_transforms.size() != nsm) fail("transforms.size() != number of sub-meshes");
  if (nsm == 0) fail("no sub-mesh");
  for (size_t f = 0; f < nf; ++f) {
    const uint3 i = s.m_indices[f];
    if (i.x >= nv || i.y >= nv || i.z >= nv) fail("vertex index out of range in face " + std::to_string(f));
    if (s.m_material_ids[f] >= s.m_materials.size()) fail("face without a valid material");
    if (s.m_instance_ids[f] >= nsm) fail("instance id out of range in face " + std::to_string(f));
  }
  // the sub-meshes must tile the face range: every face belongs to exactly one
  std::vector<uint8_t> covered(nf, 0);
  for (size_t sm = 0; sm < nsm; ++sm) {
    const uint64_t o = s.m_submesh_offsets[sm], n = s.m_submesh_n_faces[sm];
    if (o + n > nf) fail("sub-mesh " + std::to_string(sm) + " exceeds the face array");
    for (uint64_t f = o; f < o + n; ++f) {
      if (covered[f]) fail("sub-meshes overlap at face " + std::to_string(f));
      covered[f] = 1;
    }
  }
  for (size_t f = 0; f < nf; ++f)
    if (!covered[f]) fail("face " + std::to_string(f) + " belongs to no sub-mesh");
  const int nt = (int)s.m_textures.size();
  for (size_t m = 0; m < s.m_materials.size(); ++m) {
    const Material& mt = s.m_materials[m];
    const int ids[] = {mt.base_color_texture_id, mt.specular_color_texture_id, mt.specular_roughness_texture_id,
                       mt.metalness_texture_id, mt.metallic_roughness_texture_id, mt.coat_texture_id,
                       mt.coat_roughness_texture_id, mt.emission_texture_id, mt.heightmap_texture_id,
                       mt.normalmap_texture_id, mt.alpha_texture_id};
    for (int id : ids)
      if (id < -1 || id >= nt) fail("texture id " + std::to_string(id) + " out of range in material " + std::to_string(m));
  }
  for (size_t t = 0; t < s.m_textures.size(); ++t) {
    const Texture& tx = s.m_textures[t];
    if (tx.m_width == 0 || tx.m_height == 0 || tx.m_data.size() != (size_t)tx.m_width * tx.m_height)
      fail("texture " + std::to_string(t) + " has no texels or a wrong size");
  }
}

void Scene::clear() { *this = Scene(); }

void Scene::load_model(const std::filesystem::path& filepath, bool do_clear)
{
  if (do_clear) clear();
  if (filepath.extension() == ".obj") {
    load_obj(filepath);
  } else if (filepath.extension() == ".gltf") {
    load_gltf(filepath);
  } else {
    throw std::runtime_error("failed to load " + filepath.generic_string() + "\n" + "reason: invalid extension");
  }
}

void Scene::load_obj(const std::filesystem::path& filepath)
{
  const std::string text = read_file(filepath);

  std::vector<float> pos, nrm, tex;  // raw attribute pools of the file
  std::vector<MtlRecord> mtl;
  std::map<std::string, int> mtl_by_name;

  struct PendingFace {
    std::vector<ObjIndex> corners;
  };
  struct Shape {
    std::vector<ObjIndex> corners;  // 3 per triangle
    std::vector<int> material_ids;  // 1 per triangle
  };
  std::vector<Shape> shapes;
  Shape shape;
  std::vector<PendingFace> group;  // faces awaiting triangulation (same material)
  int material = -1;

  // triangulation happens when a face group is closed, against the vertex pool
  // parsed so far (the quad rule needs positions)
  auto flush_group = [&]() {
    for (const PendingFace& f : group) {
      const size_t n = f.corners.size();
      if (n < 3) continue;
      auto emit = [&](int a, int b, int c) {
        shape.corners.push_back(f.corners[a]);
        shape.corners.push_back(f.corners[b]);
        shape.corners.push_back(f.corners[c]);
        shape.material_ids.push_back(material);
      };
      if (n == 3) {
        emit(0, 1, 2);
      } else if (n == 4) {
        const float* p0 = &pos[3 * (size_t)f.corners[0].v];
        const float* p1 = &pos[3 * (size_t)f.corners[1].v];
        const float* p2 = &pos[3 * (size_t)f.corners[2].v];
        const float* p3 = &pos[3 * (size_t)f.corners[3].v];
        const float e02[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
        const float e13[3] = {p3[0] - p1[0], p3[1] - p1[1], p3[2] - p1[2]};
        const float s02 = e02[0] * e02[0] + e02[1] * e02[1] + e02[2] * e02[2];
        const float s13 = e13[0] * e13[0] + e13[1] * e13[1] + e13[2] * e13[2];
        if (s02 < s13) {
          emit(0, 1, 2);
          emit(0, 2, 3);
        } else {
          emit(0, 1, 3);
          emit(1, 2, 3);
        }
      } else {
        for (size_t k = 1; k + 1 < n; ++k) emit(0, (int)k, (int)k + 1);
      }
    }
    group.clear();
  };
  auto close_shape = [&]() {
    flush_group();
    if (!shape.corners.empty()) shapes.push_back(std::move(shape));
    shape = Shape();
  };

  const char* p = text.c_str();
  while (*p) {
    Cursor cur{p};
    cur.skip_space();
    const char* line = cur.p;
    const char* eol = line;
    while (*eol && *eol != '\n') eol++;
    p = *eol ? eol + 1 : eol;
    if (is_eol(*line) || *line == '#') continue;

    if (line[0] == 'v' && is_space(line[1])) {
      cur.p = line + 2;
      pos.push_back(cur.real());
      pos.push_back(cur.real());
      pos.push_back(cur.real());
    } else if (line[0] == 'v' && line[1] == 'n' && is_space(line[2])) {
      cur.p = line + 3;
      nrm.push_back(cur.real());
      nrm.push_back(cur.real());
      nrm.push_back(cur.real());
    } else if (line[0] == 'v' && line[1] == 't' && is_space(line[2])) {
      cur.p = line + 3;
      tex.push_back(cur.real());
      tex.push_back(cur.real());
    } else if (line[0] == 'f' && is_space(line[1])) {
      cur.p = line + 2;
      PendingFace face;
      const char *b, *e;
      bool ok = true;
      while (cur.token(b, e)) {
        ObjIndex idx;
        if (!parse_face_vertex(b, e, (int)(pos.size() / 3), (int)(tex.size() / 2), (int)(nrm.size() / 3), idx)) {
          ok = false;
          break;
        }
        face.corners.push_back(idx);
      }
      if (!ok) throw std::runtime_error("failed to load " + filepath.generic_string() + ": bad face statement");
      group.push_back(std::move(face));
    } else if (std::strncmp(line, "usemtl", 6) == 0) {
      cur.p = line + 6;
      const std::string name = cur.rest_of_line();
      const auto it = mtl_by_name.find(name);
      const int id = it != mtl_by_name.end() ? it->second : -1;
      if (id != material) {
        flush_group();
        material = id;
      }
    } else if (std::strncmp(line, "mtllib", 6) == 0 && is_space(line[6])) {
      cur.p = line + 7;
      std::stringstream ss(cur.rest_of_line());
      std::string fname;
      while (ss >> fname) {
        const std::filesystem::path mp = filepath.parent_path() / fname;
        std::ifstream probe(mp);
        if (!probe.is_open()) continue;
        probe.close();
        parse_mtl(read_file(mp), mtl, mtl_by_name);
        break;  // the first file that opens is used
      }
    } else if ((line[0] == 'g' || line[0] == 'o') && is_space(line[1])) {
      close_shape();
    }
  }
  close_shape();

  // ---- materials + textures (scene.cpp:135-312) ----
  std::unordered_map<std::string, unsigned int> unique_textures;
  auto load_texture = [&](const std::string& name, TextureType type) -> int {
    const auto it = unique_textures.find(name);
    if (it != unique_textures.end()) return (int)it->second;
    const unsigned int id = (unsigned int)m_textures.size();
    unique_textures[name] = id;
    m_textures.push_back(Texture(filepath.parent_path() / name, type));
    return (int)id;
  };
  const size_t material_base = m_materials.size();
  (void)material_base;
  for (const MtlRecord& m : mtl) {
    Material mat;
    auto extra_f = [&](const char* key, float& dst) {
      const auto it = m.extra.find(key);
      if (it != m.extra.end()) dst = std::stof(it->second);
    };
    auto extra_f3 = [&](const char* key, float3& dst) {
      const auto it = m.extra.find(key);
      if (it != m.extra.end()) dst = parse_float3_string(it->second);
    };
    extra_f("diffuse", mat.diffuse);
    extra_f("diffuse_roughness", mat.diffuse_roughness);
    mat.base_color = make_float3(m.Kd[0], m.Kd[1], m.Kd[2]);
    if (!m.map_Kd.empty()) mat.base_color_texture_id = load_texture(m.map_Kd, TextureType::COLOR);
    mat.specular_color = make_float3(m.Ks[0], m.Ks[1], m.Ks[2]);
    if (!m.map_Ks.empty()) mat.specular_color_texture_id = load_texture(m.map_Ks, TextureType::COLOR);
    if (m.roughness > 0) mat.specular_roughness = m.roughness;
    if (!m.map_Pr.empty()) mat.specular_roughness_texture_id = load_texture(m.map_Pr, TextureType::NONCOLOR);
    mat.metalness = m.metallic;
    if (!m.map_Pm.empty()) mat.metalness_texture_id = load_texture(m.map_Pm, TextureType::NONCOLOR);
    if (m.clearcoat_thickness > 0) mat.coat = m.clearcoat_thickness;
    if (m.clearcoat_roughness > 0) mat.coat_roughness = m.clearcoat_thickness;  // sic (scene.cpp:240-242)
    mat.transmission = std::max(1.0f - m.dissolve, 0.0f);
    if (m.Tf[0] > 0 || m.Tf[1] > 0 || m.Tf[2] > 0) mat.transmission_color = make_float3(m.Tf[0], m.Tf[1], m.Tf[2]);
    extra_f("sheen", mat.sheen);
    extra_f3("sheen_color", mat.sheen_color);
    extra_f("sheen_roughness", mat.sheen_roughness);
    extra_f("subsurface", mat.subsurface);
    extra_f3("subsurface_color", mat.subsurface_color);
    extra_f("thin_walled", mat.thin_walled);
    if (m.Ke[0] > 0 || m.Ke[1] > 0 || m.Ke[2] > 0) {
      mat.emission = 1.0f;
      mat.emission_color = make_float3(m.Ke[0], m.Ke[1], m.Ke[2]);
    }
    if (!m.map_bump.empty()) mat.heightmap_texture_id = load_texture(m.map_bump, TextureType::NONCOLOR);
    if (!m.map_norm.empty()) mat.normalmap_texture_id = load_texture(m.map_norm, TextureType::NONCOLOR);
    if (!m.map_d.empty()) mat.alpha_texture_id = load_texture(m.map_d, TextureType::NONCOLOR);
    m_materials.push_back(mat);
  }

  // ---- geometry (scene.cpp:314-437) ----
  // Large meshes on a machine with a CUDA device: expansion, face normals and vertex de-duplication run as
  // kernels (mesh_prep.cu) and produce the same arrays bit for bit; FRD_GPU_MESH_PREP=0 / 1 forces the host loop
  // / the kernels.
  {
    size_t total_corners = 0;
    for (const Shape& s : shapes) total_corners += s.corners.size();
    const char* e = getenv("FRD_GPU_MESH_PREP");
    const bool forced_on = e && e[0] == '1', forced_off = e && e[0] == '0';
    const bool first_model = m_vertices.empty() && m_indices.empty();  // appending to a scene keeps the host loop
    if (!forced_off && first_model && total_corners > 0 && (forced_on || total_corners >= 300000) &&
        frd::mesh_prep_available()) {
      std::vector<frd::ObjCorner> corners;
      corners.reserve(total_corners);
      for (const Shape& s : shapes)
        for (const ObjIndex& c : s.corners) corners.push_back(frd::ObjCorner{c.v, c.vt, c.vn});
      frd::PreparedMesh pm;
      try {
        frd::prepare_mesh_gpu(pos, nrm, tex, corners, pm);
      } catch (const std::runtime_error& err) {
        throw std::runtime_error("failed to load " + filepath.generic_string() + ": " + err.what());
      }
      m_vertices = std::move(pm.vertices);
      m_normals = std::move(pm.normals);
      m_texcoords = std::move(pm.texcoords);
      m_indices = std::move(pm.indices);
      for (const Shape& s : shapes) {
        m_submesh_offsets.push_back((unsigned int)m_material_ids.size());
        for (int id : s.material_ids) m_material_ids.push_back((unsigned int)id);
        m_submesh_n_faces.push_back((unsigned int)s.material_ids.size());
        m_transforms.push_back(mat4::identity());
      }
      m_instance_ids.assign(m_indices.size(), 0u);  // .obj has no instancing (scene.cpp:425-427)
      return;
    }
  }
  std::vector<VKey> unique;
  std::unordered_map<VKey, uint32_t, VKeyHash> lookup;
  lookup.reserve(pos.size() / 3 + 16);
  for (const Shape& s : shapes) {
    const size_t first_face = m_indices.size();
    m_submesh_offsets.push_back((unsigned int)first_face);
    const size_t n_tris = s.corners.size() / 3;
    for (size_t f = 0; f < n_tris; ++f) {
      const ObjIndex* c = &s.corners[3 * f];
      float P[3][3], N[3][3], T[3][2];
      int n_normals = 0, n_tex = 0;
      for (int k = 0; k < 3; ++k) {
        if (c[k].v < 0 || 3 * (size_t)c[k].v + 2 >= pos.size())
          throw std::runtime_error("failed to load " + filepath.generic_string() + ": vertex index out of range");
        for (int a = 0; a < 3; ++a) P[k][a] = pos[3 * (size_t)c[k].v + a];
        if (c[k].vn >= 0 && 3 * (size_t)c[k].vn + 2 < nrm.size()) {
          for (int a = 0; a < 3; ++a) N[n_normals][a] = nrm[3 * (size_t)c[k].vn + a];
          n_normals++;
        }
        if (c[k].vt >= 0 && 2 * (size_t)c[k].vt + 1 < tex.size()) {
          for (int a = 0; a < 2; ++a) T[n_tex][a] = tex[2 * (size_t)c[k].vt + a];
          n_tex++;
        }
      }
      if (n_normals < 3) {
        // face normal from normalized edges (scene.cpp:363-372)
        auto nrmz = [](const vec3& v) {
          const float inv = 1.0f / std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
          return vec3(v.x * inv, v.y * inv, v.z * inv);
        };
        const vec3 e1 = nrmz(vec3(P[1][0] - P[0][0], P[1][1] - P[0][1], P[1][2] - P[0][2]));
        const vec3 e2 = nrmz(vec3(P[2][0] - P[0][0], P[2][1] - P[0][1], P[2][2] - P[0][2]));
        const vec3 n = nrmz(cross(e1, e2));
        for (int k = 0; k < 3; ++k) {
          N[k][0] = n.x;
          N[k][1] = n.y;
          N[k][2] = n.z;
        }
      }
      if (n_tex < 3) {
        T[0][0] = 0, T[0][1] = 0;
        T[1][0] = 1, T[1][1] = 0;
        T[2][0] = 0, T[2][1] = 1;
      }
      uint32_t vid[3];
      for (int k = 0; k < 3; ++k) {
        VKey key{{P[k][0], P[k][1], P[k][2], N[k][0], N[k][1], N[k][2], T[k][0], T[k][1]}};
        const auto it = lookup.find(key);
        if (it == lookup.end()) {
          vid[k] = (uint32_t)unique.size();
          lookup.emplace(key, vid[k]);
          unique.push_back(key);
        } else {
          vid[k] = it->second;
        }
        // quirk (scene.cpp:380-387): a vertex with a NaN component (the face normal of a degenerate triangle) equals
        // nothing, not even itself, so the reference's `indices.push_back(unique_vertices[vertex])` inserts a fresh
        // map entry with value 0 -- the vertex is appended, but the corner refers to vertex 0
        for (int a = 0; a < 8; ++a)
          if (key.v[a] != key.v[a]) vid[k] = 0;
      }
      m_indices.push_back(make_uint3(vid[0], vid[1], vid[2]));
      m_material_ids.push_back((unsigned int)s.material_ids[f]);
      m_instance_ids.push_back(0);  // .obj has no instancing (scene.cpp:425-427)
    }
    m_submesh_n_faces.push_back((unsigned int)(m_indices.size() - first_face));
    m_transforms.push_back(mat4::identity());
  }
  for (const VKey& k : unique) {
    m_vertices.push_back(make_float3(k.v[0], k.v[1], k.v[2]));
    m_normals.push_back(make_float3(k.v[3], k.v[4], k.v[5]));
    m_texcoords.push_back(make_float2(k.v[6], k.v[7]));
  }
}

}  // namespace fredholm
