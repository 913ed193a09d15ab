// File-backed parts of the scene interface: image textures, glTF 2.0 scenes and their
// TRS animations.
//
// Behavioural spec = the reference's Scene::load_gltf / load_gltf_node / update_transform /
// update_animation (fredholm/src/scene.cpp:445-898) on top of tinygltf's LoadASCIIFromFile
// (externals/tinygltf, pinned in .SUBMODULES.json) and glm for the matrix algebra.  This
// file carries its own JSON reader, base64 / data-URI decoding and accessor resolution, and
// reproduces the reference's conventions and quirks:
//   * only scenes[0] is instantiated; every node that references a mesh gets its OWN copy
//     of the mesh data (no sharing) and becomes one sub-mesh = one instance;
//   * indices must be unsigned short, positions / normals float3, texcoords float2, all
//     tightly packed; texcoords are stored as (u, 1 - v); a primitive without a material
//     gets id 0xffffffff;
//   * material: baseColorFactor, baseColorTexture, roughnessFactor, metallicFactor,
//     metallicRoughnessTexture, KHR_materials_clearcoat factors, emissiveFactor (always
//     present in tinygltf, so emission = 1), emissiveTexture, normalTexture; the clearcoat
//     TEXTURE entries are objects that the reference reads as a number, which yields 0;
//   * every texture is loaded as NONCOLOR (no sRGB decode) from images[source].uri;
//   * node matrix = T * R * S (glm::translate, mat4_cast, glm::scale), replaced by `matrix`
//     when that is present; world transform = product down the hierarchy;
//   * an animation drives the node of its FIRST channel, which must be a root node of the
//     scene (the reference's recursive search drops results found in children); keyframes
//     are interpolated with weight (t - t0) -- not normalised by the key spacing -- with
//     t = fmod(time, last key); rotations interpolate with glm's quaternion mix.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <stdexcept>

#include "fredholm/scene.h"
#include "image_codec.h"

namespace fredholm
{

// ---- textures ------------------------------------------------------------------------------
Texture::Texture(const std::filesystem::path& filepath, const TextureType& texture_type)
    : m_texture_type(texture_type)
{
  // stbi_set_flip_vertically_on_load(true): row 0 of m_data is the bottom row of the file
  const codec::Image8 img = codec::load_image8(filepath.generic_string());
  m_width = (uint32_t)img.width;
  m_height = (uint32_t)img.height;
  m_data.resize((size_t)m_width * m_height);
  for (uint32_t j = 0; j < m_height; ++j) {
    const uint8_t* src = img.rgba.data() + (size_t)(m_height - 1 - j) * m_width * 4;
    for (uint32_t i = 0; i < m_width; ++i)
      m_data[(size_t)j * m_width + i] = make_uchar4(src[4 * i], src[4 * i + 1], src[4 * i + 2], src[4 * i + 3]);
  }
}

FloatTexture::FloatTexture(const std::filesystem::path& filepath)
{
  const codec::ImageF img = codec::load_imagef(filepath.generic_string());
  m_width = (uint32_t)img.width;
  m_height = (uint32_t)img.height;
  m_data.resize((size_t)m_width * m_height);
  for (size_t i = 0; i < m_data.size(); ++i)
    m_data[i] = make_float4(img.rgba[4 * i], img.rgba[4 * i + 1], img.rgba[4 * i + 2], img.rgba[4 * i + 3]);
}

namespace
{

// ---- JSON ------------------------------------------------------------------------------------
struct Json {
  enum Type { Null, Bool, Number, String, Array, Object } type = Null;
  bool b = false;
  double num = 0.0;
  std::string str;
  std::vector<Json> arr;
  std::vector<std::pair<std::string, Json>> obj;

  const Json* find(const char* key) const
  {
    if (type != Object) return nullptr;
    for (const auto& kv : obj)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
  bool has(const char* key) const { return find(key) != nullptr; }
  double number(const char* key, double fallback) const
  {
    const Json* v = find(key);
    return (v && v->type == Number) ? v->num : fallback;
  }
  int integer(const char* key, int fallback) const
  {
    const Json* v = find(key);
    return (v && v->type == Number) ? (int)v->num : fallback;
  }
  std::string string(const char* key) const
  {
    const Json* v = find(key);
    return (v && v->type == String) ? v->str : std::string();
  }
  const std::vector<Json>& array(const char* key) const
  {
    static const std::vector<Json> empty;
    const Json* v = find(key);
    return (v && v->type == Array) ? v->arr : empty;
  }
};

struct JsonParser {
  const char* p;
  const char* end;
  [[noreturn]] void error(const char* what) const { throw std::runtime_error(std::string("JSON: ") + what); }
  void ws()
  {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
  }
  static void utf8(std::string& s, uint32_t cp)
  {
    if (cp < 0x80) {
      s.push_back((char)cp);
    } else if (cp < 0x800) {
      s.push_back((char)(0xC0 | (cp >> 6)));
      s.push_back((char)(0x80 | (cp & 63)));
    } else if (cp < 0x10000) {
      s.push_back((char)(0xE0 | (cp >> 12)));
      s.push_back((char)(0x80 | ((cp >> 6) & 63)));
      s.push_back((char)(0x80 | (cp & 63)));
    } else {
      s.push_back((char)(0xF0 | (cp >> 18)));
      s.push_back((char)(0x80 | ((cp >> 12) & 63)));
      s.push_back((char)(0x80 | ((cp >> 6) & 63)));
      s.push_back((char)(0x80 | (cp & 63)));
    }
  }
  uint32_t hex4()
  {
    if (end - p < 4) error("truncated \\u escape");
    uint32_t v = 0;
    for (int i = 0; i < 4; ++i) {
      const char c = *p++;
      v <<= 4;
      if (c >= '0' && c <= '9')
        v |= (uint32_t)(c - '0');
      else if (c >= 'a' && c <= 'f')
        v |= (uint32_t)(c - 'a' + 10);
      else if (c >= 'A' && c <= 'F')
        v |= (uint32_t)(c - 'A' + 10);
      else
        error("bad \\u escape");
    }
    return v;
  }
  std::string string()
  {
    if (p >= end || *p != '"') error("expected string");
    ++p;
    std::string s;
    while (p < end && *p != '"') {
      if (*p == '\\') {
        if (++p >= end) error("truncated escape");
        const char c = *p++;
        switch (c) {
          case '"': s.push_back('"'); break;
          case '\\': s.push_back('\\'); break;
          case '/': s.push_back('/'); break;
          case 'b': s.push_back('\b'); break;
          case 'f': s.push_back('\f'); break;
          case 'n': s.push_back('\n'); break;
          case 'r': s.push_back('\r'); break;
          case 't': s.push_back('\t'); break;
          case 'u': {
            uint32_t cp = hex4();
            if (cp >= 0xD800 && cp < 0xDC00 && end - p >= 6 && p[0] == '\\' && p[1] == 'u') {
              p += 2;
              const uint32_t lo = hex4();
              cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
            }
            utf8(s, cp);
            break;
          }
          default: error("bad escape");
        }
      } else {
        s.push_back(*p++);
      }
    }
    if (p >= end) error("unterminated string");
    ++p;
    return s;
  }
  Json value(int depth = 0)
  {
    if (depth > 256) error("nesting too deep");
    ws();
    if (p >= end) error("unexpected end of input");
    Json v;
    const char c = *p;
    if (c == '{') {
      ++p;
      v.type = Json::Object;
      ws();
      if (p < end && *p == '}') {
        ++p;
        return v;
      }
      for (;;) {
        ws();
        std::string key = string();
        ws();
        if (p >= end || *p != ':') error("expected ':'");
        ++p;
        v.obj.emplace_back(std::move(key), value(depth + 1));
        ws();
        if (p < end && *p == ',') {
          ++p;
          continue;
        }
        if (p < end && *p == '}') {
          ++p;
          return v;
        }
        error("expected ',' or '}'");
      }
    }
    if (c == '[') {
      ++p;
      v.type = Json::Array;
      ws();
      if (p < end && *p == ']') {
        ++p;
        return v;
      }
      for (;;) {
        v.arr.push_back(value(depth + 1));
        ws();
        if (p < end && *p == ',') {
          ++p;
          continue;
        }
        if (p < end && *p == ']') {
          ++p;
          return v;
        }
        error("expected ',' or ']'");
      }
    }
    if (c == '"') {
      v.type = Json::String;
      v.str = string();
      return v;
    }
    if (end - p >= 4 && std::strncmp(p, "true", 4) == 0) {
      p += 4;
      v.type = Json::Bool;
      v.b = true;
      return v;
    }
    if (end - p >= 5 && std::strncmp(p, "false", 5) == 0) {
      p += 5;
      v.type = Json::Bool;
      return v;
    }
    if (end - p >= 4 && std::strncmp(p, "null", 4) == 0) {
      p += 4;
      return v;
    }
    if (c == '-' || (c >= '0' && c <= '9')) {
      const char* q = p;
      while (q < end && (*q == '-' || *q == '+' || *q == '.' || *q == 'e' || *q == 'E' || (*q >= '0' && *q <= '9'))) ++q;
      const std::string tok(p, q);
      char* stop = nullptr;
      v.num = std::strtod(tok.c_str(), &stop);
      if (stop == tok.c_str()) error("bad number");
      p = q;
      v.type = Json::Number;
      return v;
    }
    error("unexpected character");
  }
};

// ---- URIs / buffers -----------------------------------------------------------------------------
std::string percent_decode(const std::string& s)
{
  std::string o;
  for (size_t i = 0; i < s.size(); ++i) {
    if (s[i] == '%' && i + 2 < s.size() + 0 && std::isxdigit((unsigned char)s[i + 1]) && std::isxdigit((unsigned char)s[i + 2])) {
      o.push_back((char)std::stoi(s.substr(i + 1, 2), nullptr, 16));
      i += 2;
    } else {
      o.push_back(s[i]);
    }
  }
  return o;
}

bool decode_data_uri(const std::string& uri, std::vector<uint8_t>& out)
{
  if (uri.compare(0, 5, "data:") != 0) return false;
  const size_t comma = uri.find(',');
  if (comma == std::string::npos || uri.find(";base64") == std::string::npos || uri.find(";base64") > comma)
    throw std::runtime_error("unsupported data URI");
  int val = 0, bits = -8;
  for (size_t i = comma + 1; i < uri.size(); ++i) {
    const char c = uri[i];
    int d;
    if (c >= 'A' && c <= 'Z')
      d = c - 'A';
    else if (c >= 'a' && c <= 'z')
      d = c - 'a' + 26;
    else if (c >= '0' && c <= '9')
      d = c - '0' + 52;
    else if (c == '+' || c == '-')
      d = 62;
    else if (c == '/' || c == '_')
      d = 63;
    else
      continue;  // padding / whitespace
    val = (val << 6) | d;
    bits += 6;
    if (bits >= 0) {
      out.push_back((uint8_t)((val >> bits) & 0xff));
      bits -= 8;
    }
  }
  return true;
}

struct GltfModel {
  Json root;
  std::vector<std::vector<uint8_t>> buffers;

  // accessor -> pointer to its first element, element stride and count
  // (tinygltf Accessor::ByteStride: the bufferView's byteStride, or the packed element size)
  const uint8_t* accessor_data(int accessor_id, int& stride, int& count) const
  {
    const auto& accessors = root.array("accessors");
    if (accessor_id < 0 || accessor_id >= (int)accessors.size()) throw std::runtime_error("accessor index out of range");
    const Json& acc = accessors[accessor_id];
    const int view_id = acc.integer("bufferView", -1);
    const auto& views = root.array("bufferViews");
    if (view_id < 0 || view_id >= (int)views.size()) throw std::runtime_error("accessor without a bufferView");
    const Json& view = views[view_id];
    const int buffer_id = view.integer("buffer", -1);
    if (buffer_id < 0 || buffer_id >= (int)buffers.size()) throw std::runtime_error("bufferView without a buffer");
    int comp_size;
    switch (acc.integer("componentType", 0)) {
      case 5120: case 5121: comp_size = 1; break;
      case 5122: case 5123: comp_size = 2; break;
      case 5124: case 5125: case 5126: comp_size = 4; break;
      case 5130: comp_size = 8; break;
      default: comp_size = -1;
    }
    const std::string type = acc.string("type");
    const int n_comp = type == "SCALAR" ? 1 : type == "VEC2" ? 2 : type == "VEC3" ? 3 : type == "VEC4" ? 4
                     : type == "MAT2" ? 4 : type == "MAT3" ? 9 : type == "MAT4" ? 16 : -1;
    const int view_stride = view.integer("byteStride", 0);
    if (comp_size <= 0 || n_comp <= 0)
      stride = -1;
    else if (view_stride == 0)
      stride = comp_size * n_comp;
    else
      stride = (view_stride % comp_size) ? -1 : view_stride;
    count = acc.integer("count", 0);
    const size_t offset = (size_t)view.number("byteOffset", 0.0) + (size_t)acc.number("byteOffset", 0.0);
    const std::vector<uint8_t>& buf = buffers[buffer_id];
    if (stride > 0 && count > 0 && offset + (size_t)stride * (size_t)(count - 1) + (size_t)comp_size * n_comp > buf.size())
      throw std::runtime_error("accessor reaches past the end of its buffer");
    return buf.data() + offset;
  }
};

// ---- glm-compatible algebra -----------------------------------------------------------------------
mat4 trs_matrix(const vec3& t, const quat& q, const vec3& s)
{
  // glm::translate(I, t)
  mat4 m = mat4::identity();
  m[3][0] = t.x;
  m[3][1] = t.y;
  m[3][2] = t.z;
  // *= glm::mat4_cast(q)
  mat4 r = mat4::identity();
  const float qxx = q.x * q.x, qyy = q.y * q.y, qzz = q.z * q.z, qxz = q.x * q.z, qxy = q.x * q.y, qyz = q.y * q.z,
              qwx = q.w * q.x, qwy = q.w * q.y, qwz = q.w * q.z;
  r[0][0] = 1.0f - 2.0f * (qyy + qzz);
  r[0][1] = 2.0f * (qxy + qwz);
  r[0][2] = 2.0f * (qxz - qwy);
  r[1][0] = 2.0f * (qxy - qwz);
  r[1][1] = 1.0f - 2.0f * (qxx + qzz);
  r[1][2] = 2.0f * (qyz + qwx);
  r[2][0] = 2.0f * (qxz + qwy);
  r[2][1] = 2.0f * (qyz - qwx);
  r[2][2] = 1.0f - 2.0f * (qxx + qyy);
  m = m * r;
  // glm::scale(m, s): columns scaled
  for (int k = 0; k < 4; ++k) {
    m[0][k] *= s.x;
    m[1][k] *= s.y;
    m[2][k] *= s.z;
  }
  return m;
}

float mixf(float x, float y, float a) { return x * (1.0f - a) + y * a; }
vec3 mix3(const vec3& x, const vec3& y, float a) { return vec3(mixf(x.x, y.x, a), mixf(x.y, y.y, a), mixf(x.z, y.z, a)); }
quat mixq(const quat& x, const quat& y, float a)
{
  const float cos_theta = (x.w * y.w + x.x * y.x) + (x.y * y.y + x.z * y.z);  // glm::dot(qua) association
  quat r;
  if (cos_theta > 1.0f - std::numeric_limits<float>::epsilon()) {
    r.w = mixf(x.w, y.w, a);
    r.x = mixf(x.x, y.x, a);
    r.y = mixf(x.y, y.y, a);
    r.z = mixf(x.z, y.z, a);
  } else {
    const float angle = std::acos(cos_theta);
    const float s0 = std::sin((1.0f - a) * angle), s1 = std::sin(a * angle), d = std::sin(angle);
    r.w = (x.w * s0 + y.w * s1) / d;
    r.x = (x.x * s0 + y.x * s1) / d;
    r.y = (x.y * s0 + y.y * s1) / d;
    r.z = (x.z * s0 + y.z * s1) / d;
  }
  return r;
}

template <typename T, typename Mix>
T keyframe(const std::vector<float>& input, const std::vector<T>& output, float time, Mix mix)
{
  const float t = std::fmod(time, input[input.size() - 1]);
  const int idx1 = (int)(std::lower_bound(input.begin(), input.end(), t) - input.begin());
  const int idx0 = std::max(idx1 - 1, 0);
  const float h = t - input[idx0];  // sic: not divided by the key spacing (scene.h:174)
  return mix(output[idx0], output[std::min<size_t>(idx1, output.size() - 1)], h);
}

void apply_node_transforms(Scene& sc, const Node& node, const mat4& parent)
{
  const mat4 m = parent * node.transform;
  if (node.camera_id != -1) {
    sc.m_has_camera_transform = true;
    sc.m_camera_transform = m;
  }
  if (node.submesh_id != -1) sc.m_transforms[node.submesh_id] = m;
  for (const Node& c : node.children) apply_node_transforms(sc, c, m);
}

struct GltfLoader {
  Scene& sc;
  const GltfModel& model;
  int indices_offset = 0;
  int prev_indices_size = 0;

  std::vector<uint8_t> on_path;  // nodes of the current root-to-node chain (cycle guard)

  Node load_node(int node_idx, int depth = 0)
  {
    const auto& nodes = model.root.array("nodes");
    if (node_idx < 0 || node_idx >= (int)nodes.size()) throw std::runtime_error("node index out of range");
    // a glTF node hierarchy is a forest; a file that lists an ancestor as a child would recurse for ever
    if (on_path.size() != nodes.size()) on_path.assign(nodes.size(), 0);
    if (on_path[node_idx] || depth > 256) throw std::runtime_error("node hierarchy has a cycle or is too deep");
    on_path[node_idx] = 1;
    const Json& node = nodes[node_idx];
    Node n;
    n.idx = node_idx;

    vec3 translation(0, 0, 0), scale(1, 1, 1);
    quat rotation;
    const auto& jt = node.array("translation");
    if (jt.size() == 3) translation = vec3((float)jt[0].num, (float)jt[1].num, (float)jt[2].num);
    const auto& jr = node.array("rotation");
    if (jr.size() == 4) {
      rotation.x = (float)jr[0].num;
      rotation.y = (float)jr[1].num;
      rotation.z = (float)jr[2].num;
      rotation.w = (float)jr[3].num;
    }
    const auto& js = node.array("scale");
    if (js.size() == 3) scale = vec3((float)js[0].num, (float)js[1].num, (float)js[2].num);
    n.transform = trs_matrix(translation, rotation, scale);
    const auto& jm = node.array("matrix");
    if (jm.size() == 16)
      for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) n.transform[c][r] = (float)jm[4 * c + r].num;

    const int mesh_id = node.integer("mesh", -1);
    if (mesh_id != -1) {
      const auto& meshes = model.root.array("meshes");
      if (mesh_id < 0 || mesh_id >= (int)meshes.size()) throw std::runtime_error("mesh index out of range");
      n.submesh_id = (int)sc.m_submesh_offsets.size();
      for (const Json& prim : meshes[mesh_id].array("primitives")) {
        int stride = 0, count = 0;
        const uint8_t* raw = model.accessor_data(prim.integer("indices", -1), stride, count);
        if (stride != 2) throw std::runtime_error("indices stride is not ushort");
        const int n_tris = count / 3;
        for (int i = 0; i < n_tris; ++i) {
          uint16_t id[3];
          std::memcpy(id, raw + 6 * (size_t)i, 6);
          sc.m_indices.push_back(make_uint3(id[0] + indices_offset, id[1] + indices_offset, id[2] + indices_offset));
        }
        int n_vertices = 0;
        const Json* attrs = prim.find("attributes");
        if (attrs && attrs->type == Json::Object) {
          for (const auto& kv : attrs->obj) {
            if (kv.second.type != Json::Number) continue;
            const int acc = (int)kv.second.num;
            if (kv.first == "POSITION") {
              const uint8_t* d = model.accessor_data(acc, stride, count);
              if (stride != 12) throw std::runtime_error("positions stride is not float3");
              for (int i = 0; i < count; ++i) {
                float v[3];
                std::memcpy(v, d + 12 * (size_t)i, 12);
                sc.m_vertices.push_back(make_float3(v[0], v[1], v[2]));
              }
              n_vertices += count;
            } else if (kv.first == "NORMAL") {
              const uint8_t* d = model.accessor_data(acc, stride, count);
              if (stride != 12) throw std::runtime_error("normals stride is not float3");
              for (int i = 0; i < count; ++i) {
                float v[3];
                std::memcpy(v, d + 12 * (size_t)i, 12);
                sc.m_normals.push_back(make_float3(v[0], v[1], v[2]));
              }
            } else if (kv.first == "TEXCOORD_0") {
              const uint8_t* d = model.accessor_data(acc, stride, count);
              if (stride != 8) throw std::runtime_error("texcoord stride is not float2");
              for (int i = 0; i < count; ++i) {
                float v[2];
                std::memcpy(v, d + 8 * (size_t)i, 8);
                sc.m_texcoords.push_back(make_float2(v[0], 1.0f - v[1]));
              }
            }
          }
        }
        // A primitive without NORMAL / TEXCOORD_0 leaves those arrays shorter than the vertex array (the reference
        // does the same and then reads out of bounds on the device): pad them, with area-weighted vertex normals
        // of this primitive's triangles and zero texture coordinates.
        if (sc.m_normals.size() < sc.m_vertices.size()) {
          const size_t first = sc.m_normals.size();
          sc.m_normals.resize(sc.m_vertices.size(), make_float3(0.f, 0.f, 0.f));
          for (size_t f = sc.m_indices.size() - (size_t)n_tris; f < sc.m_indices.size(); ++f) {
            const uint3 t = sc.m_indices[f];
            if (t.x >= sc.m_vertices.size() || t.y >= sc.m_vertices.size() || t.z >= sc.m_vertices.size()) continue;
            const float3 a = sc.m_vertices[t.x], b = sc.m_vertices[t.y], c = sc.m_vertices[t.z];
            const float3 e1 = make_float3(b.x - a.x, b.y - a.y, b.z - a.z), e2 = make_float3(c.x - a.x, c.y - a.y, c.z - a.z);
            const float3 n = make_float3(e1.y * e2.z - e1.z * e2.y, e1.z * e2.x - e1.x * e2.z, e1.x * e2.y - e1.y * e2.x);
            for (uint32_t v : {t.x, t.y, t.z})
              if (v >= first) {
                sc.m_normals[v].x += n.x;
                sc.m_normals[v].y += n.y;
                sc.m_normals[v].z += n.z;
              }
          }
          for (size_t v = first; v < sc.m_normals.size(); ++v) {
            float3& n = sc.m_normals[v];
            const float len = std::sqrt(n.x * n.x + n.y * n.y + n.z * n.z);
            n = len > 0.f ? make_float3(n.x / len, n.y / len, n.z / len) : make_float3(0.f, 1.f, 0.f);
          }
        }
        if (sc.m_texcoords.size() < sc.m_vertices.size()) sc.m_texcoords.resize(sc.m_vertices.size(), make_float2(0.f, 0.f));
        const unsigned int material = (unsigned int)prim.integer("material", -1);
        for (int i = 0; i < n_tris; ++i) sc.m_material_ids.push_back(material);
        for (int i = 0; i < n_tris; ++i) sc.m_instance_ids.push_back((unsigned int)sc.m_submesh_offsets.size());
        indices_offset += n_vertices;
      }
      sc.m_submesh_offsets.push_back((unsigned int)prev_indices_size);
      sc.m_submesh_n_faces.push_back((unsigned int)(sc.m_indices.size() - prev_indices_size));
      prev_indices_size = (int)sc.m_indices.size();
    } else {
      n.submesh_id = -1;
    }
    for (const Json& child : node.array("children")) n.children.push_back(load_node((int)child.num, depth + 1));
    on_path[node_idx] = 0;
    return n;
  }
};

}  // namespace

void Scene::load_gltf(const std::filesystem::path& filepath)
{
  const std::string where = "failed to load " + filepath.generic_string();
  try {
    GltfModel model;
    {
      const std::vector<uint8_t> text = codec::read_file_bytes(filepath.generic_string());
      JsonParser parser{reinterpret_cast<const char*>(text.data()), reinterpret_cast<const char*>(text.data()) + text.size()};
      model.root = parser.value();
      if (model.root.type != Json::Object) throw std::runtime_error("root is not an object");
    }
    for (const Json& b : model.root.array("buffers")) {
      std::vector<uint8_t> bytes;
      const std::string uri = b.string("uri");
      if (uri.empty()) throw std::runtime_error("buffer without uri (GLB containers are not supported)");
      if (!decode_data_uri(uri, bytes))
        bytes = codec::read_file_bytes((filepath.parent_path() / percent_decode(uri)).generic_string());
      const size_t declared = (size_t)b.number("byteLength", 0.0);
      if (bytes.size() < declared) throw std::runtime_error("buffer shorter than its byteLength");
      model.buffers.push_back(std::move(bytes));
    }

    // ---- materials (scene.cpp:487-560) ----
    for (const Json& jm : model.root.array("materials")) {
      Material mat;
      static const Json none;
      const Json* pmr_p = jm.find("pbrMetallicRoughness");
      const Json& pmr = pmr_p ? *pmr_p : none;
      const auto& bc = pmr.array("baseColorFactor");
      if (bc.size() == 4)
        mat.base_color = make_float3((float)bc[0].num, (float)bc[1].num, (float)bc[2].num);
      else
        mat.base_color = make_float3(1.0f, 1.0f, 1.0f);
      auto texture_index = [](const Json& parent, const char* key) {
        const Json* t = parent.find(key);
        return (t && t->type == Json::Object) ? t->integer("index", -1) : -1;
      };
      int id = texture_index(pmr, "baseColorTexture");
      if (id != -1) mat.base_color_texture_id = id;
      mat.specular_roughness = (float)pmr.number("roughnessFactor", 1.0);
      mat.metalness = (float)pmr.number("metallicFactor", 1.0);
      id = texture_index(pmr, "metallicRoughnessTexture");
      if (id != -1) mat.metallic_roughness_texture_id = id;
      if (const Json* ext = jm.find("extensions")) {
        if (const Json* cc = ext->find("KHR_materials_clearcoat")) {
          // Value::GetNumberAsDouble / GetNumberAsInt: a real or an int, anything else reads as 0
          auto as_number = [](const Json& v) { return v.type == Json::Number ? v.num : 0.0; };
          if (const Json* v = cc->find("clearcoatFactor")) mat.coat = (float)as_number(*v);
          if (const Json* v = cc->find("clearcoatTexture")) mat.coat_texture_id = (int)as_number(*v);
          if (const Json* v = cc->find("clearcoatRoughnessFactor")) mat.coat_roughness = (float)as_number(*v);
          if (const Json* v = cc->find("clearcoatRoughnessTexture")) mat.coat_roughness_texture_id = (int)as_number(*v);
        }
      }
      // tinygltf always holds three emissive factors (default 0,0,0)
      const auto& ef = jm.array("emissiveFactor");
      mat.emission = 1.0f;
      mat.emission_color = ef.size() == 3 ? make_float3((float)ef[0].num, (float)ef[1].num, (float)ef[2].num)
                                          : make_float3(0.0f, 0.0f, 0.0f);
      id = texture_index(jm, "emissiveTexture");
      if (id != -1) mat.emission_texture_id = id;
      id = texture_index(jm, "normalTexture");
      if (id != -1) mat.normalmap_texture_id = id;
      m_materials.push_back(mat);
    }

    // ---- textures (scene.cpp:562-570): all NONCOLOR, from the image file ----
    const auto& images = model.root.array("images");
    for (const Json& jt : model.root.array("textures")) {
      const int source = jt.integer("source", -1);
      if (source < 0 || source >= (int)images.size()) throw std::runtime_error("texture without a source image");
      const std::string uri = images[source].string("uri");
      if (uri.empty() || uri.compare(0, 5, "data:") == 0)
        throw std::runtime_error("texture images must be external files");
      m_textures.push_back(Texture(filepath.parent_path() / percent_decode(uri), TextureType::NONCOLOR));
    }

    // ---- nodes of scene 0 (scene.cpp:572-578) ----
    const auto& scenes = model.root.array("scenes");
    if (scenes.empty()) throw std::runtime_error("no scene");
    const size_t first_root = m_nodes.size();
    GltfLoader loader{*this, model};
    for (const Json& root : scenes[0].array("nodes")) m_nodes.push_back(loader.load_node((int)root.num));
    m_transforms.resize(m_submesh_offsets.size());
    update_transform();

    // ---- animations (scene.cpp:584-660) ----
    for (const Json& ja : model.root.array("animations")) {
      Animation anim;
      const auto& channels = ja.array("channels");
      const auto& samplers = ja.array("samplers");
      if (channels.empty()) throw std::runtime_error("animation without channels");
      auto target_of = [](const Json& ch, int& node, std::string& path) {
        const Json* t = ch.find("target");
        node = t ? t->integer("node", -1) : -1;
        path = t ? t->string("path") : std::string();
      };
      int target_node;
      std::string path;
      target_of(channels[0], target_node, path);
      anim.node_idx = target_node;
      anim.root_slot = -1;
      for (size_t r = first_root; r < m_nodes.size(); ++r)
        if (m_nodes[r].idx == target_node) {
          anim.root_slot = (int)r;
          break;
        }
      if (anim.root_slot < 0) throw std::runtime_error("invalid target node");
      for (const Json& ch : channels) {
        int node_unused;
        target_of(ch, node_unused, path);
        const int sampler_id = ch.integer("sampler", -1);
        if (sampler_id < 0 || sampler_id >= (int)samplers.size()) throw std::runtime_error("animation sampler out of range");
        const Json& smp = samplers[sampler_id];
        int in_stride = 0, in_count = 0;
        const uint8_t* in_raw = model.accessor_data(smp.integer("input", -1), in_stride, in_count);
        if (in_stride != 4) throw std::runtime_error("unsupported animation input");
        std::vector<float>* in_dst = path == "translation" ? &anim.translation_input
                                   : path == "rotation"    ? &anim.rotation_input
                                   : path == "scale"       ? &anim.scale_input
                                                           : nullptr;
        for (int i = 0; i < in_count && in_dst; ++i) {
          float v;
          std::memcpy(&v, in_raw + 4 * (size_t)i, 4);
          in_dst->push_back(v);
        }
        int out_stride = 0, out_count = 0;
        const uint8_t* out_raw = model.accessor_data(smp.integer("output", -1), out_stride, out_count);
        if (in_count != out_count) throw std::runtime_error("animation input size is not equal to output size");
        if (path == "translation" || path == "scale") {
          if (out_stride != 12) throw std::runtime_error("invalid output stride");
          for (int i = 0; i < out_count; ++i) {
            float v[3];
            std::memcpy(v, out_raw + 12 * (size_t)i, 12);
            (path == "translation" ? anim.translation_output : anim.scale_output).push_back(vec3(v[0], v[1], v[2]));
          }
        } else if (path == "rotation") {
          if (out_stride != 16) throw std::runtime_error("invalid output stride");
          for (int i = 0; i < out_count; ++i) {
            float v[4];
            std::memcpy(v, out_raw + 16 * (size_t)i, 16);
            quat q;
            q.x = v[0], q.y = v[1], q.z = v[2], q.w = v[3];
            anim.rotation_output.push_back(q);
          }
        }
      }
      m_animations.push_back(anim);
    }
  } catch (const std::runtime_error& e) {
    const std::string what = e.what();
    if (what.compare(0, 14, "failed to load") == 0) throw;
    throw std::runtime_error(where + ": " + what);
  }
}

void Scene::update_transform()
{
  for (const Node& node : m_nodes) apply_node_transforms(*this, node, mat4::identity());
}

void Scene::update_animation(float time)
{
  for (const Animation& a : m_animations) {
    vec3 translation(0, 0, 0), scale(1, 1, 1);
    quat rotation;
    if (!a.translation_input.empty()) translation = keyframe(a.translation_input, a.translation_output, time, mix3);
    if (!a.rotation_input.empty()) rotation = keyframe(a.rotation_input, a.rotation_output, time, mixq);
    if (!a.scale_input.empty()) scale = keyframe(a.scale_input, a.scale_output, time, mix3);
    if (a.root_slot >= 0 && a.root_slot < (int)m_nodes.size()) m_nodes[a.root_slot].transform = trs_matrix(translation, rotation, scale);
  }
  update_transform();
}

}  // namespace fredholm
