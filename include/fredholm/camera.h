// Camera of the rendering core: camera-to-world transform + thin-lens parameters,
// WASD-style movement and look-around.
//
// Mirrors the public members and methods of the reference's fredholm::Camera
// (fredholm/include/fredholm/camera.h:22-136) so application code keeps compiling:
// m_transform, m_fov, m_F, m_focus, move(), lookAround(), get_origin(),
// set_origin().  The reference builds m_transform as inverse(lookAt(...)) with glm;
// glm is not a dependency here, so the camera-to-world matrix is written down in
// closed form (columns: right, up, -forward, origin) in a small column-major mat4
// that indexes like glm::mat4 (m[column][row]).
#pragma once
#include <cuda_runtime.h>

#include <cmath>

namespace fredholm
{

struct vec3 {
  float x = 0, y = 0, z = 0;
  vec3() = default;
  vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};
inline vec3 operator+(const vec3& a, const vec3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(const vec3& a, const vec3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator*(float s, const vec3& a) { return {s * a.x, s * a.y, s * a.z}; }
inline vec3& operator+=(vec3& a, const vec3& b)
{
  a = a + b;
  return a;
}
inline vec3& operator-=(vec3& a, const vec3& b)
{
  a = a - b;
  return a;
}
inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(const vec3& a, const vec3& b)
{
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline vec3 normalize(const vec3& a)
{
  const float inv = 1.0f / std::sqrt(dot(a, a));
  return inv * a;
}

// column-major 4x4, m[c][r] like glm::mat4
struct mat4 {
  float c[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  float* operator[](int col) { return c[col]; }
  const float* operator[](int col) const { return c[col]; }
  static mat4 identity() { return mat4(); }
};

mat4 operator*(const mat4& a, const mat4& b);
mat4 inverse(const mat4& m);

enum class CameraMovement {
  FORWARD,
  BACKWARD,
  RIGHT,
  LEFT,
  UP,
  DOWN,
};

struct Camera {
  mat4 m_transform;  // camera to world

  float m_fov;
  float m_F;      // F number
  float m_focus;  // focus distance

  float m_movement_speed;
  float m_look_around_speed;

  vec3 m_origin;
  vec3 m_forward;
  vec3 m_right;
  vec3 m_up;
  float m_phi;
  float m_theta;

  Camera()
      : m_fov(0.5f * static_cast<float>(M_PI)),
        m_F(8.0f),
        m_focus(10000.0f),
        m_movement_speed(10.0f),
        m_look_around_speed(0.1f),
        m_phi(270.0f),
        m_theta(90.0f)
  {
  }

  Camera(const float3& origin, float fov = 0.5f * static_cast<float>(M_PI), float F = 8.0f,
         float focus = 10000.0f, float movement_speed = 1.0f, float look_around_speed = 0.1f)
      : m_fov(fov),
        m_F(F),
        m_focus(focus),
        m_movement_speed(movement_speed),
        m_look_around_speed(look_around_speed),
        m_phi(270.0f),
        m_theta(90.0f)
  {
    m_origin = vec3(origin.x, origin.y, origin.z);
    m_forward = vec3(0, 0, -1);
    m_right = normalize(cross(m_forward, vec3(0, 1, 0)));
    m_up = normalize(cross(m_right, m_forward));
    update_transform();
  }

  float3 get_origin() const { return make_float3(m_origin.x, m_origin.y, m_origin.z); }

  void set_origin(const float3& origin)
  {
    m_origin = vec3(origin.x, origin.y, origin.z);
    update_transform();
  }

  void move(const CameraMovement& direction, float dt)
  {
    const float velocity = m_movement_speed * dt;
    switch (direction) {
      case CameraMovement::FORWARD: m_origin += velocity * m_forward; break;
      case CameraMovement::BACKWARD: m_origin -= velocity * m_forward; break;
      case CameraMovement::RIGHT: m_origin += velocity * m_right; break;
      case CameraMovement::LEFT: m_origin -= velocity * m_right; break;
      case CameraMovement::UP: m_origin += velocity * m_up; break;
      case CameraMovement::DOWN: m_origin -= velocity * m_up; break;
    }
    update_transform();
  }

  void lookAround(float d_phi, float d_theta)
  {
    m_phi += m_look_around_speed * d_phi;
    if (m_phi < 0.0f) m_phi = 360.0f;
    if (m_phi > 360.0f) m_phi = 0.0f;

    m_theta += m_look_around_speed * d_theta;
    if (m_theta < 0.0f) m_theta = 180.0f;
    if (m_theta > 180.0f) m_theta = 0.0f;

    const float phi_rad = m_phi / 180.0f * static_cast<float>(M_PI);
    const float theta_rad = m_theta / 180.0f * static_cast<float>(M_PI);
    m_forward = vec3(std::cos(phi_rad) * std::sin(theta_rad), std::cos(theta_rad),
                     std::sin(phi_rad) * std::sin(theta_rad));
    m_right = normalize(cross(m_forward, vec3(0.0f, 1.0f, 0.0f)));
    m_up = normalize(cross(m_right, m_forward));
    update_transform();
  }

  // Packs the 3x4 row-major camera-to-world block the kernels consume
  // (what Renderer::render does in the reference, renderer.h:678-684).
  void to_rows(float out12[12]) const
  {
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) out12[4 * r + c] = m_transform[c][r];
  }

 private:
  // camera-to-world of a view looking from m_origin along m_forward with m_up:
  // the inverse of lookAt(origin, origin + 0.01 forward, up)
  void update_transform()
  {
    const vec3 f = normalize((m_origin + 0.01f * m_forward) - m_origin);
    const vec3 s = normalize(cross(f, m_up));
    const vec3 u = cross(s, f);
    m_transform = mat4();
    m_transform[0][0] = s.x, m_transform[0][1] = s.y, m_transform[0][2] = s.z;
    m_transform[1][0] = u.x, m_transform[1][1] = u.y, m_transform[1][2] = u.z;
    m_transform[2][0] = -f.x, m_transform[2][1] = -f.y, m_transform[2][2] = -f.z;
    m_transform[3][0] = m_origin.x, m_transform[3][1] = m_origin.y, m_transform[3][2] = m_origin.z;
  }
};

}  // namespace fredholm
