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

// Public members and method names are the reference's (application code reads and writes m_fov, m_F,
// m_focus and m_transform directly); the bodies live in fredholm_b200/csrc/camera.cpp.
struct Camera {
  mat4 m_transform;  // camera to world

  float m_fov = 1.57079632679489661923f;  // radians
  float m_F = 8.0f;                        // F number
  float m_focus = 10000.0f;                // focus distance

  float m_movement_speed = 10.0f;
  float m_look_around_speed = 0.1f;

  vec3 m_origin, m_forward, m_right, m_up;
  float m_phi = 270.0f;   // degrees, azimuth
  float m_theta = 90.0f;  // degrees from +y

  Camera() = default;
  Camera(const float3& origin, float fov = 1.57079632679489661923f, float F = 8.0f, float focus = 10000.0f,
         float movement_speed = 1.0f, float look_around_speed = 0.1f);

  float3 get_origin() const;
  void set_origin(const float3& origin);
  void move(const CameraMovement& direction, float dt);
  void lookAround(float d_phi, float d_theta);

  // the 3x4 row-major camera-to-world block the kernels consume (Renderer::render, renderer.h:678-684)
  void to_rows(float out12[12]) const;

 private:
  void set_view_direction(const vec3& forward);
  void update_transform();
};

}  // namespace fredholm
