// fredholm::Camera (include/fredholm/camera.h): walk / look-around camera of the applications, mirroring the
// behaviour of the reference's Camera (fredholm/include/fredholm/camera.h:22-136): same angles, speeds and
// resulting camera-to-world matrix, checked bit for bit against the reference class in tests/test_host_logic.py.
#include "fredholm/camera.h"

namespace fredholm
{

namespace
{
constexpr float kPiF = 3.14159265358979323846f;

// the reference snaps an angle that leaves [0, limit] to the opposite end instead of wrapping it
float snap(float angle, float limit)
{
  if (angle < 0.0f) return limit;
  if (angle > limit) return 0.0f;
  return angle;
}
}  // namespace

Camera::Camera(const float3& origin, float fov, float F, float focus, float movement_speed, float look_around_speed)
{
  m_fov = fov;
  m_F = F;
  m_focus = focus;
  m_movement_speed = movement_speed;
  m_look_around_speed = look_around_speed;
  m_origin = vec3(origin.x, origin.y, origin.z);
  set_view_direction(vec3(0.0f, 0.0f, -1.0f));
}

float3 Camera::get_origin() const { return make_float3(m_origin.x, m_origin.y, m_origin.z); }

void Camera::set_origin(const float3& origin)
{
  m_origin = vec3(origin.x, origin.y, origin.z);
  update_transform();
}

void Camera::move(const CameraMovement& direction, float dt)
{
  // FORWARD/BACKWARD, RIGHT/LEFT, UP/DOWN: the odd members walk against their axis
  const int code = static_cast<int>(direction);
  if (code >= 0 && code < 6) {  // anything else moves nothing (the reference's switch has no default)
    const vec3* axes[3] = {&m_forward, &m_right, &m_up};
    const float step = m_movement_speed * dt;
    const vec3 delta = step * *axes[code / 2];
    if (code & 1)
      m_origin -= delta;
    else
      m_origin += delta;
  }
  update_transform();
}

void Camera::lookAround(float d_phi, float d_theta)
{
  m_phi = snap(m_phi + m_look_around_speed * d_phi, 360.0f);
  m_theta = snap(m_theta + m_look_around_speed * d_theta, 180.0f);
  const float azimuth = m_phi / 180.0f * kPiF, polar = m_theta / 180.0f * kPiF;
  const float sin_polar = std::sin(polar);
  set_view_direction(vec3(std::cos(azimuth) * sin_polar, std::cos(polar), std::sin(azimuth) * sin_polar));
}

void Camera::to_rows(float out12[12]) const
{
  for (int row = 0; row < 3; ++row)
    for (int col = 0; col < 4; ++col) out12[4 * row + col] = m_transform[col][row];
}

void Camera::set_view_direction(const vec3& forward)
{
  m_forward = forward;
  m_right = normalize(cross(m_forward, vec3(0.0f, 1.0f, 0.0f)));
  m_up = normalize(cross(m_right, m_forward));
  update_transform();
}

// camera-to-world of a view from m_origin along m_forward with m_up: the inverse of
// lookAt(origin, origin + 0.01 forward, up), written in closed form (columns: side, up, -forward, origin)
void Camera::update_transform()
{
  const vec3 f = normalize((m_origin + 0.01f * m_forward) - m_origin);
  const vec3 side = normalize(cross(f, m_up));
  const vec3 up = cross(side, f);
  const vec3 cols[4] = {side, up, vec3(-f.x, -f.y, -f.z), m_origin};
  m_transform = mat4();
  for (int c = 0; c < 4; ++c) {
    m_transform[c][0] = cols[c].x;
    m_transform[c][1] = cols[c].y;
    m_transform[c][2] = cols[c].z;
  }
}

}  // namespace fredholm
