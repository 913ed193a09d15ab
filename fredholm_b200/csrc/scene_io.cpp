// File-backed parts of the scene interface: image textures and glTF.
#include <stdexcept>

#include "fredholm/scene.h"

namespace fredholm
{

Texture::Texture(const std::filesystem::path& filepath, const TextureType& texture_type)
    : m_texture_type(texture_type)
{
  throw std::runtime_error("failed to load " + filepath.generic_string() + ": image decoding not available yet");
}

FloatTexture::FloatTexture(const std::filesystem::path& filepath)
{
  throw std::runtime_error("failed to load " + filepath.generic_string() + ": image decoding not available yet");
}

void Scene::load_gltf(const std::filesystem::path& filepath)
{
  throw std::runtime_error("failed to load " + filepath.generic_string() + ": glTF not available yet");
}

void Scene::update_animation(float) {}

}  // namespace fredholm
