"""Shared generator of BSDF known-answer cases (n x 40 floats: ShadingParams[30], wo[3],
entering, wi[3], u, v[2])."""
import numpy as np

from fredholm_b200.types import shading_params

MATERIAL_CLASSES = {
    "lambert": dict(base_color=(0.7, 0.6, 0.5), specular_color=(0, 0, 0)),
    "oren_nayar": dict(base_color=(0.7, 0.6, 0.5), specular_color=(0, 0, 0), diffuse_roughness=0.6),
    "dielectric": dict(base_color=(0.2, 0.5, 0.8), specular_color=(1, 1, 1), specular_roughness=0.3),
    "glossy": dict(base_color=(0.2, 0.5, 0.8), specular_color=(0.9, 0.8, 0.7), specular_roughness=0.05),
    "metal": dict(base_color=(0.95, 0.64, 0.54), specular_color=(1, 1, 1), metalness=1.0, specular_roughness=0.2),
    "half_metal": dict(base_color=(0.9, 0.9, 0.2), specular_color=(0.8, 0.8, 0.8), metalness=0.5,
                       specular_roughness=0.4),
    "coat": dict(base_color=(0.7, 0.05, 0.05), specular_color=(1, 1, 1), coat=1.0, coat_roughness=0.05,
                 coat_color=(1.0, 0.9, 0.8), specular_roughness=0.3),
    "glass": dict(base_color=(1, 1, 1), specular_color=(1, 1, 1), transmission=0.9,
                  transmission_color=(0.9, 1.0, 0.95), specular_roughness=0.15),
    "clear_glass": dict(base_color=(1, 1, 1), specular_color=(1, 1, 1), transmission=1.0, specular_roughness=0.01),
    "sheen": dict(base_color=(0.1, 0.15, 0.5), specular_color=(0, 0, 0), sheen=1.0, sheen_color=(0.8, 0.8, 1.0),
                  sheen_roughness=0.3),
    "thin_sss": dict(base_color=(0.8, 0.4, 0.3), specular_color=(0.5, 0.5, 0.5), subsurface=0.6,
                     subsurface_color=(1.0, 0.5, 0.4), thin_walled=1.0),
    "everything": dict(diffuse=0.8, base_color=(0.6, 0.5, 0.4), diffuse_roughness=0.2, specular=0.9,
                       specular_color=(0.9, 0.9, 0.8), specular_roughness=0.25, metalness=0.3, coat=0.7,
                       coat_color=(0.9, 0.95, 1.0), coat_roughness=0.1, transmission=0.4,
                       transmission_color=(0.8, 0.9, 1.0), sheen=0.5, sheen_color=(1, 0.9, 0.8),
                       sheen_roughness=0.4, subsurface=0.3, subsurface_color=(1, 0.6, 0.5), thin_walled=1.0),
}


def make_cases(n_per_class=64, seed=7):
    rng = np.random.default_rng(seed)
    rows, labels = [], []
    for name, kw in MATERIAL_CLASSES.items():
        sp = shading_params(**kw)
        for i in range(n_per_class):
            wo = rng.normal(size=3)
            wo[1] = abs(wo[1]) + 0.05          # the integrator always flips the frame towards the viewer
            wo /= np.linalg.norm(wo)
            wi = rng.normal(size=3)
            wi /= np.linalg.norm(wi)
            entering = 1.0 if (i % 4) != 3 else 0.0
            u = rng.uniform()
            v = rng.uniform(size=2)
            rows.append(np.concatenate([sp, wo, [entering], wi, [u], v]))
            labels.append(name)
    return np.asarray(rows, dtype=np.float32), labels


def make_cases_below_horizon(n_per_class=32, seed=13):
    """wo BELOW the shading horizon (wo.y < 0).  The integrator flips the frame towards the viewer with the
    GEOMETRIC normal; a normal / bump map can still tilt the shading normal away from the viewer, so these
    directions do reach BSDF::eval (pt.cu:709-742, 762-763).  Used for the eval columns only: sampling from
    below the horizon is ill-conditioned in the reference itself (0/0 pdfs)."""
    rng = np.random.default_rng(seed)
    rows, labels = [], []
    for name, kw in MATERIAL_CLASSES.items():
        sp = shading_params(**kw)
        for i in range(n_per_class):
            wo = rng.normal(size=3)
            wo[1] = -(abs(wo[1]) + 0.02)
            wo /= np.linalg.norm(wo)
            wi = rng.normal(size=3)
            wi /= np.linalg.norm(wi)
            entering = 1.0 if (i % 4) != 3 else 0.0
            rows.append(np.concatenate([sp, wo, [entering], wi, [rng.uniform()], rng.uniform(size=2)]))
            labels.append(name)
    return np.asarray(rows, dtype=np.float32), labels
