"""Synthetic stand-ins for BASELINE.json's configs 4 and 5: the reference ships neither the XYZ dragon nor the
Sci-Fi city + Trees models (main.cpp:641-646 names them, models/ does not hold them), and there is no network. Same primitive
counts, same primitive kind the reference would make of them (one small sphere per vertex, main.cpp:668-692) + its ground sphere.
Used by bench.py --workload config4|config5, tools/run_configs.py and the GPU tests."""
import numpy as np

GROUND = (0.93591022, -105.47120094, -43.2363205, 100.0)   # main.cpp:703
# config 5's "multi-light": main.cpp:775's light + two more (position, radius, emission)
LIGHTS3 = np.asarray([[0, 3, 30, 10, 1, 1, 1], [20, 30, -40, 1, 0.5, 0.4, 0.3], [-30, 5, -70, 1, 0.2, 0.3, 0.6]], np.float32)


def torus_knot_scene(n, seed=7):
    """config 4 stand-in: n spheres on a displaced (2,3) torus-knot tube inside the camera frustum."""
    rng = np.random.default_rng(seed)
    t = rng.uniform(0, 2 * np.pi, n)
    phi = rng.uniform(0, 2 * np.pi, n)
    R, r, tube = 9.0, 3.5, 1.2 + 0.25 * np.sin(7 * t)
    cx = (R + r * np.cos(3 * t)) * np.cos(2 * t)
    cy = (R + r * np.cos(3 * t)) * np.sin(2 * t)
    cz = r * np.sin(3 * t)
    # tube cross-section in a frame that is good enough for a point cloud
    nx, ny, nz = np.cos(2 * t) * np.cos(phi), np.sin(2 * t) * np.cos(phi), np.sin(phi)
    p = np.stack([cx + tube * nx, cy + tube * ny, cz + tube * nz], 1) + rng.normal(size=(n, 3)) * 0.01
    sph = np.zeros((n + 1, 4), np.float32)
    sph[:n, :3] = p.astype(np.float32) * np.float32(0.9) + np.asarray([0, 0, -70], np.float32)
    sph[:n, 3] = 0.02
    sph[n] = np.asarray(GROUND, np.float32)
    mat = np.zeros_like(sph)
    mat[:n, 0], mat[:n, 1] = 0.8, 0.7
    return sph, mat


def city_trees_scene(seed=5):
    """config 5 stand-in: 90,811 'city' prims (axis-aligned boxes of points) + 252,178 'tree' prims (clustered blobs),
    10 % REFLECTION_AND_REFRACTION and 10 % REFLECTION materials."""
    rng = np.random.default_rng(seed)
    n_city, n_tree = 90811, 252178
    b = rng.integers(0, 60, n_city)
    origin = np.stack([(b % 10) * 4.0 - 20, np.full(60, -8.0)[b], -(b // 10) * 6.0 - 50], 1)
    size = np.stack([rng.uniform(1, 3, 60), rng.uniform(2, 14, 60), rng.uniform(1, 3, 60)], 1)[b]
    city = origin + rng.uniform(0, 1, (n_city, 3)) * size
    k = rng.integers(0, 300, n_tree)
    tc = np.stack([rng.uniform(-25, 25, 300), rng.uniform(-6, 2, 300), rng.uniform(-95, -45, 300)], 1)[k]
    trees = tc + rng.normal(size=(n_tree, 3)) * rng.uniform(0.3, 1.2, (300, 1))[k]
    p = np.concatenate([city, trees]).astype(np.float32)
    n = p.shape[0]
    sph = np.zeros((n + 1, 4), np.float32)
    sph[:n, :3] = p
    sph[:n, 3] = 0.05
    sph[n] = np.asarray(GROUND, np.float32)
    mat = np.zeros_like(sph)
    mat[:n, :3] = rng.uniform(0.1, 0.9, (n, 3)).astype(np.float32)
    u = rng.uniform(size=n)
    mat[:n, 3] = np.where(u < 0.1, 1.0, np.where(u < 0.2, 2.0, 0.0))
    return sph, mat
