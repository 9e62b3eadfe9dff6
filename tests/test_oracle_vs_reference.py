"""CPU, build container only: the oracle port against the compiled UNMODIFIED reference on inputs the golden
files do not cover (random synthetic scenes through ref_scene_from_spheres, arbitrary rays, 4 spp)."""
import numpy as np
import pytest

import conftest as T

rt = T.rtds_b200


@pytest.mark.parametrize("n,seed,q", [(500, 1, 0.0), (3000, 2, 0.0), (2000, 3, 1.0), (64, 4, 0.5)])
def test_port_tree_equals_reference_tree(ref, oracle, n, seed, q):
    sph, mat = T.synthetic_scene(n, seed)
    if q:
        sph[:n, :3] = np.round(sph[:n, :3] / np.float32(q)) * np.float32(q)
    ref.scene_from_spheres(sph, mat)
    for acc, n_use in ((rt.BVH, sph.shape[0]), (rt.LBVH, sph.shape[0] - 1)):
        total, _ = ref.build(acc)
        nodes_r, objs_r, _ = ref.bvh_linear()
        rc, nodes, order, _ = oracle.build_bvh(sph, n_use)
        assert rc == 0 and total == nodes.shape[0]
        assert np.array_equal(order, objs_r) and nodes.tobytes() == nodes_r.tobytes()


def test_port_trace_and_render_equal_reference(ref, oracle):
    sph, mat = T.synthetic_scene(4000, 7)
    ref.scene_from_spheres(sph, mat)
    ref.build(rt.BVH)
    rc, nodes, order, _ = oracle.build_bvh(sph)
    rng = np.random.default_rng(8)
    d = rng.normal(size=(5000, 3)).astype(np.float32) * np.float32([0.15, 0.15, 0]) + np.float32([0, 0, -1])
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = np.zeros((1, 3), np.float32)
    h_r, t_r, cand_r = ref.trace(rt.BVH, o, d)
    h_o, t_o, cand_o = oracle.trace(sph, nodes, order, o, d)
    assert np.array_equal(h_r, h_o) and t_r.tobytes() == t_o.tobytes() and cand_r == cand_o
    h_rn, t_rn, _ = ref.trace(rt.NONE, o, d[:300])
    h_on, t_on, _ = oracle.trace(sph, None, None, o, d[:300])
    assert np.array_equal(h_rn, h_on) and t_rn.tobytes() == t_on.tobytes()
    rgb_r, _, acc_r, _ = ref.render_rows(rt.BVH, 200, 150, 4, 40, 110, want_accum=True)
    rgb_o, _, acc_o, _ = oracle.render_rows(sph, mat, nodes, order, 200, 150, 4, 40, 110, want_accum=True)
    assert acc_r.tobytes() == acc_o.tobytes() and np.array_equal(rgb_r, rgb_o)


def test_port_jitter_equals_libstdcxx(ref, oracle):
    assert np.array_equal(ref.jitter(100000), oracle.jitter(100000))


def test_port_castray_materials_equal_reference(ref, oracle):
    """REFLECTION_AND_REFRACTION / REFLECTION branches + recursion depth (main.cpp:417-447), 3 lights, through the
    reference's own castRay."""
    sph, mat = T.material_scene(1500, 11)
    lights = np.asarray([[0, 3, 30, 10, 1, 1, 1], [20, 30, -40, 1, 0.5, 0.4, 0.3], [-30, 5, -70, 1, 0.2, 0.3, 0.6]], np.float32)
    ref.scene_from_spheres(sph, mat)
    ref.lib.ref_set_lights(lights.ctypes.data_as(T.C.c_void_p), 3)
    ref.build(rt.BVH)
    rc, nodes, order, _ = oracle.build_bvh(sph)
    rgb_r, _, acc_r, _ = ref.render_rows(rt.BVH, 240, 180, 2, want_accum=True)
    rgb_o, _, acc_o, _ = oracle.render_rows(sph, mat, nodes, order, 240, 180, 2, lights=lights, want_accum=True)
    assert oracle.last_ray_counts[2] > 100                     # secondary rays were actually traced
    # glibc's powf(x, 25) (main.cpp:475) is not correctly rounded: 0.09 % of inputs come out 1 ulp away from the exact
    # x^25 the port (and the GPU) compute, so float sums may differ in the last place where a highlight contributes;
    # north_star's tolerance is 1/255 per channel.
    assert np.allclose(acc_r, acc_o, rtol=3e-7, atol=1e-7)
    assert np.count_nonzero(acc_r != acc_o) < 0.001 * acc_r.size
    assert np.abs(rgb_r.astype(int) - rgb_o.astype(int)).max() <= 1
    ref.lib.ref_set_lights(np.asarray([[0, 3, 30, 10, 1, 1, 1]], np.float32).ctypes.data_as(T.C.c_void_p), 1)


def phantom_none_hits(sph, d, hit_none, where):
    """For the rays `where` (flat indices): True where the NONE loop's hit is a float phantom — in float64 the ray
    passes OUTSIDE the sphere (raySphereIntersect's d2 = l.l - tca*tca cancels catastrophically at |l| ~ 60:
    accelerators.h:85-87), so no spatial subdivision can hold the 'hit point' in a cell the sphere overlaps."""
    out = np.zeros(len(where), bool)
    for k, r in enumerate(where):
        s = sph[hit_none.reshape(-1)[r]].astype(np.float64)
        dv = d[r].astype(np.float64)
        dv /= np.linalg.norm(dv)
        tca = s[:3] @ dv
        out[k] = np.sqrt(max(0.0, s[:3] @ s[:3] - tca * tca)) > s[3]
    return out


def test_port_kd_closest_hit_on_the_reference_kd_tree(ref, oracle):
    """Closest-hit KD traversal (extension; the reference's KD path is any-hit): the restatement walks the reference's
    OWN KdAccelNode[] for the default config's 307,200 primary rays and must find the NONE loop's hits (golden, from
    the unmodified reference) — except where NONE's hit is a float phantom — and agree with kdtreeIntersect's any-hit mask."""
    sph, mat = T.bunny_scene()
    ref.scene_from_spheres(sph, mat)
    ref.build(rt.KDTREE)
    nodes, idx, bounds = ref.kd_dump()
    W, H = 640, 480
    rc, bn, bo, _ = oracle.build_bvh(sph)
    _, _, _, dirs = oracle.render_rows(sph, mat, bn, bo, W, H, 1, want_dirs=True)      # main.cpp:554-557's directions
    d = dirs.reshape(-1, 3)
    gold = np.load(T.GOLDEN + "/bunny_hits_640x480.npz")
    hit, t, tests = oracle.kd_closest(sph, nodes, idx, bounds, np.zeros((1, 3), np.float32), d)
    where = np.nonzero(hit != gold["hit_none"].reshape(-1))[0]
    print("KD closest hit on the reference's tree: %d of %d pixels differ from the NONE golden hits, %.2f prim tests/ray" %
          (len(where), W * H, tests / (W * H)))
    assert len(where) <= 31                                           # 0.01 %
    assert phantom_none_hits(sph, d, gold["hit_none"], where).all()    # every one of them: the true ray misses NONE's sphere
    mask = np.unpackbits(np.load(T.GOLDEN + "/bunny_kd_mask_640x480.npy")).astype(bool).reshape(-1)
    assert np.array_equal(hit >= 0, mask)                             # same hit/miss as the reference's kdtreeIntersect
    same = hit == gold["hit_none"].reshape(-1)
    t_none = oracle.trace(sph, None, None, np.zeros((1, 3), np.float32), d[same][::97])[1]
    assert t[same][::97].tobytes() == t_none.tobytes()                # and the same tnear bits


def test_port_geometric_triangle_test_equals_reference_class_triangle(ref, oracle):
    """the reference's class Triangle (main.cpp:107-216) is never instantiated by its loader, but its
    rayTriangleIntersect IS compiled (the geometric branch; MOLLER_TRUMBORE is never defined). The harness instantiates the class
    and calls it; the port's restatement (prim_type 2) must give the same hit triangle and the same t bits - rays from the camera
    origin (what render() casts) and from arbitrary origins (where main.cpp:182's `N.orig + d` is wrong: kept bug for bug)."""
    tris, _ = T.triangle_scene(3000, 7, ground=True)
    rng = np.random.default_rng(8)
    n = 6000
    cen = tris.reshape(-1, 3, 3).mean(1)
    tgt = cen[rng.integers(0, cen.shape[0], n)] + rng.normal(size=(n, 3)).astype(np.float32) * np.float32(0.2)
    o = np.zeros((n, 3), np.float32)
    o[n // 2:] = rng.normal(size=(n - n // 2, 3)).astype(np.float32) * np.float32(5)
    d = tgt - o
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    d[:50] = np.asarray([0, 0, -1], np.float32)                  # axis-parallel rays: zero components in the edge tests
    d[50:60, 1] = 0
    h_r, t_r, cnt_r = ref.triangle_trace(tris, o, d)
    h_p, t_p, _ = oracle.trace(tris, None, None, o, d, prim_type=2)
    assert np.array_equal(h_r, h_p) and t_r.tobytes() == t_p.tobytes()
    assert (h_r[: n // 2] >= 0).mean() > 0.5 and cnt_r.max() >= 2
    # and it is a different function from Moeller-Trumbore (prim_type 1): same triangles for origin-0 rays up to edge cases,
    # different answers for displaced origins
    h_m, t_m, _ = oracle.trace(tris, None, None, o, d, prim_type=1)
    assert (h_m[: n // 2] == h_p[: n // 2]).mean() > 0.99
    assert (h_m[n // 2:] != h_p[n // 2:]).mean() > 0.05
