"""GPU parity for triangle primitives + Möller–Trumbore (extension — the reference never
instantiates class Triangle, main.cpp:107-216: PARITY UNPINNED by the reference).  Oracle: the CPU restatement
(brute-force and BVH traversal over the same trees)."""
import numpy as np
import pytest

import conftest as T

rt = T.rtds_b200
pytestmark = pytest.mark.gpu


def _rays(n, seed, tris):
    rng = np.random.default_rng(seed)
    cen = tris.reshape(-1, 3, 3).mean(1)
    tgt = cen[rng.integers(0, cen.shape[0], n)] + rng.normal(size=(n, 3)).astype(np.float32) * np.float32(0.03)
    o = np.zeros((n, 3), np.float32)
    o[n // 2:] = rng.normal(size=(n - n // 2, 3)).astype(np.float32) * np.float32(20)
    d = tgt - o
    return o, (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)


@pytest.mark.parametrize("builder", ["median", "lbvh30", "lbvh63", "sah"])
def test_triangle_trees_bit_exact_and_hits(gpu_ctx, oracle, builder):
    # The reference's median split DROPS a whole range when no centre lies below the midpoint of its bounds
    # (accelerators.h:321-327) — which two scene-spanning ground triangles provoke at the second level — so the
    # median builder gets the soup without the ground; the other builders get it with.
    tris, mat = T.triangle_scene(20000, 51, ground=(builder != "median"))
    gpu_ctx.set_triangles(tris, mat)
    if builder == "median":
        st = gpu_ctx.build(rt.BVH)
        rc, nodes_o, order_o, _ = oracle.build_bvh(tris, prim_type=1)
        acc, tie = rt.BVH, 0
    elif builder == "sah":
        st = gpu_ctx.build(rt.BVH, mode=rt.MODE_SAH)
        nodes_o, order_o, _ = oracle.build_sah(tris, prim_type=1)
        acc, tie = rt.BVH, 1
    else:
        bits = int(builder[4:])
        st = gpu_ctx.build(rt.LBVH, mode=rt.MODE_TRUE, morton_bits=bits)
        nodes_o, order_o, _, _ = oracle.build_lbvh(tris, bits, prim_type=1)
        acc, tie = rt.LBVH, 1
    nodes, order = gpu_ctx.export_bvh()
    assert np.array_equal(order, order_o) and nodes.tobytes() == nodes_o.tobytes()
    o, d = _rays(20000, 52, tris)
    h_o, t_o, cand = oracle.trace(tris, nodes, order, o, d, tie_by_objid=tie, prim_type=1)
    for exact in (True, False):
        h, t, st = gpu_ctx.trace(acc, o, d, exact=exact)
        assert np.array_equal(h, h_o), f"{np.count_nonzero(h != h_o)} hit ids differ (exact={exact})"
        assert t.tobytes() == t_o.tobytes()
    hn, tn, _ = gpu_ctx.trace(rt.NONE, o, d)
    hn_o, tn_o, _ = oracle.trace(tris, None, None, o, d, prim_type=1)
    assert np.array_equal(hn, hn_o) and tn.tobytes() == tn_o.tobytes()
    assert (h_o >= 0).mean() > 0.3
    # The BVH path may only lose grazing hits to the float slab test. (Not for the reference's median split: with
    # primitives of different sizes it DROPS every range whose centres all lie at or above the midpoint of the
    # range's bounds, accelerators.h:321-327 — reproduced bit for bit above, so those triangles are simply absent.)
    if builder != "median":
        assert np.count_nonzero(h_o != hn_o) <= 0.002 * h_o.size
    else:
        print("median split kept %d of %d triangles (the reference's dropped ranges)" % (order.size, tris.shape[0]))


@pytest.mark.parametrize("acc,exact,shadows", [(rt.LBVH, False, 0), (rt.BVH, True, 0), (rt.LBVH, False, 1), (rt.NONE, False, 0)])
def test_triangle_render_matches_restatement(gpu_ctx, oracle, acc, exact, shadows):
    n = 6000 if acc != rt.NONE else 600
    tris, mat = T.triangle_scene(n, 53, ground=(acc != rt.BVH))
    gpu_ctx.set_triangles(tris, mat)
    if acc == rt.NONE:
        nodes = order = None
        tie = 0
    else:
        gpu_ctx.build(acc, mode=rt.MODE_TRUE if acc == rt.LBVH else rt.MODE_COMPAT)
        nodes, order = gpu_ctx.export_bvh()
        tie = 1 if acc == rt.LBVH else 0
    W, H, spp = 256, 192, 2
    rgb, hit, accum, st = gpu_ctx.render(acc, W, H, spp, want_hit=True, want_accum=True, exact=exact, shadows=shadows)
    rgb_o, hit_o, accum_o, _ = oracle.render_rows(tris, mat, nodes, order, W, H, spp, tie_by_objid=tie, want_accum=True,
                                                   shadows=shadows, prim_type=1)
    assert np.array_equal(hit, hit_o)
    assert accum.tobytes() == accum_o.tobytes() and np.array_equal(rgb, rgb_o)
    assert (hit >= 0).mean() > 0.2


def test_triangle_kdtree_any_hit(gpu_ctx):
    # small triangles: a breadth-first KD build holds every node of a level at once, and triangles much larger than
    # their spacing are duplicated into exponentially many cells (the builder reports RTDS_ERR_CAPACITY beyond 128 n)
    tris, mat = T.triangle_scene(20000, 55, ground=False)
    c = tris.reshape(-1, 3, 3).mean(1, keepdims=True)
    tris = np.ascontiguousarray((c + (tris.reshape(-1, 3, 3) - c) * np.float32(0.25)).reshape(-1, 9), np.float32)
    gpu_ctx.set_triangles(tris, mat)
    st = gpu_ctx.build(rt.KDTREE)
    o, d = _rays(20000, 56, tris)
    h_kd, _, _ = gpu_ctx.trace(rt.KDTREE, o, d)
    h_none, _, _ = gpu_ctx.trace(rt.NONE, o, d)
    diff = np.count_nonzero((h_kd > 0) != (h_none >= 0))
    assert diff <= 0.001 * h_none.size, f"KD any-hit disagrees with brute force on {diff} rays"


def test_geometric_triangle_test_equals_reference_class_triangle(gpu_ctx, ref, oracle):
    """SURVEY 8a14, PINNED: with tri_geometric the GPU evaluates Triangle::rayTriangleIntersect as the reference compiles it (the
    geometric branch, main.cpp:163-215). The compiled, unmodified reference instantiates its own class Triangle and answers the
    same rays (oracle/ref_harness.cpp: ref_triangle_trace): hit triangle and t bits must be identical - brute force (NONE loop,
    main.cpp:376-386) and through every BVH builder (unpruned reference traversal and the ordered one), rays from the camera origin
    and from displaced origins (where main.cpp:182's plane distance is wrong: reproduced bug for bug)."""
    tris, mat = T.triangle_scene(3000, 7, ground=True)
    gpu_ctx.set_triangles(tris, mat)
    o, d = _rays(8000, 9, tris)
    d[:40] = np.asarray([0, 0, -1], np.float32)
    h_r, t_r, cnt_r = ref.triangle_trace(tris, o, d)
    h_n, t_n, _ = gpu_ctx.trace(rt.NONE, o, d, tri_geometric=True)
    assert np.array_equal(h_n, h_r) and t_n.tobytes() == t_r.tobytes()
    h_p, t_p, _ = oracle.trace(tris, None, None, o, d, prim_type=2)
    assert np.array_equal(h_p, h_r) and t_p.tobytes() == t_r.tobytes()
    assert (h_r[:4000] >= 0).mean() > 0.5 and cnt_r.max() >= 2
    # through the trees: candidates are the leaves whose boxes the reference's slab test passes (boxIntersect), then the same test.
    # Origin-0 rays only (what render() casts): for displaced origins the reference's t does not lie on the ray, so a box test
    # along the ray says nothing about it.
    o0, d0 = o[:4000], d[:4000]
    for acc, kw, tie in ((rt.LBVH, dict(mode=rt.MODE_TRUE), 1), (rt.BVH, dict(mode=rt.MODE_SAH), 1)):
        gpu_ctx.build(acc, **kw)
        nodes, order = gpu_ctx.export_bvh()
        h_o, t_o, _ = oracle.trace(tris, nodes, order, o0, d0, tie_by_objid=tie, prim_type=2)
        for exact in (True, False):
            h, t, _ = gpu_ctx.trace(acc, o0, d0, exact=exact, tri_geometric=True)
            assert np.array_equal(h, h_o) and t.tobytes() == t_o.tobytes(), (acc, exact)
        lost = np.count_nonzero(h_o != h_r[:4000])
        assert lost <= 0.003 * 4000, lost              # only grazing hits the float slab test rejects
    # rendered frame with the geometric test == the restatement's (shading uses the same v0v1 x v0v2 normal, main.cpp:165-168)
    W, H, spp = 200, 150, 2
    rgb, hit, accum, _ = gpu_ctx.render(rt.BVH, W, H, spp, want_hit=True, want_accum=True, tri_geometric=1)
    rgb_o, hit_o, accum_o, _ = oracle.render_rows(tris, mat, nodes, order, W, H, spp, tie_by_objid=1, want_accum=True, prim_type=2)
    assert np.array_equal(hit, hit_o) and accum.tobytes() == accum_o.tobytes() and np.array_equal(rgb, rgb_o)
    rgb_mt = gpu_ctx.render(rt.BVH, W, H, spp)[0]
    assert (np.abs(rgb.astype(int) - rgb_mt.astype(int)).max(axis=2) > 0).mean() < 0.01     # same picture as Moeller-Trumbore but for edges
