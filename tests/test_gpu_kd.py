"""GPU parity: KD-tree build (K8) + any-hit traversal against the reference's KDTREE path.  std::sort's order of
equal edges is implementation-defined in the reference, so topology is not pinned; parity is on hit results
(SURVEY.md §7.6): the any-hit mask of the default config, and — when the compiled reference travelled to this
box — random scenes and rays through the real kdtreeIntersect."""
import os

import numpy as np
import pytest

import conftest as T

rt = T.rtds_b200
G = T.load_golden_json()
pytestmark = pytest.mark.gpu


def _kd_invariants(nodes, idx, n_prims):
    leaf = (nodes["w1"] & 3) == 3
    inner = ~leaf
    n = nodes.shape[0]
    above = (nodes["w1"][inner] >> 2).astype(np.int64)
    me = np.nonzero(inner)[0]
    assert np.all(above > me + 1) and np.all(above < n)          # left child = me+1, right child later in DFS order
    np_leaf = (nodes["w1"][leaf] >> 2).astype(np.int64)
    assert np.array_equal(np_leaf, nodes["w2"][leaf])            # nPrims == nPrimitivesTest
    multi = np_leaf > 1
    assert int(np_leaf[multi].sum()) == idx.shape[0]
    if idx.size:
        assert idx.min() >= 0 and idx.max() < n_prims
    one = nodes["w0"][leaf][np_leaf == 1]
    seen = np.zeros(n_prims, bool)
    seen[one] = True
    seen[idx] = True
    assert seen.all(), "every primitive must be referenced by at least one leaf"


def test_kd_default_config_mask_and_counts(gpu_ctx):
    sph, mat = T.bunny_scene()
    gpu_ctx.set_spheres(sph, mat)
    st = gpu_ctx.build(rt.KDTREE)
    assert st["max_depth"] == 28                                   # "Depth is: 28"
    nodes, idx, bounds = gpu_ctx.export_kd()
    _kd_invariants(nodes, idx, sph.shape[0])
    rgb, hit, _, rs = gpu_ctx.render(rt.KDTREE, 640, 480, 1, want_hit=True)
    mask = hit > 0
    gold = np.unpackbits(np.load(T.GOLDEN + "/bunny_kd_mask_640x480.npy")).astype(bool).reshape(480, 640)
    diff = int(np.count_nonzero(mask != gold))
    print("KD: %d counted nodes (reference %d), %d allocated; mask differs on %d px; md5 %s (reference %s)" %
          (st["total_nodes"], G["kd"]["total_nodes"], st["alloc_nodes"], diff, T.ppm_md5(rgb), G["kd"]["ppm_md5"]))
    assert diff == 0, f"{diff} pixels differ from the reference's kdtreeIntersect mask"
    assert T.ppm_md5(rgb) == G["kd"]["ppm_md5"] == "f7feeee16ced45991d070b88fd5ac3fd"   # the reference's KDTREE frame
    assert st["total_nodes"] == G["kd"]["total_nodes"] == 46475                          # "Number of nodes: 46475"
    assert abs(st["total_nodes"] - G["kd"]["total_nodes"]) <= 0.01 * G["kd"]["total_nodes"]
    # hit pixels are black, misses are the sky (main.cpp:365-369)
    assert np.all(rgb[mask] == 0) and np.all(rgb[~mask] == np.array([153, 204, 255], np.uint8))


def test_kd_vs_compiled_reference_random_scenes(gpu_ctx):
    if not os.path.exists(os.path.join(T.ROOT, "oracle", "_ref", "libref_oracle.so")):
        pytest.skip("compiled reference not on this box")
    ref = T.Ref()
    rng = np.random.default_rng(41)
    for n, seed in ((200, 1), (5000, 2), (30000, 3)):
        sph, mat = T.synthetic_scene(n, seed)
        ref.scene_from_spheres(sph, mat)
        total_ref, _ = ref.build(rt.KDTREE)
        gpu_ctx.set_spheres(sph, mat)
        st = gpu_ctx.build(rt.KDTREE)
        nodes, idx, bounds = gpu_ctx.export_kd()
        _kd_invariants(nodes, idx, sph.shape[0])
        m = 20000
        tgt = sph[rng.integers(0, n, m), :3] + rng.normal(size=(m, 3)).astype(np.float32) * np.float32(0.06)
        o = np.zeros((m, 3), np.float32)
        o[m // 2:] = rng.normal(size=(m - m // 2, 3)).astype(np.float32) * np.float32(15)
        d = tgt - o
        d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        # structure vs the reference's own node array (informational where std::sort ties decide membership)
        nref = ref.lib.ref_kd_dump(None, 0, None, 0, None, None)
        ref_nodes = np.zeros(nref, rt.KD_NODE_DTYPE)
        nidx = T.C.c_int()
        ref_idx = np.zeros(max(idx.shape[0] * 2, 16), np.int32)
        ref.lib.ref_kd_dump(ref_nodes.ctypes.data_as(T.C.c_void_p), nref, ref_idx.ctypes.data_as(T.C.c_void_p), ref_idx.size, T.C.byref(nidx), None)
        same_struct = nref == nodes.shape[0] and np.array_equal(ref_nodes["w1"], nodes["w1"])
        inner = (nodes["w1"] & 3) != 3
        same_split = same_struct and np.array_equal(ref_nodes["w0"][inner], nodes["w0"][inner])
        print("n=%d: node array length ref %d / gpu %d; flags+children identical: %s; split planes identical: %s" %
              (n, nref, nodes.shape[0], same_struct, same_split))
        h_ref, _, _ = ref.trace(rt.KDTREE, o, d)
        h, t, _ = gpu_ctx.trace(rt.KDTREE, o, d)
        diff = int(np.count_nonzero((h > 0) != (h_ref > 0)))
        print("n=%d: reference %d counted nodes, GPU %d; any-hit differs on %d of %d rays" % (n, total_ref, st["total_nodes"], diff, m))
        assert diff <= m // 2000
        assert abs(st["total_nodes"] - total_ref) <= max(4, 0.02 * total_ref)


# ---- closest-hit KD traversal + shading (extension, rtds_render_params.kd_closest / RTDS_TRACE_KD_CLOSEST) ---------------
def test_kd_closest_hit_default_config(gpu_ctx, oracle):
    """The default config through the KD-tree with closest-hit traversal: hit ids and tnear bit-exact against the CPU
    restatement walking the SAME exported node array; against the reference's NONE hits (golden) only float-phantom
    rays differ; where the hit is the BVH path's the shaded pixel is the BVH frame's (= the reference's output.ppm)."""
    from test_oracle_vs_reference import phantom_none_hits
    sph, mat = T.bunny_scene()
    gpu_ctx.set_spheres(sph, mat)
    gpu_ctx.build(rt.KDTREE)
    nodes, idx, bounds = gpu_ctx.export_kd()
    W, H = 640, 480
    rgb, hit, accum, rs = gpu_ctx.render(rt.KDTREE, W, H, 1, want_hit=True, want_accum=True, kd_closest=1)
    rc, bn, bo, _ = oracle.build_bvh(sph)
    rgb_b, hit_b, acc_b, dirs = oracle.render_rows(sph, mat, bn, bo, W, H, 1, want_dirs=True, want_accum=True)
    d = dirs.reshape(-1, 3)
    hit_o, t_o, tests_o = oracle.kd_closest(sph, nodes, idx, bounds, np.zeros((1, 3), np.float32), d)
    assert np.array_equal(hit.reshape(-1), hit_o)
    assert rs["prim_tests"] == tests_o                       # the same walk: identical number of primitive tests
    h_t, t_t, _ = gpu_ctx.trace(rt.KDTREE, np.zeros((1, 3), np.float32), d, kd_closest=True)
    assert np.array_equal(h_t, hit_o) and t_t.tobytes() == t_o.tobytes()
    gold = np.load(T.GOLDEN + "/bunny_hits_640x480.npz")
    where = np.nonzero(hit_o != gold["hit_none"].reshape(-1))[0]
    print("KD closest hit: %d of %d pixels differ from the reference's NONE hits (all float-phantom hits of NONE); "
          "%.2f prim tests/ray, kernel %.3f ms" % (len(where), W * H, tests_o / (W * H), rs["ms_kernel"]))
    assert len(where) <= 31 and phantom_none_hits(sph, d, gold["hit_none"], where).all()
    same = hit == hit_b
    assert same.mean() > 0.999
    assert np.array_equal(rgb[same], rgb_b[same]) and accum[same].tobytes() == acc_b[same].tobytes()
    # any-hit mode is unchanged by the flag's existence
    rgb_a, _, _, _ = gpu_ctx.render(rt.KDTREE, W, H, 1)
    assert T.ppm_md5(rgb_a) == G["kd"]["ppm_md5"]


def test_kd_closest_hit_random_rays_and_multisample(gpu_ctx, oracle):
    rng = np.random.default_rng(77)
    for n, seed in ((300, 5), (20000, 6)):
        sph, mat = T.synthetic_scene(n, seed)
        gpu_ctx.set_spheres(sph, mat)
        gpu_ctx.build(rt.KDTREE)
        nodes, idx, bounds = gpu_ctx.export_kd()
        m = 30000
        tgt = sph[rng.integers(0, n, m), :3] + rng.normal(size=(m, 3)).astype(np.float32) * np.float32(0.06)
        o = np.zeros((m, 3), np.float32)
        o[m // 2:] = rng.normal(size=(m - m // 2, 3)).astype(np.float32) * np.float32(15)     # origins inside and outside the tree bounds
        d = tgt - o
        d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        h, t, st = gpu_ctx.trace(rt.KDTREE, o, d, kd_closest=True)
        h_o, t_o, tests_o = oracle.kd_closest(sph, nodes, idx, bounds, o, d)
        assert np.array_equal(h, h_o) and t.tobytes() == t_o.tobytes() and st["prim_tests"] == tests_o
        h_n, t_n, _ = gpu_ctx.trace(rt.NONE, o, d)
        diff = int(np.count_nonzero(h != h_n))
        print("n=%d: closest-hit KD differs from brute force on %d of %d rays" % (n, diff, m))
        assert diff <= m // 1000
        agree = h == h_n
        assert t[agree].tobytes() == t_n[agree].tobytes()
    # 4 spp frame: float sums equal the LBVH path's wherever all samples hit the same primitives (checked through the frames)
    W, H = 320, 240
    rgb_k, hit_k, acc_k, _ = gpu_ctx.render(rt.KDTREE, W, H, 4, want_hit=True, want_accum=True, kd_closest=1)
    gpu_ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
    rgb_l, hit_l, acc_l, _ = gpu_ctx.render(rt.LBVH, W, H, 4, want_hit=True, want_accum=True)
    differ = np.count_nonzero(np.any(acc_k != acc_l, axis=-1))
    print("4 spp: %d of %d pixels differ between the closest-hit KD frame and the LBVH frame" % (differ, W * H))
    # per ray the KD walk and the BVH walk each deviate from NONE on grazing rays only (~0.005 % / ~0.04 %); a pixel has 4 rays
    assert differ <= W * H // 100
    assert np.abs(rgb_k.astype(int) - rgb_l.astype(int)).max(axis=-1).reshape(-1)[np.all(acc_k == acc_l, axis=-1).reshape(-1)].max() == 0


def test_kd_closest_hit_triangles(gpu_ctx, oracle):
    tris, mat = T.triangle_scene(20000, 55, ground=False)
    c = tris.reshape(-1, 3, 3).mean(1, keepdims=True)
    tris = np.ascontiguousarray((c + (tris.reshape(-1, 3, 3) - c) * np.float32(0.25)).reshape(-1, 9), np.float32)
    gpu_ctx.set_triangles(tris, mat)
    gpu_ctx.build(rt.KDTREE)
    nodes, idx, bounds = gpu_ctx.export_kd()
    rng = np.random.default_rng(56)
    m = 20000
    tgt = tris.reshape(-1, 3, 3).mean(1)[rng.integers(0, tris.shape[0], m)] + rng.normal(size=(m, 3)).astype(np.float32) * np.float32(0.1)
    d = (tgt / np.linalg.norm(tgt, axis=1, keepdims=True)).astype(np.float32)
    o = np.zeros((1, 3), np.float32)
    h, t, _ = gpu_ctx.trace(rt.KDTREE, o, d, kd_closest=True)
    h_o, t_o, _ = oracle.kd_closest(tris, nodes, idx, bounds, o, d, prim_type=1)
    assert np.array_equal(h, h_o) and t.tobytes() == t_o.tobytes()
    h_n, t_n, _ = gpu_ctx.trace(rt.NONE, o, d)
    assert np.count_nonzero(h != h_n) <= m // 1000 and (h >= 0).mean() > 0.2
    with pytest.raises(rt.RtdsError):                      # primary rays only on the KD path
        gpu_ctx.render(rt.KDTREE, 64, 48, 1, kd_closest=1, shadows=1)
