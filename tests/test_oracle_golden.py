"""CPU: the oracle port (oracle/oracle.cpp) against the golden vectors generated from the unmodified reference
(tests/golden/make_golden.py).  Runs anywhere g++ exists; no GPU, no /root/reference."""
import hashlib

import numpy as np
import pytest

import conftest as T

rt = T.rtds_b200
G = T.load_golden_json()


def tree_sha(nodes, order):
    return hashlib.sha256(nodes.tobytes() + np.ascontiguousarray(order, np.int32).tobytes()).hexdigest()


def test_loader_restatement_matches_golden_scene():
    v = T.bunny_vertices()
    assert v.shape[0] == G["bunny"]["n_vertices"]
    assert hashlib.sha256(v.tobytes()).hexdigest() == G["bunny"]["vertices_sha256"]
    sph, mat = rt.scene_from_vertices(v, 1)
    assert hashlib.sha256(sph.tobytes()).hexdigest() == G["bunny"]["scene_sha256"]


def test_oracle_scene_equals_numpy_scene(oracle):
    v = T.bunny_vertices()
    for clones in (1, 3):
        a, am = oracle.scene_from_vertices(v, clones)
        b, bm = rt.scene_from_vertices(v, clones)
        assert a.tobytes() == b.tobytes() and am.tobytes() == bm.tobytes()
    assert hashlib.sha256(rt.scene_from_vertices(v, 3)[0].tobytes()).hexdigest() == G["trees"]["bunny_clones3"]["scene_sha256"]


def test_oracle_bvh_tree_bunny(oracle):
    sph, _ = T.bunny_scene()
    rc, nodes, order, depth = oracle.build_bvh(sph)
    assert rc == 0
    e = G["default_config"]["BVH"]
    assert nodes.shape[0] == e["total_nodes"]
    assert tree_sha(nodes, order) == e["tree_sha256"]
    assert np.array_equal(order, np.load(T.GOLDEN + "/bunny_bvh_prim_order.npy"))


def test_oracle_lbvh_compat_tree_bunny(oracle):
    sph, _ = T.bunny_scene()
    rc, nodes, order, _ = oracle.build_bvh(sph, sph.shape[0] - 1)     # accelerators.h:583 drops the last object
    e = G["default_config"]["LBVH"]
    assert rc == 0 and nodes.shape[0] == e["total_nodes"] and tree_sha(nodes, order) == e["tree_sha256"]


def _scene(name):
    import make_golden  # noqa
    if name == "bunny_clones3":
        return T.bunny_scene(3)
    if name.startswith("synthetic_"):
        _, n, seed = name.split("_")
        n = int(n)
        return T.synthetic_scene(n if n != 2 else 1, int(seed[4:]))
    if name.startswith("quantised_"):
        _, n, seed, q = name.split("_")
        return make_golden.quantised_scene(int(n), int(seed[4:]), float(q[1:]))
    return None


@pytest.mark.parametrize("name", [k for k in G["trees"] if k != "armadillo"])
def test_oracle_trees_other_scenes(oracle, name):
    import sys
    sys.path.insert(0, T.GOLDEN)
    sph, _ = _scene(name)
    e = G["trees"][name]
    assert hashlib.sha256(sph.tobytes()).hexdigest() == e["scene_sha256"]
    rc, nodes, order, _ = oracle.build_bvh(sph)
    assert rc == 0 and nodes.shape[0] == e["BVH"]["total_nodes"]
    assert order.shape[0] == e["BVH"]["n_leaves"]          # < n when the reference dropped ranges (accelerators.h:321-327)
    assert tree_sha(nodes, order) == e["BVH"]["tree_sha256"]
    if "LBVH" in e:
        rc, nodes, order, _ = oracle.build_bvh(sph, sph.shape[0] - 1)
        assert rc == 0 and nodes.shape[0] == e["LBVH"]["total_nodes"]
        assert tree_sha(nodes, order) == e["LBVH"]["tree_sha256"]


def test_oracle_morton_kats(oracle):
    m = G["morton"]
    assert oracle.morton30([[.5, .5, .5]])[0] == m["half"] == 939524096
    assert oracle.morton30([[1, 1, 1]])[0] == m["ones"] == 0x3FFFFFFF
    rng = np.random.default_rng(5)
    pts = rng.uniform(-0.1, 1.1, size=(64, 3)).astype(np.float32)
    assert oracle.morton30(pts).tolist() == m["points_seed5"]
    sph, _ = T.bunny_scene()
    q = (sph[:1, :3] + np.float32(30)) / np.float32(1000)
    assert oracle.morton30(q)[0] == m["bunny_v0_refnorm"] == 84002


def test_oracle_jitter_kats(oracle):
    j = oracle.jitter(1000004)
    g = G["jitter"]
    assert [float(x).hex() for x in j[:8]] == g["first8_hex"]
    assert [float(x).hex() for x in j[1000000:1000004]] == g["at_1000000_hex"]
    assert hashlib.sha256(j[:1000000].tobytes()).hexdigest() == g["sha256_first_1e6"]
    assert np.array_equal(oracle.jitter(4, first=1000000), j[1000000:1000004])


def test_oracle_render_default_config_is_byte_exact(oracle):
    """The whole restated path (jitter, ray gen, traversal, sphere test, shading, quantise) reproduces the
    reference's shipped output.ppm byte for byte (md5 c69c6637...)."""
    sph, mat = T.bunny_scene()
    rc, nodes, order, _ = oracle.build_bvh(sph)
    rgb, hit, _, _ = oracle.render_rows(sph, mat, nodes, order, 640, 480, 1)
    assert T.ppm_md5(rgb) == G["default_config"]["BVH"]["ppm_md5"] == "c69c66375f2c6bda433f9f457a4b2b2e"
    gold = np.load(T.GOLDEN + "/bunny_hits_640x480.npz")
    assert np.array_equal(hit, gold["hit_bvh"])


def test_oracle_render_lbvh_compat_is_byte_exact(oracle):
    sph, mat = T.bunny_scene()
    rc, nodes, order, _ = oracle.build_bvh(sph, sph.shape[0] - 1)
    rgb, _, _, _ = oracle.render_rows(sph, mat, nodes, order, 640, 480, 1)
    assert T.ppm_md5(rgb) == G["default_config"]["LBVH"]["ppm_md5"]


def test_oracle_none_rows_match_golden_hits(oracle):
    """NONE brute force on a band of rows through the bunny (full frame takes a minute on one core)."""
    sph, mat = T.bunny_scene()
    y0, y1 = 236, 244
    rgb, hit, _, _ = oracle.render_rows(sph, mat, None, None, 640, 480, 1, y0, y1)
    gold = np.load(T.GOLDEN + "/bunny_hits_640x480.npz")
    assert np.array_equal(hit, gold["hit_none"][y0:y1])


def test_wide4_collapse_keeps_boxes_leaves_and_hits(oracle):
    """Definition of the 4-wide collapse (oracle.cpp, groundwork for the wide-node traversal of DESIGN.md 10.1): every leaf once,
    child boxes = the binary nodes' boxes, 2..4 children (largest-area entry expanded first), pre-order numbering; the collect-all trace over the wide tree returns
    the binary tree's hits and tnear bits (the candidate criterion is leaf-local) with fewer box tests per ray."""
    sph, mat = T.bunny_scene()
    for build in ("median", "lbvh"):
        if build == "median":
            rc, nodes, order, _ = oracle.build_bvh(sph)
            tie = 0
        else:
            nodes, order, _, _ = oracle.build_lbvh(sph, 30)
            tie = 1
        wide = oracle.collapse4(nodes)
        nc = wide["n_children"]
        assert nc.min() >= 2 and nc.max() <= 4 and nc.mean() > 3.0 and wide.shape[0] < 0.5 * ((nodes.shape[0] - 1) // 2)
        used = np.arange(4)[None, :] < nc[:, None]
        ch = wide["child"]
        leaves = ~ch[used & (ch < 0)]
        assert np.array_equal(np.sort(leaves), np.arange(order.shape[0]))                 # every leaf position exactly once
        inner = ch[used & (ch >= 0)]
        assert np.array_equal(np.sort(inner), np.arange(1, wide.shape[0]))                # every wide node but the root has one parent
        assert np.all(inner > np.repeat(np.arange(wide.shape[0]), (used & (ch >= 0)).sum(1)))   # pre-order: children after parents
        assert np.all(ch[~used] == np.iinfo(np.int32).max)
        # child boxes are the binary nodes' boxes: a leaf child's box is the leaf node's
        leaf_nodes = nodes[nodes["nPrimitives"] > 0]
        lb = {int(o): (a.tobytes(), b.tobytes()) for o, a, b in zip(leaf_nodes["offset"], leaf_nodes["bmin"], leaf_nodes["bmax"])}
        wi, ki = np.nonzero(used & (ch < 0))
        for w, k in list(zip(wi, ki))[::97]:
            assert (wide["bmin"][w, k].tobytes(), wide["bmax"][w, k].tobytes()) == lb[int(~ch[w, k])]
        rng = np.random.default_rng(5)
        m = 20000
        tgt = sph[rng.integers(0, sph.shape[0] - 1, m), :3] + rng.normal(size=(m, 3)).astype(np.float32) * np.float32(0.07)
        d = (tgt / np.linalg.norm(tgt, axis=1, keepdims=True)).astype(np.float32)
        o = np.zeros((1, 3), np.float32)
        h2, t2, cand = oracle.trace(sph, nodes, order, o, d, tie_by_objid=tie)
        h4, t4, tests4 = oracle.trace_wide4(sph, wide, order, o, d, tie_by_objid=tie)
        assert np.array_equal(h2, h4) and t2.tobytes() == t4.tobytes()
        assert (h2 >= 0).mean() > 0.5
        print("%s: %d binary nodes -> %d wide nodes (%.2f children each), %.1f box tests per ray on the wide tree" %
              (build, nodes.shape[0], wide.shape[0], nc.mean(), tests4 / m))
