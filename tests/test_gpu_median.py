"""GPU parity: the median-split BVH builder (K6) against constructBVHNew — bit-exact topology, AABB bit patterns
and primitive order, ties and dropped ranges included — via the oracle port (same libstdc++ algorithms) and the
sha256 of the unmodified reference's trees (tests/golden/golden.json), then the default-config image."""
import hashlib
import os
import sys

import numpy as np
import pytest

import conftest as T

sys.path.insert(0, T.GOLDEN)
import make_golden  # noqa: E402  (scene generators only)

rt = T.rtds_b200
G = T.load_golden_json()
pytestmark = pytest.mark.gpu


def tree_sha(nodes, order):
    return hashlib.sha256(nodes.tobytes() + np.ascontiguousarray(order, np.int32).tobytes()).hexdigest()


def _check_against_oracle(gpu_ctx, oracle, sph, mat, acc=rt.BVH):
    n_use = sph.shape[0] - (1 if acc == rt.LBVH else 0)
    gpu_ctx.set_spheres(sph, mat)
    st = gpu_ctx.build(acc)
    rc, nodes_o, order_o, depth_o = oracle.build_bvh(sph, n_use)
    assert rc == 0
    nodes, order = gpu_ctx.export_bvh()
    assert st["total_nodes"] == nodes_o.shape[0]
    assert np.array_equal(order, order_o), f"primitive order differs at {np.count_nonzero(order != order_o)} of {order.size} leaves"
    assert nodes.tobytes() == nodes_o.tobytes()
    assert st["max_depth"] == depth_o
    return st, nodes, order


def test_bvh_bunny_matches_reference_tree(gpu_ctx, oracle):
    sph, mat = T.bunny_scene()
    st, nodes, order = _check_against_oracle(gpu_ctx, oracle, sph, mat)
    e = G["default_config"]["BVH"]
    assert st["total_nodes"] == e["total_nodes"] == 71895
    assert tree_sha(nodes, order) == e["tree_sha256"]                      # the UNMODIFIED reference's tree
    assert np.array_equal(order, np.load(T.GOLDEN + "/bunny_bvh_prim_order.npy"))


def test_lbvh_compat_bunny_matches_reference_tree(gpu_ctx, oracle):
    sph, mat = T.bunny_scene()
    st, nodes, order = _check_against_oracle(gpu_ctx, oracle, sph, mat, rt.LBVH)
    e = G["default_config"]["LBVH"]
    assert st["total_nodes"] == e["total_nodes"] == 71893 and tree_sha(nodes, order) == e["tree_sha256"]


def _golden_scene(name):
    if name == "bunny_clones3":
        return T.bunny_scene(3)
    if name.startswith("synthetic_"):
        _, n, seed = name.split("_")
        return T.synthetic_scene(int(n) if int(n) != 2 else 1, int(seed[4:]))
    _, n, seed, q = name.split("_")
    return make_golden.quantised_scene(int(n), int(seed[4:]), float(q[1:]))


@pytest.mark.parametrize("name", [k for k in G["trees"] if k != "armadillo"])
def test_bvh_golden_scenes(gpu_ctx, oracle, name):
    """Includes the tie-heavy quantised scenes where the reference DROPS ranges (fewer than 2n-1 nodes)."""
    sph, mat = _golden_scene(name)
    e = G["trees"][name]
    st, nodes, order = _check_against_oracle(gpu_ctx, oracle, sph, mat)
    assert st["total_nodes"] == e["BVH"]["total_nodes"] and order.size == e["BVH"]["n_leaves"]
    assert tree_sha(nodes, order) == e["BVH"]["tree_sha256"]
    if "LBVH" in e:
        st, nodes, order = _check_against_oracle(gpu_ctx, oracle, sph, mat, rt.LBVH)
        assert tree_sha(nodes, order) == e["LBVH"]["tree_sha256"]


@pytest.mark.parametrize("small", ["1", "8", "64", "100000000"])
def test_bvh_tie_fuzz_all_code_paths(gpu_ctx, oracle, small):
    """Random sizes x tie densities, with the sequential/parallel switch-over forced to every regime:
    small=1 -> every node by the block-parallel Hoare emulation; huge -> one thread runs the literal algorithms."""
    gpu_ctx.set_option("median_small", int(small))
    try:
        rng = np.random.default_rng(int(small) % 1000 + 7)
        for trial in range(40):
            n = int(rng.choice([2, 3, 4, 5, 7, 16, 33, 100, 257, 1000, 2049, 4097, 6000]))
            q = float(rng.choice([0.0, 0.25, 1.0, 3.0]))
            sph, mat = T.synthetic_scene(n, int(rng.integers(1 << 30)), ground=bool(rng.integers(2)))
            if q > 0:
                sph[:n, :3] = np.round(sph[:n, :3] / np.float32(q)) * np.float32(q)
            rc = oracle.build_bvh(sph)[0]
            if rc != 0:
                continue
            _check_against_oracle(gpu_ctx, oracle, sph, mat)
    finally:
        gpu_ctx.set_option("median_small", 0)


def test_bvh_degenerate_input_is_an_error(gpu_ctx, oracle):
    """std::partition returns endIndex -> the reference recurses forever (accelerators.h:311-330); the library
    reports RTDS_ERR_DEGENERATE instead."""
    c = np.float32(174.0289764404297)
    sph = np.zeros((4, 4), np.float32)
    sph[0] = [c, 0, 0, 100]
    for i in range(1, 4):
        sph[i] = [c - np.float32(i), np.float32(i) * 0.5, -np.float32(i) * 0.25, 0.05]
    assert oracle.build_bvh(sph)[0] == -6
    gpu_ctx.set_spheres(sph, None)
    with pytest.raises(rt.RtdsError) as e:
        gpu_ctx.build(rt.BVH)
    assert e.value.code == -6


def test_bvh_million_prims_bit_exact(gpu_ctx, oracle):
    """BASELINE config 3 scale: 30 bunny clones, 1,078,411 prims, 2,156,821 nodes."""
    sph, mat = T.bunny_scene(30)
    st, nodes, order = _check_against_oracle(gpu_ctx, oracle, sph, mat)
    assert st["total_nodes"] == 2 * sph.shape[0] - 1 == 2156821
    print("median-split BVH build 1,078,411 prims: %.3f ms (%.3f ms/Mprim), %d launches" % (st["ms"], st["ms"] / (sph.shape[0] / 1e6), st["kernel_launches"]))


@pytest.mark.parametrize("exact", [True, False])
def test_render_default_config_equals_shipped_output_ppm(gpu_ctx, exact):
    """settings.h defaults (640x480, aa_samples 1, dataStructure BVH, bunny): the GPU frame is byte-identical to
    the reference's shipped output.ppm (md5 c69c6637...), and the hit ids equal the reference BVH path's."""
    sph, mat = T.bunny_scene()
    gpu_ctx.set_spheres(sph, mat)
    gpu_ctx.build(rt.BVH)
    rgb, hit, _, st = gpu_ctx.render(rt.BVH, 640, 480, 1, want_hit=True, exact=exact)
    gold = np.load(T.GOLDEN + "/bunny_hits_640x480.npz")
    assert np.array_equal(hit, gold["hit_bvh"])
    assert T.ppm_md5(rgb) == G["default_config"]["BVH"]["ppm_md5"] == "c69c66375f2c6bda433f9f457a4b2b2e"
    if exact:
        assert st["prim_tests"] == G["default_config"]["BVH"]["candidates"] == 323685   # = the reference's printed test count


def test_render_lbvh_compat_equals_reference_lbvh_ppm(gpu_ctx):
    sph, mat = T.bunny_scene()
    gpu_ctx.set_spheres(sph, mat)
    gpu_ctx.build(rt.LBVH)                     # compat: the reference's LBVH drops the ground sphere
    for exact in (True, False):
        rgb, _, _, st = gpu_ctx.render(rt.LBVH, 640, 480, 1, exact=exact)
        assert T.ppm_md5(rgb) == G["default_config"]["LBVH"]["ppm_md5"]
        if exact:
            assert st["prim_tests"] == 93644   # Report/Performance.xlsx row 7-8 col F (and today's binary)
