"""GPU parity: Morton codes, onesweep sort, Karras emission, atomic refit, pre-order flattening — against the
oracle port and the reference's golden KATs.  Bit-exact (integer / index / IEEE bit patterns)."""
import numpy as np
import pytest

import conftest as T

rt = T.rtds_b200
G = T.load_golden_json()
pytestmark = pytest.mark.gpu


def test_morton30_kats_and_random(gpu_ctx, oracle):
    m = G["morton"]
    assert gpu_ctx.morton30([[.5, .5, .5]])[0] == m["half"]
    assert gpu_ctx.morton30([[1, 1, 1]])[0] == m["ones"]
    rng = np.random.default_rng(5)
    pts = rng.uniform(-0.1, 1.1, size=(64, 3)).astype(np.float32)
    assert gpu_ctx.morton30(pts).tolist() == m["points_seed5"]
    pts = rng.uniform(-0.5, 1.5, size=(200000, 3)).astype(np.float32)
    assert np.array_equal(gpu_ctx.morton30(pts), oracle.morton30(pts))


@pytest.mark.parametrize("bits", [30, 63])
@pytest.mark.parametrize("scene", ["bunny", "synthetic_20000", "dup_5000", "tiny_2", "tiny_1", "ragged_4097"])
def test_lbvh_true_tree_bit_exact(gpu_ctx, oracle, scene, bits):
    if scene == "bunny":
        sph, mat = T.bunny_scene()
    elif scene == "synthetic_20000":
        sph, mat = T.synthetic_scene(20000, 11)
    elif scene == "dup_5000":          # many identical Morton keys AND identical centres
        sph, mat = T.synthetic_scene(5000, 12)
        sph[:5000, :3] = np.round(sph[:5000, :3])
    elif scene == "tiny_2":
        sph, mat = T.synthetic_scene(1, 1)
    elif scene == "tiny_1":
        sph, mat = T.synthetic_scene(1, 1, ground=False)
    else:
        sph, mat = T.synthetic_scene(4096, 13)   # 4097 prims: one key past a full sort tile
    gpu_ctx.set_spheres(sph, mat)
    st = gpu_ctx.build(rt.LBVH, mode=rt.MODE_TRUE, morton_bits=bits)
    nodes_o, order_o, keys_o, depth_o = oracle.build_lbvh(sph, bits)
    keys, ids = gpu_ctx.export_morton()
    assert np.array_equal(keys, keys_o), "sorted Morton keys differ"
    assert np.array_equal(ids, order_o), "sort is not the stable ascending sort (payload order differs)"
    nodes, order = gpu_ctx.export_bvh()
    assert st["total_nodes"] == nodes_o.shape[0] == 2 * sph.shape[0] - 1
    assert np.array_equal(order, order_o)
    assert nodes.tobytes() == nodes_o.tobytes(), "flattened LBVH differs from the sequential restatement"
    assert st["max_depth"] == depth_o


def test_lbvh_true_large_sort_property(gpu_ctx):
    """1M-prim scale (30 bunny clones = BASELINE config 3): sortedness + permutation + box containment."""
    sph, mat = T.bunny_scene(30)
    gpu_ctx.set_spheres(sph, mat)
    st = gpu_ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
    keys, ids = gpu_ctx.export_morton()
    assert np.all(keys[1:] >= keys[:-1])
    assert np.array_equal(np.sort(ids), np.arange(sph.shape[0]))
    eq = keys[1:] == keys[:-1]
    assert np.all(ids[1:][eq] > ids[:-1][eq]), "equal keys must keep input order (stable)"
    nodes, order = gpu_ctx.export_bvh()
    assert nodes.shape[0] == 2 * sph.shape[0] - 1
    # root box = union of all primitive boxes (exact min/max)
    lo = (sph[:, :3] - sph[:, 3:4]).min(0)
    hi = (sph[:, :3] + sph[:, 3:4]).max(0)
    assert np.array_equal(nodes[0]["bmin"], lo) and np.array_equal(nodes[0]["bmax"], hi)
    # every interior node's box is the union of its children's boxes
    inner = np.nonzero(nodes["nPrimitives"] == 0)[0]
    l, r = inner + 1, nodes["offset"][inner]
    assert np.array_equal(nodes["bmin"][inner], np.minimum(nodes["bmin"][l], nodes["bmin"][r]))
    assert np.array_equal(nodes["bmax"][inner], np.maximum(nodes["bmax"][l], nodes["bmax"][r]))
    leaves = np.nonzero(nodes["nPrimitives"] == 1)[0]
    assert np.array_equal(nodes["offset"][leaves], np.arange(sph.shape[0]))
    print("LBVH build 1,078,411 prims: %.3f ms (%.3f ms/Mprim), %d launches" % (st["ms"], st["ms"] / (sph.shape[0] / 1e6), st["kernel_launches"]))
