"""GPU parity on the reference's other shipped models (Igea 134,345 vertices with 2,897 median ties, dragon, armadillo,
lucy, teapot with duplicate vertices -> dropped ranges, woody): trees bit-exact against the sha256 of the UNMODIFIED
reference's trees (tests/golden/model_trees.json) and against the oracle port; frames against the oracle port.
The sphere tables are derived data under oracle/_ref/models (written by oracle/dump_models.py in the build container)."""
import hashlib
import json
import os

import numpy as np
import pytest

import conftest as T

rt = T.rtds_b200
pytestmark = pytest.mark.gpu
MODELS_DIR = os.path.join(T.ROOT, "oracle", "_ref", "models")
with open(os.path.join(T.GOLDEN, "model_trees.json")) as f:
    G = json.load(f)


def tree_sha(nodes, order):
    return hashlib.sha256(nodes.tobytes() + np.ascontiguousarray(order, np.int32).tobytes()).hexdigest()


def _load(name):
    path = os.path.join(MODELS_DIR, name + ".f32")
    if not os.path.exists(path):
        pytest.skip(f"{path} not present (derived from /root/reference in the build container)")
    sph = np.fromfile(path, np.float32).reshape(-1, 4)
    assert hashlib.sha256(sph.tobytes()).hexdigest() == G[name]["scene_sha256"]
    mat = np.zeros_like(sph)
    mat[:-1, 0], mat[:-1, 1] = 0.8, 0.7
    return sph, mat


@pytest.mark.parametrize("name", sorted(G))
def test_model_trees_equal_reference(gpu_ctx, oracle, name):
    sph, mat = _load(name)
    gpu_ctx.set_spheres(sph, mat)
    for acc, key, n_use in ((rt.BVH, "BVH", sph.shape[0]), (rt.LBVH, "LBVH", sph.shape[0] - 1)):
        e = G[name][key]
        st = gpu_ctx.build(acc)
        nodes, order = gpu_ctx.export_bvh()
        assert st["total_nodes"] == e["total_nodes"] and order.size == e["n_leaves"]
        assert tree_sha(nodes, order) == e["tree_sha256"], f"{name} {key}: tree differs from the reference's"
        rc, nodes_o, order_o, _ = oracle.build_bvh(sph, n_use)
        assert rc == 0 and nodes.tobytes() == nodes_o.tobytes() and np.array_equal(order, order_o)
        print("%s %s: %d nodes in %.3f ms (reference %.1f ms on one core)" % (name, key, st["total_nodes"], st["ms"], 1e3 * e["ref_build_s"]))


@pytest.mark.parametrize("name", ["Igea", "dragon", "teapot"])
def test_model_frames_equal_oracle(gpu_ctx, oracle, name):
    sph, mat = _load(name)
    gpu_ctx.set_spheres(sph, mat)
    gpu_ctx.build(rt.BVH)
    nodes, order = gpu_ctx.export_bvh()
    W, H = 320, 240
    rgb_o, hit_o, accum_o, _ = oracle.render_rows(sph, mat, nodes, order, W, H, 1, want_accum=True)
    for exact in (True, False):
        rgb, hit, accum, st = gpu_ctx.render(rt.BVH, W, H, 1, want_hit=True, want_accum=True, exact=exact)
        assert np.array_equal(hit, hit_o) and accum.tobytes() == accum_o.tobytes() and np.array_equal(rgb, rgb_o)
    assert (hit_o >= 0).mean() > 0.1
