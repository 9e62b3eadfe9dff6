"""tools/perf_model.cpp (performance model, a dev tool - not the oracle, not the product) still agrees with the oracle's hits."""
import os
import sys

import numpy as np

import conftest as T

sys.path.insert(0, os.path.join(T.ROOT, "tools"))
from perf_model import PerfModel  # noqa: E402


def test_packet_model_hits_and_wide_tree_saves_visits(oracle):
    """The CPU model of the ordered packet traversal (tools/perf_model.cpp, design evidence for DESIGN.md 10.1) returns the unpruned
    traversal's hits over the binary AND the 4-wide tree, and the wide tree needs well under two thirds of the interior visits."""
    model = PerfModel()
    sph, mat = T.bunny_scene()
    nodes, order, _, _ = oracle.build_lbvh(sph, 30)
    wide = oracle.collapse4(nodes)
    W, H, spp = 640, 480, 4
    _, _, _, dirs = oracle.render_rows(sph, mat, nodes, order, W, H, spp, 200, 260, tie_by_objid=1, want_dirs=True)
    d = dirs.reshape(-1, 4, 3)
    h_exact, _, _ = oracle.trace(sph, nodes, order, np.zeros((1, 3), np.float32), d.reshape(-1, 3), tie_by_objid=1)
    hb, sb = model.packet_model(sph, nodes, wide, order, d, use_wide=False)
    hw, sw = model.packet_model(sph, nodes, wide, order, d, use_wide=True)
    assert np.array_equal(hb.reshape(-1), h_exact) and np.array_equal(hw.reshape(-1), h_exact)
    assert sb["packets"] == sw["packets"] > 0.9 * d.shape[0] and abs(sb["prim_tests"] - sw["prim_tests"]) < 0.01 * sb["prim_tests"]   # pruning order differs
    assert sw["interior_visits"] < 0.66 * sb["interior_visits"] and sw["box_tests"] < 1.1 * sb["box_tests"]
    assert (h_exact >= 0).mean() > 0.2
