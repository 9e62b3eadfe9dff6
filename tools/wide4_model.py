"""CPU model (tools/perf_model.cpp): dependent node visits of the ordered packet traversal on the BENCH frame
(bunny x30 clones, 3840x2160, 4 spp, LBVH) over the binary tree and over its 4-wide collapse - design evidence for
DESIGN.md section 10.1. Runs on the CPU only (rows sub-sampled). usage: python tools/wide4_model.py [row_step=48]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest as T  # noqa: E402

rt = T.rtds_b200
step = int(sys.argv[1]) if len(sys.argv) > 1 else 48
clones = int(os.environ.get("CLONES", "30"))
shift = int(os.environ.get("CLONE_SHIFT", "20"))
W, H, SPP = 3840, 2160, 4
oracle = T.Oracle()
sys.path.insert(0, os.path.join(ROOT, "tools"))
from perf_model import PerfModel  # noqa: E402
model = PerfModel()
if os.environ.get("SCENE") == "config4":          # the 7 M-sphere torus-knot stand-in, 1 spp: "packets" of four identical rays
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import run_configs
    sph, mat = run_configs.torus_knot_scene(int(os.environ.get("N_PRIMS", "7000000")))
    SPP = 1
else:
    sph, mat = rt.scene_from_vertices(T.bunny_vertices(), clones, clone_shift=shift)
t0 = time.time()
nodes, order, _, _ = oracle.build_lbvh(sph, 30)
RULE, ORDER = int(os.environ.get("COLLAPSE_RULE", "0")), int(os.environ.get("ORDER_MODE", "0"))
QUANT = int(os.environ.get("QUANT_BITS", "0"))      # wide tree: child boxes quantised outward to this many bits per plane
wide = oracle.collapse4(nodes, RULE)
print("LBVH %d nodes -> %d wide nodes (%.2f children each) in %.1f s; collapse rule %d, child order mode %d" %
      (nodes.shape[0], wide.shape[0], wide["n_children"].mean(), time.time() - t0, RULE, ORDER))
tot = {False: None, True: None}
hits_ok = True
for y in range(step // 2, H, step):
    _, hit_ref, _, dirs = oracle.render_rows(sph, mat, nodes, order, W, H, SPP, y, y + 1, tie_by_objid=1, want_dirs=True)
    d = dirs.reshape(-1, 4, 3) if SPP == 4 else np.repeat(dirs.reshape(-1, 1, 3), 4, axis=1)
    # the exact per-ray hits of all four samples (render_rows reports the last sample's hit only)
    h_exact, _, _ = oracle.trace(sph, nodes, order, np.zeros((1, 3), np.float32), np.ascontiguousarray(d).reshape(-1, 3), tie_by_objid=1)
    for use_wide in (False, True):
        hit, st = model.packet_model(sph, nodes, wide, order, d, use_wide=use_wide, order_mode=ORDER, quant_bits=QUANT if use_wide else 0)
        hits_ok &= bool(np.array_equal(hit.reshape(-1), h_exact))
        tot[use_wide] = st if tot[use_wide] is None else {k: (max(tot[use_wide][k], v) if k == "max_stack" else tot[use_wide][k] + v) for k, v in st.items()}
print("model hits == unpruned reference traversal on every sampled ray:", hits_ok)
for use_wide in (False, True):
    s = tot[use_wide]
    p = s["packets"]
    print("%-7s per packet: %.2f interior visits (dependent node loads), %.2f leaf visits, %.1f hull tests, %.2f prim tests; deepest stack %d; "
          "node bytes %.0f" % ("wide4" if use_wide else "binary", s["interior_visits"] / p, s["leaf_visits"] / p, s["box_tests"] / p,
                               s["prim_tests"] / p, s["max_stack"], s["interior_visits"] / p * ((64 if QUANT == 8 else 128) if use_wide else 64)))
