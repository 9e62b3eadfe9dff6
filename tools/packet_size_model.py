"""CPU model (tools/perf_model.cpp): interior visits / hull tests per RAY of the ordered packet traversal on the bench frame for packets of
1 pixel x 4 spp (the current kernel), 2x1 and 2x2 pixels x 4 spp, over the binary and the 4-wide tree. Design evidence, DESIGN.md 10.
`python tools/packet_size_model.py spp1` : the 1-spp case instead (bunny 1080p, or CLONES=30 at 4K): single rays vs 2x2 / 4x1 pixel packets."""
import os, sys, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests')); import conftest as T
rt = T.rtds_b200; oracle = T.Oracle()
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); from perf_model import PerfModel; model = PerfModel()
if len(sys.argv) > 1 and sys.argv[1] == "spp1":
    clones = int(os.environ.get("CLONES", "1"))
    sph, mat = rt.scene_from_vertices(T.bunny_vertices(), clones)
    nodes, order, _, _ = oracle.build_lbvh(sph, 30); wide = oracle.collapse4(nodes)
    W, H = (1920, 1080) if clones == 1 else (3840, 2160)
    tot = {}
    for y in range(20, H - 2, 64):
        _, _, _, dirs = oracle.render_rows(sph, mat, nodes, order, W, H, 1, y, y + 2, tie_by_objid=1, want_dirs=True)
        d = dirs.reshape(2, W, 3)
        single = np.repeat(d.reshape(-1, 1, 3), 4, axis=1)                      # one ray per "packet"
        quad = d.reshape(2, W // 2, 2, 3).transpose(1, 0, 2, 3).reshape(-1, 4, 3)  # 2x2 pixels per packet
        row4 = d.reshape(2, W // 4, 4, 3).reshape(-1, 4, 3)                     # 4x1 pixels per packet
        for name, pk, rays_per in (("single ray", single, 1), ("2x2 pixels", quad, 4), ("4x1 pixels", row4, 4)):
            for uw in (False, True):
                _, st = model.packet_model(sph, nodes, wide, order, pk, use_wide=uw)
                k = (name, uw); a = tot.setdefault(k, [0, 0, 0, 0])
                a[0] += st["packets"] * rays_per; a[1] += st["interior_visits"]; a[2] += st["box_tests"]; a[3] += st["leaf_visits"]
    for (name, uw), (rays, iv, bt, lv) in tot.items():
        print("%-11s %-6s per RAY: %.2f interior visits, %.2f hull tests, %.2f leaf visits" % (name, "wide4" if uw else "binary", iv / rays, bt / rays, lv / rays))
    sys.exit(0)
sph, mat = rt.scene_from_vertices(T.bunny_vertices(), 30, clone_shift=int(os.environ.get("CLONE_SHIFT", "20")))
nodes, order, _, _ = oracle.build_lbvh(sph, 30); wide = oracle.collapse4(nodes)
W, H, SPP = 3840, 2160, 4
tot = {}
for y in range(40, H - 2, 128):
    _, _, _, dirs = oracle.render_rows(sph, mat, nodes, order, W, H, SPP, y, y + 2, tie_by_objid=1, want_dirs=True)
    d = dirs.reshape(2, W, 4, 3)
    cases = {"1 pixel x 4 spp (now)": d.reshape(-1, 4, 3),
             "2x1 pixels x 4 spp": d.reshape(2, W // 2, 8, 3).reshape(-1, 8, 3),
             "2x2 pixels x 4 spp": d.reshape(2, W // 2, 2, 4, 3).transpose(1, 0, 2, 3, 4).reshape(-1, 16, 3)}
    for name, pk in cases.items():
        for uw in (False, True):
            _, st = model.packet_model(sph, nodes, wide, order, pk, use_wide=uw)
            a = tot.setdefault((name, uw), [0, 0, 0, 0]); nr = pk.shape[1]
            a[0] += st["packets"] * nr; a[1] += st["interior_visits"]; a[2] += st["box_tests"]; a[3] += st["leaf_visits"]
for (name, uw), (rays, iv, bt, lv) in tot.items():
    print("%-22s %-6s per RAY: %.2f interior visits, %.2f hull tests, %.3f leaf visits" % (name, "wide4" if uw else "binary", iv / rays, bt / rays, lv / rays))
