"""Times every builder on the bench scene (bunny x30 clones, 1,078,411 spheres); run under ncu for a launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
rt = g.load_rtds()
v = np.fromfile(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "bunny_vertices.f32"), np.float32).reshape(-1, 3)
clones = int(sys.argv[1]) if len(sys.argv) > 1 else 30
which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["lbvh", "median", "sah", "kd"]
sph, mat = rt.scene_from_vertices(v, clones)
ctx = rt.Rtds(0)
ctx.set_spheres(sph, mat)
n = sph.shape[0]
for name in which:
    reps = 3
    for r in range(reps):
        if name == "lbvh": st = ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
        elif name == "lbvh63": st = ctx.build(rt.LBVH, mode=rt.MODE_TRUE, morton_bits=63)
        elif name == "median": st = ctx.build(rt.BVH)
        elif name == "sah": st = ctx.build(rt.BVH, mode=rt.MODE_SAH)
        elif name == "kd": st = ctx.build(rt.KDTREE)
    print("%-7s n=%d: %.3f ms (%.3f ms/Mprim), %d launches, depth %d, nodes %d" % (name, n, st["ms"], st["ms"] / (n / 1e6), st["kernel_launches"], st["max_depth"], st["total_nodes"]), flush=True)
