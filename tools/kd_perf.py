import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, __graft_entry__ as g, conftest as T
rt = g.load_rtds(); ctx = rt.Rtds(0)
sph, mat = T.bunny_scene(); ctx.set_spheres(sph, mat)
st = ctx.build(rt.KDTREE); print("KD build %.2f ms, %d launches" % (st["ms"], st["kernel_launches"]))
for kw in ({}, {"kd_closest": 1}):
    ms = [ctx.render(rt.KDTREE, 1920, 1080, 1, **kw)[3] for _ in range(6)]
    s = ms[-1]; print(kw, "kernel %.3f ms, %.2f prim tests/ray, %.1f node visits/ray" % (min(m["ms_kernel"] for m in ms), s["prim_tests"] / s["rays"], s["node_visits"] / s["rays"]))
ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
ms = [ctx.render(rt.LBVH, 1920, 1080, 1)[3] for _ in range(6)]; s = ms[-1]
print("LBVH", "kernel %.3f ms, %.2f prim tests/ray, %.1f node visits/ray" % (min(m["ms_kernel"] for m in ms), s["prim_tests"] / s["rays"], s["node_visits"] / s["rays"]))
