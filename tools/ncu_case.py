"""One BASELINE config (stand-in) rendered a few times, for an ncu capture of its render kernel (GPU box):
  ncu --set full --clock-control none --import-source on -k regex:render -s 2 -c 1 -o gpurun_out/x python tools/ncu_case.py 4
usage: python tools/ncu_case.py {4|5} [frames]   (scenes from tools/run_configs.py)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import run_configs as RC  # noqa: E402

rt = RC.rt
case = sys.argv[1] if len(sys.argv) > 1 else "4"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = rt.Rtds(0)
if case == "4":
    sph, mat = RC.torus_knot_scene(7_000_000)
    ctx.set_spheres(sph, mat)
    st = ctx.build(rt.BVH, mode=rt.MODE_SAH)
    acc, W, H, spp, sh = rt.BVH, 3840, 2160, 1, 0
else:
    sph, mat = RC.city_trees_scene()
    ctx.set_spheres(sph, mat)
    ctx.set_lights(RC.LIGHTS3)
    st = ctx.build(rt.LBVH, mode=rt.MODE_TRUE)
    acc, W, H, spp, sh = rt.LBVH, 3840, 2160, 16, 1
print("build %.2f ms, %d nodes" % (st["ms"], st["total_nodes"]))
for i in range(frames):
    _, _, _, rs = ctx.render(acc, W, H, spp, shadows=sh)
    print("frame %d: kernel %.3f ms, %d rays, %.2f slab tests / ray, %.3f prim tests / ray, %.2f node visits / ray" %
          (i, rs["ms_kernel"], rs["rays"], rs["node_tests"] / rs["rays"], rs["prim_tests"] / rs["rays"], rs["node_visits"] / rs["rays"]))
