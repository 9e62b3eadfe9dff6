set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sah.py tests/test_gpu_triangles.py tests/test_gpu_errors_edges.py -q -m gpu -s 2>&1 | grep -v "^$" | tail -8
timeout 300 python tools/build_profile.py 30 sah
timeout 600 python /dev/stdin <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tools'))
import run_configs as RC
rt = RC.rt
ctx = rt.Rtds(0)
sph, mat = RC.torus_knot_scene(7_000_000)
ctx.set_spheres(sph, mat)
for i in range(3):
    st = ctx.build(rt.BVH, mode=rt.MODE_SAH)
    print("SAH 7M: %.2f ms (%.2f ms/Mprim), %d launches, depth %d" % (st["ms"], st["ms"] / 7.0, st["kernel_launches"], st["max_depth"]))
import hashlib
nodes, order = ctx.export_bvh()
print("tree sha", hashlib.sha256(nodes.tobytes() + order.tobytes()).hexdigest()[:16])
PY
