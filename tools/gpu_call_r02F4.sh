# round 2 final: the whole GPU suite + sanitizers on the last build (wavefront + wide option in)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/r02F_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02F_pytest_gpu.log
timeout 120 python tools/sanitize_smoke.py 2>&1 | tail -3
SAN_TIMEOUT=400 bash tools/sanitize.sh 2>&1 | tail -12
python - <<'PY'
import sys; sys.path.insert(0,'tests')
import conftest as T
rt=T.rtds_b200
ctx=rt.Rtds(0)
sph,mat=T.bunny_scene(30)
ctx.set_spheres(sph,mat)
for w in (0,1,0,1):
    ctx.set_option("wide", w)
    st=[ctx.build(rt.LBVH, mode=rt.MODE_TRUE)["ms"] for _ in range(5)]
    print("wide", w, "LBVH build ms", ["%.3f"%x for x in st])
PY
