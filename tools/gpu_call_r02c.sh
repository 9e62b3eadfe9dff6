set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_render.py tests/test_gpu_shared_frame.py tests/test_gpu_host_main.py tests/test_gpu_bench_parity.py tests/test_gpu_triangles.py tests/test_gpu_materials.py tests/test_gpu_errors_edges.py -q -m gpu > gpurun_out/r02c_pytest.log 2>&1
tail -15 gpurun_out/r02c_pytest.log
mkdir -p /tmp/cold/models && python - <<'PY'
import numpy as np
v=np.fromfile('tests/golden/bunny_vertices.f32',np.float32).reshape(-1,3)
open('/tmp/cold/models/bunny.obj','w').write(''.join('v %.9g %.9g %.9g\n'%tuple(r) for r in v))
PY
for i in 1 2; do (cd /tmp/cold && RTDS_TIMING=1 RTDS_MODELS_DIR=/tmp/cold/models RTDS_OUT=/tmp/cold/out.ppm $OLDPWD/raytracer-data-structures_b200/rtds_main > /dev/null 2>> $OLDPWD/gpurun_out/r02c_cold_main.log); done
cat gpurun_out/r02c_cold_main.log
timeout 600 python tools/ab_frame.py frame_graph=0,1 l2_prefetch=0,1 > gpurun_out/r02c_ab_frame_w1.txt 2>&1
cat gpurun_out/r02c_ab_frame_w1.txt
RTDS_LIB=$PWD/raytracer-data-structures_b200/variants/librtds_ldcs.so timeout 600 python tools/ab_frame.py frame_graph=1 l2_prefetch=0,1 > gpurun_out/r02c_ab_frame_w1_ldcs.txt 2>&1
cat gpurun_out/r02c_ab_frame_w1_ldcs.txt
for tr in 8 16 32; do
WORLD=8 TILE_ROWS=$tr ITERS=8 timeout 900 python tools/ab_frame.py frame_graph=0,1 l2_prefetch=0,1 > gpurun_out/r02c_ab_frame_w8_t$tr.txt 2>&1
cat gpurun_out/r02c_ab_frame_w8_t$tr.txt
done
WORLD=8 ITERS=8 RTDS_LIB=$PWD/raytracer-data-structures_b200/variants/librtds_ldcs.so timeout 600 python tools/ab_frame.py frame_graph=1 l2_prefetch=0,1 > gpurun_out/r02c_ab_frame_w8_ldcs.txt 2>&1
cat gpurun_out/r02c_ab_frame_w8_ldcs.txt
