set -x
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_sah.py tests/test_gpu_triangles.py -q -m gpu --timeout 60 -x 2>&1 | tail -5
timeout 120 python tools/build_profile.py 30 sah
timeout 400 bash tools/gpu_call_r02o.sh 2>&1 | grep -v "^+" | tail -20
