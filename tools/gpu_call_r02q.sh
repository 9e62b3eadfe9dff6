set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sah.py tests/test_gpu_triangles.py -q -m gpu -s 2>&1 | grep -v "^$" | tail -6
timeout 300 python tools/build_profile.py 30 sah
bash tools/gpu_call_r02o.sh 2>&1 | grep -v "^+" | tail -20
