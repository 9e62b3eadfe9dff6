set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/r02t_pytest_gpu.log 2>&1
tail -5 gpurun_out/r02t_pytest_gpu.log
timeout 120 python tools/sanitize_smoke.py 2>&1 | tail -3
SAN_TIMEOUT=400 bash tools/sanitize.sh 2>&1 | tail -20
