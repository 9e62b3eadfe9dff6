set -x
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/r02F_pytest_gpu.log 2>&1
tail -2 gpurun_out/r02F_pytest_gpu.log
timeout 600 python bench.py --workload config5 --steps 5 --warmup 3 > gpurun_out/r02F_bench_config5_1gpu.json 2> gpurun_out/r02F_bench_config5_1gpu.err; tail -c 400 gpurun_out/r02F_bench_config5_1gpu.json | head -c 100
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02F_bench_config5_1gpu.json').read().strip().splitlines()[-1]); print('config5', d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d['frame_sha256'][:12])
PY
