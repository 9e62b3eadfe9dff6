set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_render.py tests/test_gpu_shared_frame.py -q -m gpu 2>&1 | tail -4
for w in 1 8; do WORLD=$w ITERS=10 timeout 600 python tools/ab_frame.py lpt=0,1,2 >> gpurun_out/r02l_ab_lpt.txt 2>&1; done
WORKLOAD=config4 ITERS=8 timeout 600 python tools/ab_frame.py lpt=0,1,2 >> gpurun_out/r02l_ab_lpt.txt 2>&1
cut -c1-330 gpurun_out/r02l_ab_lpt.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; tail -2 gpurun_out/r02l_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02l_bench.json').read().strip().splitlines()[-1])
print('value %.0f (%.4f ms) e2e %.0f (%.4f ms) kernel %.4f shadows %.4f ms match %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['with_shadows']['ms_per_step'], d['frame_matches_single_rank']))
PY
