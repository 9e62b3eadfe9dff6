set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_errors_edges.py -q -m gpu 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02n_bench_config3_2gpu.json 2> gpurun_out/r02n_bench_config3_2gpu.err
tail -4 gpurun_out/r02n_bench_config3_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02n_bench_config3_2gpu.json').read().strip().splitlines()[-1])
print('value %.0f ms %.3f | e2e %.0f ms %.3f | sliced %s | via gpu frame %s | match %s sliced match %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('scene_exchanged_over_nvlink'), (d['e2e'].get('via_gpu_assembled_frame') or {}).get('ms_per_step'), d.get('frame_matches_single_rank'), d.get('e2e_sliced_frame_matches')))
PY
