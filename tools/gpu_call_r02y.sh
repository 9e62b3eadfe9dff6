N=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus $N --steps 10 --warmup 10 > gpurun_out/r02F_bench_config3_${N}gpu.json 2> gpurun_out/r02F_bench_config3_${N}gpu.err
tail -3 gpurun_out/r02F_bench_config3_${N}gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02F_bench_config3_2gpu.json').read().strip().splitlines()[-1])
print('value %.0f ms %.3f | e2e %.0f ms %.3f %s | other %s | match %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['variant'][:40], (d['e2e'].get('scene_exchanged_over_nvlink') or {}).get('ms_per_step'), d['frame_matches_single_rank']))
PY
