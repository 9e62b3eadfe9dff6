set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR bench.py --gpus 8 --steps 10 --warmup 10 > gpurun_out/r02F_bench_config3_8gpu.json 2> gpurun_out/r02F_bench_config3_8gpu.err
tail -c 900 gpurun_out/r02F_bench_config3_8gpu.json; tail -3 gpurun_out/r02F_bench_config3_8gpu.err
timeout 900 $TR bench.py --gpus 8 --workload config4 --steps 5 --warmup 10 > gpurun_out/r02F_bench_config4_8gpu.json 2> gpurun_out/r02F_bench_config4_8gpu.err
tail -c 600 gpurun_out/r02F_bench_config4_8gpu.json; tail -3 gpurun_out/r02F_bench_config4_8gpu.err
timeout 900 $TR bench.py --gpus 8 --workload config5 --steps 5 --warmup 10 > gpurun_out/r02F_bench_config5_8gpu.json 2> gpurun_out/r02F_bench_config5_8gpu.err
tail -c 600 gpurun_out/r02F_bench_config5_8gpu.json; tail -3 gpurun_out/r02F_bench_config5_8gpu.err
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542"
timeout 600 $TR4 bench.py --gpus 4 --steps 10 --warmup 10 > gpurun_out/r02F_bench_config3_4gpu.json 2> gpurun_out/r02F_bench_config3_4gpu.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02F_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'value %.0f ms %.3f | e2e %.0f ms %.3f (via gpu frame %s) | sha %s match %s %s %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], (d['e2e'].get('via_gpu_assembled_frame') or {}).get('ms_per_step'), d['frame_sha256'][:12], d.get('frame_matches_single_rank'), d.get('p2p_matches_single_rank_fresh_jitter'), d.get('nccl_gather_matches_single_rank_fresh_jitter')))
    except Exception as e: print(f, 'ERR', e)
PY
