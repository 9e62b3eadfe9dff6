set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_shared_frame.py tests/test_gpu_host_main.py -q -m gpu -rs > gpurun_out/r02g_pytest_2gpu.log 2>&1
tail -8 gpurun_out/r02g_pytest_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02g_bench_config3_2gpu.json 2> gpurun_out/r02g_bench_config3_2gpu.err
tail -c 1800 gpurun_out/r02g_bench_config3_2gpu.json; tail -3 gpurun_out/r02g_bench_config3_2gpu.err
RTDS_FRAME_GRAPH=1 timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02g_bench_config3_2gpu_graph.json 2> gpurun_out/r02g_bench_config3_2gpu_graph.err
tail -c 600 gpurun_out/r02g_bench_config3_2gpu_graph.json
RTDS_LPT=0 timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02g_bench_config3_2gpu_lpt0.json 2> gpurun_out/r02g_bench_config3_2gpu_lpt0.err
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --gather nccl > gpurun_out/r02g_bench_config3_2gpu_nccl.json 2> gpurun_out/r02g_bench_config3_2gpu_nccl.err
timeout 900 $TR bench.py --gpus 2 --workload config4 --steps 5 --warmup 3 > gpurun_out/r02g_bench_config4_2gpu.json 2> gpurun_out/r02g_bench_config4_2gpu.err
tail -c 600 gpurun_out/r02g_bench_config4_2gpu.json; tail -3 gpurun_out/r02g_bench_config4_2gpu.err
timeout 900 $TR bench.py --gpus 2 --workload config5 --steps 3 --warmup 3 > gpurun_out/r02g_bench_config5_2gpu.json 2> gpurun_out/r02g_bench_config5_2gpu.err
tail -c 600 gpurun_out/r02g_bench_config5_2gpu.json; tail -3 gpurun_out/r02g_bench_config5_2gpu.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02g_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'value %.0f ms %.3f | e2e %.0f ms %.3f (via gpu frame %s) | sha %s match %s %s %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], (d['e2e'].get('via_gpu_assembled_frame') or {}).get('ms_per_step'), d['frame_sha256'][:12], d.get('frame_matches_single_rank'), d.get('p2p_matches_single_rank_fresh_jitter'), d.get('nccl_gather_matches_single_rank_fresh_jitter')))
    except Exception as e: print(f, 'ERR', e)
PY
