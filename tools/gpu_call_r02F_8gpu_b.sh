# round 2 final: config 3 at 8, 4 and 2 GPUs of one box (the side measurements now have their own warm-up)
set -x
mkdir -p gpurun_out
for N in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29550+N)) bench.py --gpus $N --steps 10 --warmup 10 > gpurun_out/r02F_bench_config3_${N}gpu.json 2> gpurun_out/r02F_bench_config3_${N}gpu.err
  tail -c 300 gpurun_out/r02F_bench_config3_${N}gpu.json; tail -2 gpurun_out/r02F_bench_config3_${N}gpu.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02F_bench_config3_?gpu.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'value %.0f ms %.3f | e2e %.0f ms %.3f %s | shadows %s | sha %s match %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['variant'][:30], d['with_shadows'] and round(d['with_shadows']['value']), d['frame_sha256'][:12], d.get('frame_matches_single_rank')))
    except Exception as e: print(f, 'ERR', e)
PY
