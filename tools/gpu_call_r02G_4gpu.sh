# A/B on a real 4-GPU box: heavy-tile split off / on, learned order off
set -x
mkdir -p gpurun_out
N=4
for v in "RTDS_LPT_SPLIT=0" "RTDS_LPT_SPLIT=128" "RTDS_LPT=0" "RTDS_LPT_SPLIT=0" "RTDS_LPT_SPLIT=128"; do
  env $v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --steps 20 --warmup 10 > gpurun_out/r02G_ab.json 2> gpurun_out/r02G_ab.err
  python - "$v" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r02G_ab.json').read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.0f ms %.4f' % (d['value'], d['ms_per_step']), 'kernel', d['per_rank']['render_kernel_ms'], 'total', d['per_rank']['device_total_ms'], 'launches', d['gpu_launches'], 'match', d['frame_matches_single_rank'])
PY
done
