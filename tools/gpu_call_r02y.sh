set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_render.py -q -m gpu --timeout 120 -x -k "wide" 2>&1 | tail -15
RTDS_WIDE=1 WORKLOAD=config4 ITERS=10 timeout 600 python tools/ab_frame.py wide=0,1 2>&1 | cut -c1-230
for c in "1 640 480" "1 1920 1080" "30 3840 2160"; do set -- $c; AB_CLONES=$1 AB_W=$2 AB_H=$3 AB_SPP=1 RTDS_WIDE=1 AB_DEVICE=1 timeout 300 python tools/ab_render.py wide=0,1 | head -2; done
