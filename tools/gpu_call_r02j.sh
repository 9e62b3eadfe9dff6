set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_render.py tests/test_gpu_shared_frame.py tests/test_gpu_materials.py -q -m gpu 2>&1 | tail -4
export RTDS_LIB=$PWD/raytracer-data-structures_b200/variants/librtds_bt.so
for l in 0 1; do LPT=$l WORLD=8 RANK=5 timeout 300 python tools/block_timeline.py >> gpurun_out/r02j_block_timeline_lpt.txt 2>&1; done
for l in 0 1; do LPT=$l WORLD=1 RANK=0 timeout 300 python tools/block_timeline.py >> gpurun_out/r02j_block_timeline_lpt.txt 2>&1; done
unset RTDS_LIB
cut -c1-500 gpurun_out/r02j_block_timeline_lpt.txt | grep -v "^blocks per SM"
for w in 1 2 4 8; do WORLD=$w ITERS=10 timeout 600 python tools/ab_frame.py lpt=0,1 >> gpurun_out/r02j_ab_lpt.txt 2>&1; done
WORKLOAD=config4 ITERS=8 timeout 600 python tools/ab_frame.py lpt=0,1 >> gpurun_out/r02j_ab_lpt.txt 2>&1
WORKLOAD=config5 ITERS=5 timeout 600 python tools/ab_frame.py lpt=0,1 >> gpurun_out/r02j_ab_lpt.txt 2>&1
cut -c1-330 gpurun_out/r02j_ab_lpt.txt
