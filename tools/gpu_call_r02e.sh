set -x
mkdir -p gpurun_out
export RTDS_LIB=$PWD/raytracer-data-structures_b200/variants/librtds_bt.so
for r in 5 0; do WORLD=8 RANK=$r timeout 300 python tools/block_timeline.py >> gpurun_out/r02e_block_timeline.txt 2>&1; done
WORLD=1 RANK=0 timeout 300 python tools/block_timeline.py >> gpurun_out/r02e_block_timeline.txt 2>&1
cat gpurun_out/r02e_block_timeline.txt
unset RTDS_LIB
timeout 600 python -m pytest tests/test_gpu_materials.py tests/test_gpu_render.py -q -m gpu 2>&1 | tail -3
