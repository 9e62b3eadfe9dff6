set -x
mkdir -p gpurun_out
export RTDS_LIB=$PWD/raytracer-data-structures_b200/variants/librtds_bt.so
for o in 2 3; do for r in 5 0; do ORDER=$o WORLD=8 RANK=$r timeout 300 python tools/block_timeline.py >> gpurun_out/r02f_block_timeline.txt 2>&1; done; done
for o in 2 3; do ORDER=$o WORLD=1 RANK=0 timeout 300 python tools/block_timeline.py >> gpurun_out/r02f_block_timeline.txt 2>&1; done
unset RTDS_LIB
cat gpurun_out/r02f_block_timeline.txt | cut -c1-600
for w in 8 4 2 1; do WORLD=$w ITERS=8 timeout 600 python tools/ab_frame.py block_order=2,3 >> gpurun_out/r02f_ab_order.txt 2>&1; done
cat gpurun_out/r02f_ab_order.txt
