set -x
mkdir -p gpurun_out
export RTDS_TEST_EXPERIMENTAL=1
timeout 900 python -m pytest tests/test_gpu_render.py -q -m gpu -k "two_pixel or quad" > gpurun_out/r02a_exp_tests.log 2>&1
tail -5 gpurun_out/r02a_exp_tests.log
AB_DEVICE=1 timeout 300 python tools/ab_render.py RTDS_PACKET2=0,1 > gpurun_out/r02a_ab_packet2_c3.txt 2>&1
AB_DEVICE=1 AB_CLONES=1 AB_W=640 AB_H=480 AB_SPP=1 timeout 300 python tools/ab_render.py RTDS_QUAD=0,1 > gpurun_out/r02a_ab_quad_c1.txt 2>&1
AB_DEVICE=1 AB_CLONES=1 AB_W=1920 AB_H=1080 AB_SPP=1 timeout 300 python tools/ab_render.py RTDS_QUAD=0,1 > gpurun_out/r02a_ab_quad_c2.txt 2>&1
AB_DEVICE=1 AB_CLONES=1 AB_W=1920 AB_H=1080 AB_SPP=4 timeout 300 python tools/ab_render.py RTDS_PACKET2=0,1 > gpurun_out/r02a_ab_packet2_c2spp4.txt 2>&1
AB_DEVICE=1 AB_SPP=1 timeout 300 python tools/ab_render.py RTDS_QUAD=0,1 > gpurun_out/r02a_ab_quad_c3spp1.txt 2>&1
cat gpurun_out/r02a_ab_*.txt
RTDS_QUAD=1 REPS=2 timeout 600 python tools/run_configs.py 4 gpurun_out/r02a_config4_quad.json > gpurun_out/r02a_config4_quad.log 2>&1
tail -3 gpurun_out/r02a_config4_quad.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 2 -c 1 -f -o gpurun_out/r02a_config4_render python tools/ncu_case.py 4 > gpurun_out/r02a_ncu_config4.log 2>&1
tail -5 gpurun_out/r02a_ncu_config4.log
