set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_materials.py tests/test_gpu_render.py tests/test_gpu_shared_frame.py -q -m gpu --timeout 120 -x 2>&1 | tail -5
WORKLOAD=config5 ITERS=5 timeout 600 python tools/ab_frame.py shadow_packets=0,1 > gpurun_out/r02v_ab_shadow_packets_c5.txt 2>&1
cut -c1-330 gpurun_out/r02v_ab_shadow_packets_c5.txt
AB_DEVICE=1 timeout 300 python tools/ab_render.py shadow_packets=0,1 shadows > gpurun_out/r02v_ab_shadow_packets_c3.txt 2>&1
cat gpurun_out/r02v_ab_shadow_packets_c3.txt
