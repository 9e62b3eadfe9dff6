set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_materials.py tests/test_gpu_render.py tests/test_gpu_shared_frame.py -q -m gpu --timeout 120 -x 2>&1 | tail -12
WORKLOAD=config5 ITERS=9 timeout 600 python tools/ab_frame.py wavefront=0,2,1 > gpurun_out/r02x_ab_wavefront_c5.txt 2>&1
cut -c1-330 gpurun_out/r02x_ab_wavefront_c5.txt
AB_DEVICE=1 timeout 300 python tools/ab_render.py wavefront=0,2,1 shadows > gpurun_out/r02x_ab_wavefront_c3.txt 2>&1
cat gpurun_out/r02x_ab_wavefront_c3.txt
