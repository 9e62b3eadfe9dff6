set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 -x 2>&1 | tail -8 | tee gpurun_out/r02z_pytest_gpu.log
WORKLOAD=config5 ITERS=9 timeout 600 python tools/ab_frame.py wavefront=0,2 > gpurun_out/r02z_ab_wavefront_c5.txt 2>&1
cut -c1-330 gpurun_out/r02z_ab_wavefront_c5.txt
AB_DEVICE=1 timeout 300 python tools/ab_render.py wavefront=0,2 shadows > gpurun_out/r02z_ab_wavefront_c3.txt 2>&1
cat gpurun_out/r02z_ab_wavefront_c3.txt
