set -x
timeout 900 python -m pytest tests/test_gpu_materials.py tests/test_gpu_render.py tests/test_gpu_errors_edges.py -q -m gpu --timeout 120 -x 2>&1 | tail -4
WORKLOAD=config5 ITERS=8 timeout 600 python tools/ab_frame.py wavefront=2 2>&1 | cut -c1-230 | tail -1
AB_DEVICE=1 timeout 300 python tools/ab_render.py wavefront=0 shadows | head -1
