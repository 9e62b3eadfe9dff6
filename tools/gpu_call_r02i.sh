set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_render.py tests/test_gpu_shared_frame.py -q -m gpu 2>&1 | tail -4
for w in 1 8; do WORLD=$w ITERS=10 timeout 600 python tools/ab_frame.py dirs_ahead=0,1,2,3 >> gpurun_out/r02i_ab_dirs_ahead.txt 2>&1; done
cat gpurun_out/r02i_ab_dirs_ahead.txt | cut -c1-330
