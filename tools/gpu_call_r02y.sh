set -x
timeout 900 python -m pytest tests/test_gpu_render.py tests/test_gpu_shared_frame.py -q -m gpu --timeout 120 -x 2>&1 | tail -4
for w in 8 4 2 1; do echo "== world $w auto"; WORLD=$w ITERS=16 timeout 600 python tools/ab_frame.py lpt=0,1 2>&1 | cut -c1-230 | tail -2; done
