set -x
timeout 600 python -m pytest tests/test_gpu_render.py -q -m gpu --timeout 120 -x -k "grazing" 2>&1 | tail -15
