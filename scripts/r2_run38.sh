# round 2, GPU call 38: the index builder's sort / scan / selection primitives against numpy
set -x
timeout 900 python -m pytest tests/test_gpu_sortscan.py -m gpu -x -q 2>&1 | tail -8
