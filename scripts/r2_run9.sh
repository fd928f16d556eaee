set -x
timeout 900 python -m pytest tests/test_cli.py tests/test_gpu_pipeline.py -m gpu -x -q -k "fused or allreduce or interleave or failed_batch or intervals" 2>&1 | tail -40
