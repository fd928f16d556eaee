# round 2, GPU call 49: other k-mer lengths through the GPU pipeline
set -x
timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q -k "other_kmer" 2>&1 | tail -12
