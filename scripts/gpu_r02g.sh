set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q --durations=5 2>&1 | tail -15
