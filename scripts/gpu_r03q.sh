# z_forward compiled for three resident blocks per SM (80 registers) against two (128): timing on one GPU + slab parity
set -x
HS2_ZF_BPS=2 python scripts/slab_bench.py 2 0 10 2>&1 | tail -1
HS2_ZF_BPS=3 python scripts/slab_bench.py 2 0 10 2>&1 | tail -1
HS2_ZF_BPS=3 timeout 60 python -m pytest tests/test_gpu_more.py -x -q -m gpu -k "slab_kernels_on_one_gpu" 2>&1 | tail -2
