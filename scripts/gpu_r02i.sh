# two-lines-per-lane x kernel: parity of every shape, A/B
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sizes.py -x -q -k "warp_x_kernel_shapes and 23" 2>&1 | tail -12
timeout 300 python scripts/ab_sweeps.py --reps 20 w42=HS2_XW_SHAPE:42 w24=HS2_XW_SHAPE:24 2>&1 | grep -v "^{" | tail -4
timeout 300 python scripts/ab_sweeps.py --shape 512,512,512 --problem steelonwater --reps 20 w42=HS2_XW_SHAPE:42 w24=HS2_XW_SHAPE:24 2>&1 | grep -v "^{" | tail -4
timeout 300 python scripts/ab_sweeps.py --shape 256,512,512 --problem composite --reps 20 w42=HS2_XW_SHAPE:42 w24=HS2_XW_SHAPE:24 2>&1 | grep -v "^{" | tail -4
timeout 300 python scripts/ab_sweeps.py --reps 20 w23=HS2_XW_SHAPE:23 2>&1 | grep -v "^{" | tail -3
