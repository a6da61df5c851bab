# warp-per-line x kernel: parity first (bounded), then A/B against the patch kernel
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sizes.py -x -q -k "long_lines or warp or patch or folded" 2>&1 | tail -15
timeout 300 python scripts/ab_sweeps.py --reps 20 warp33=HS2_XW_SHAPE:33 warp52=HS2_XW_SHAPE:52 warp42=HS2_XW_SHAPE:42 tma=HS2_X_KERNEL:tma 2>&1 | grep -v "^{" | tail -8
