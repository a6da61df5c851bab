# warp-per-line x kernel: parity first (bounded), then A/B against the patch kernel
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sizes.py -x -q --timeout 100 --timeout-method=thread -k "long_lines or warp or patch or folded" 2>&1 | tail -15
timeout 300 python scripts/ab_sweeps.py --reps 20 warp=HS2_X_KERNEL:warp tma=HS2_X_KERNEL:tma 2>&1 | grep -v "^{" | tail -8
timeout 300 python scripts/ab_sweeps.py --shape 256,256,256 --reps 50 warp=HS2_X_KERNEL:warp tma=HS2_X_KERNEL:tma 2>&1 | grep -v "^{" | tail -8
timeout 300 python scripts/ab_sweeps.py --shape 256,512,512 --reps 20 warp=HS2_X_KERNEL:warp tma=HS2_X_KERNEL:tma 2>&1 | grep -v "^{" | tail -8
