# warp x kernel: restored fast path (LPW) + two warps per line (nx = 1024)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tail -12
timeout 300 python scripts/ab_sweeps.py --reps 20 warp=HS2_X_KERNEL:warp 2>&1 | grep -v "^{" | tail -3
timeout 300 python scripts/ab_sweeps.py --shape 256,256,256 --reps 50 warp=HS2_X_KERNEL:warp 2>&1 | grep -v "^{" | tail -3
timeout 300 python scripts/ab_sweeps.py --shape 128,1024,1024 --reps 20 warp13=HS2_XW_SHAPE:13 warp12=HS2_XW_SHAPE:12 fold=HS2_X_KERNEL:fold 2>&1 | grep -v "^{" | tail -4
timeout 300 python scripts/ab_sweeps.py --shape 512,512,512 --problem steelonwater --reps 20 warp=HS2_X_KERNEL:warp tma=HS2_X_KERNEL:tma 2>&1 | grep -v "^{" | tail -3
