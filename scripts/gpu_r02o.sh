# strided kernels: resident blocks per SM for short lines (256 rows)
set -x
mkdir -p gpurun_out
timeout 300 python scripts/ab_sweeps.py --shape 256,256,256 --reps 50 bps4=HS2_CHUNK:32 2>&1 | grep -v "^{" | tail -2
HS2_STRIDED_BPS=2 timeout 300 python scripts/ab_sweeps.py --shape 256,256,256 --reps 50 bps2=HS2_CHUNK:32 2>&1 | grep -v "^{" | tail -2
timeout 300 python scripts/ab_sweeps.py --shape 256,256,256 --reps 50 m16=HS2_CHUNK:16 2>&1 | grep -v "^{" | tail -2
timeout 300 python scripts/ab_sweeps.py --shape 256,512,512 --problem composite --reps 20 bps4=HS2_CHUNK:32 2>&1 | grep -v "^{" | tail -2
HS2_STRIDED_BPS=2 timeout 300 python scripts/ab_sweeps.py --shape 256,512,512 --problem composite --reps 20 bps2=HS2_CHUNK:32 2>&1 | grep -v "^{" | tail -2
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
