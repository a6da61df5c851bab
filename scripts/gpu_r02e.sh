# A/B: where did the warp x kernel lose 0.56 -> 0.70 ms?  (a) native vs numpy tables, (b) current vs first committed kernel
set -x
mkdir -p gpurun_out
timeout 300 python scripts/ab_sweeps.py --reps 20 native=HS2_TABLES:native numpy=HS2_TABLES:numpy 2>&1 | grep -v "^{" | tail -4
HS2_B200_LIB=$PWD/heatsim2_b200/_ab/libhs2b200_xwold.so timeout 300 python scripts/ab_sweeps.py --reps 20 old_numpy=HS2_TABLES:numpy 2>&1 | grep -v "^{" | tail -4
timeout 1200 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -25
