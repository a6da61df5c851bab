# call c: lean addressing in the z sweep's T_in / store phase, cp.async pieces paced by the forward elimination
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_more.py -x -q -m gpu -k "tile_kernels or constant_bank" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_sizes.py -x -q -m gpu -k "long_lines or survey" 2>&1 | tail -3
for S in 512,512,512 256,256,256; do
  HS2_B200_LIB=$PWD/heatsim2_b200/libhs2b200_base.so python scripts/ab_sweeps.py --shape $S base= 2>&1 | grep -v "^{"
  HS2_Z_PACE=0 python scripts/ab_sweeps.py --shape $S burst= 2>&1 | grep -v "^{"
  HS2_Z_PACE=1 python scripts/ab_sweeps.py --shape $S paced= 2>&1 | grep -v "^{"
  HS2_Z_PACE=1 HS2_PREFETCH=0 python scripts/ab_sweeps.py --shape $S paced_nopf= 2>&1 | grep -v "^{"
done
HS2_B200_LIB=$PWD/heatsim2_b200/libhs2b200_base.so python scripts/ab_sweeps.py --problem steelonwater base= 2>&1 | grep -v "^{"
python scripts/ab_sweeps.py --problem steelonwater new= 2>&1 | grep -v "^{"
HS2_B200_LIB=$PWD/heatsim2_b200/libhs2b200_phase.so python profiles/phase_timing_strided.py 2>&1 | tail -12
