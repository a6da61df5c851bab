# call b: metadata one tile ahead + sliced cp.async in the persistent strided kernels
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_more.py -x -q -m gpu -k "native_step_loop or device_resident or tile_kernels or constant_bank" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_sizes.py -x -q -m gpu 2>&1 | tail -3
for S in 512,512,512 256,256,256; do
  HS2_B200_LIB=$PWD/heatsim2_b200/libhs2b200_base.so python scripts/ab_sweeps.py --shape $S base= 2>&1 | grep -v "^{"
  HS2_Z_SLICES=1 python scripts/ab_sweeps.py --shape $S slices1= 2>&1 | grep -v "^{"
  HS2_Z_SLICES=2 python scripts/ab_sweeps.py --shape $S slices2= 2>&1 | grep -v "^{"
  HS2_Z_SLICES=4 python scripts/ab_sweeps.py --shape $S slices4= 2>&1 | grep -v "^{"
  HS2_Z_PREFETCH=2 python scripts/ab_sweeps.py --shape $S tma= 2>&1 | grep -v "^{"
done
HS2_B200_LIB=$PWD/heatsim2_b200/libhs2b200_phase.so python profiles/phase_timing_strided.py 2>&1 | tail -24
