# call d: z sweep without the L2 warm-up of T_in (default), lean addressing in the slab z kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_more.py tests/test_gpu_sizes.py -x -q -m gpu -k "tile_kernels or constant_bank or long_lines or survey or slab" 2>&1 | tail -3
for S in 512,512,512 256,256,256 1024,256,512; do
  HS2_B200_LIB=$PWD/heatsim2_b200/libhs2b200_base.so python scripts/ab_sweeps.py --shape $S base= 2>&1 | grep -v "^{"
  HS2_Z_PACE=0 python scripts/ab_sweeps.py --shape $S burst= 2>&1 | grep -v "^{"
  HS2_Z_PACE=1 python scripts/ab_sweeps.py --shape $S paced= 2>&1 | grep -v "^{"
done
HS2_B200_LIB=$PWD/heatsim2_b200/libhs2b200_base.so python scripts/ab_sweeps.py --problem steelonwater base= 2>&1 | grep -v "^{"
python scripts/ab_sweeps.py --problem steelonwater new= 2>&1 | grep -v "^{"
HS2_B200_LIB=$PWD/heatsim2_b200/libhs2b200_phase.so python profiles/phase_timing_strided.py 2>&1 | tail -12
HS2_B200_LIB=$PWD/heatsim2_b200/libhs2b200_base.so python scripts/slab_bench.py 2 0 10 2>&1 | tail -8
python scripts/slab_bench.py 2 0 10 2>&1 | tail -8
HS2_B200_LIB=$PWD/heatsim2_b200/libhs2b200_base.so python scripts/slab_bench.py 8 3 10 2>&1 | tail -8
python scripts/slab_bench.py 8 3 10 2>&1 | tail -8
