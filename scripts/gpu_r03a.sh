# round-2 session 4, call a: native step loop tests, z-sweep fetch variants (A/B in one box), phase marks, C4 bench
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_more.py -x -q -m gpu -k "native_step_loop or device_resident" 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_sizes.py -x -q -m gpu -k "survey_sizes" 2>&1 | tail -5
for S in 512,512,512 256,256,256; do
  HS2_B200_LIB=$PWD/heatsim2_b200/libhs2b200_base.so python scripts/ab_sweeps.py --shape $S base= 2>&1 | grep -v "^{"
  python scripts/ab_sweeps.py --shape $S cpasync= 2>&1 | grep -v "^{"
  HS2_Z_PREFETCH=3 python scripts/ab_sweeps.py --shape $S rows= 2>&1 | grep -v "^{"
  HS2_Z_PREFETCH=2 python scripts/ab_sweeps.py --shape $S tma= 2>&1 | grep -v "^{"
done
HS2_B200_LIB=$PWD/heatsim2_b200/libhs2b200_phase.so python profiles/phase_timing_strided.py 2>&1 | tail -24
HS2_Z_PREFETCH=3 HS2_B200_LIB=$PWD/heatsim2_b200/libhs2b200_phase.so python profiles/phase_timing_strided.py 2>&1 | tail -12
python bench.py --workload c4_composite_256x512x512 --no-cpu-baseline --steps 30 > gpurun_out/bench_c4_r03a.json 2> gpurun_out/bench_c4_r03a.err
python scripts/bench_line.py c4 < gpurun_out/bench_c4_r03a.json || tail -5 gpurun_out/bench_c4_r03a.err
