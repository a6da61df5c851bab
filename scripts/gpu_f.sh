set -x
mkdir -p gpurun_out
for L in libhs2b200_prev.so libhs2b200.so libhs2b200_prev.so libhs2b200.so; do
  HS2_B200_LIB=$PWD/heatsim2_b200/$L timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_f.err | python scripts/bench_line.py "$L"
done
for L in libhs2b200_prev.so libhs2b200.so; do
HS2_B200_LIB=$PWD/heatsim2_b200/$L timeout 300 python scripts/slab_bench.py 8 3 10 2>&1 | tail -1
done
timeout 600 python -m pytest tests/test_gpu_more.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
