set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for K in f o; do
  HS2_X_KERNEL=$K timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_$K.err | python scripts/bench_line.py "x-kernel=$K"
done
timeout 300 python scripts/slab_bench.py 8 3 10 2>&1 | tail -1
timeout 300 python scripts/slab_bench.py 2 0 10 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sweep_xf|strided_sweep_tma" -s 6 -c 3 -o gpurun_out/prof_c -f python profiles/run_steps.py 512 3 > gpurun_out/prof_c.log 2>&1
tail -1 gpurun_out/prof_c.log
