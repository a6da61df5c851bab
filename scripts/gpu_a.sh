set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for K in f o; do
  HS2_X_KERNEL=$K timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_$K.err | python scripts/bench_line.py "x-kernel=$K"
done
for SH in 128,1024,1024 256,1024,512; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --shape $SH 2>>gpurun_out/bench_shapes.err | python scripts/bench_line.py "shape=$SH"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sweep_xf" -s 2 -c 1 -o gpurun_out/prof_xf -f python profiles/run_steps.py 512 3 > gpurun_out/prof_xf.log 2>&1
tail -2 gpurun_out/prof_xf.log
