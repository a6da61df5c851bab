set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_e.err | python scripts/bench_line.py "512^3"
for R in 8 4; do
HS2_X_R=$R timeout 300 python scripts/slab_bench.py 8 3 10 2>&1 | tail -1
done
HS2_X_R=4 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_e4.err | python scripts/bench_line.py "512^3 R=4"
