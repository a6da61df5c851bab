set -x
mkdir -p gpurun_out
for K in f o; do for PF in 0 1; do
  HS2_X_KERNEL=$K HS2_X_PREFETCH=$PF timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_$K$PF.err | python scripts/bench_line.py "x-kernel=$K prefetch=$PF"
done; done
timeout 300 python scripts/slab_bench.py 8 3 10 2>&1 | tail -2
timeout 300 python scripts/slab_bench.py 4 1 10 2>&1 | tail -2
timeout 300 python scripts/slab_bench.py 2 0 10 2>&1 | tail -2
