# round-2 session-2 first call: GPU tests, x->y L2 hand-off experiment, bench workloads
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python scripts/xy_pipeline_bench.py 512 20 4 8 16 32 > gpurun_out/xy_pipeline_r02c.log 2>&1
tail -12 gpurun_out/xy_pipeline_r02c.log
python bench.py > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err
python scripts/bench_line.py "bench" < gpurun_out/bench_r02c.json
for W in c2_256 c3_steelonwater_512 c4_composite_256x512x512; do
  timeout 900 python bench.py --workload $W --no-cpu-baseline > gpurun_out/bench_${W}_r02c.json 2> gpurun_out/bench_${W}_r02c.err
  python scripts/bench_line.py "$W" < gpurun_out/bench_${W}_r02c.json || tail -5 gpurun_out/bench_${W}_r02c.err
done
