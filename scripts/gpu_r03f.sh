# call f: full GPU suite + bench lines after the strided-kernel changes (AHEAD, lean T_in addressing, no L2 warm-up)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for S in 512,512,512 1024,256,512; do
  HS2_B200_LIB=$PWD/heatsim2_b200/libhs2b200_base.so python scripts/ab_sweeps.py --shape $S base= 2>&1 | grep -v "^{"
  python scripts/ab_sweeps.py --shape $S new= 2>&1 | grep -v "^{"
done
python scripts/ab_sweeps.py --problem steelonwater new= 2>&1 | grep -v "^{"
python scripts/slab_bench.py 8 3 10 2>&1 | tail -1
python bench.py > gpurun_out/bench_r03f.json 2> gpurun_out/bench_r03f.err
python scripts/bench_line.py "bench" < gpurun_out/bench_r03f.json || tail -5 gpurun_out/bench_r03f.err
for W in c2_256 c3_steelonwater_512 c4_composite_256x512x512; do
  timeout 900 python bench.py --workload $W --no-cpu-baseline > gpurun_out/bench_${W}_r03f.json 2> gpurun_out/bench_${W}_r03f.err
  python scripts/bench_line.py "$W" < gpurun_out/bench_${W}_r03f.json || tail -5 gpurun_out/bench_${W}_r03f.err
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
