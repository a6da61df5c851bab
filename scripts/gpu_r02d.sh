# round-2 session-3 first call: native plan builder + warp x kernel everywhere: tests, bench lines, ncu
TAG=r02d
set -x
mkdir -p gpurun_out
date
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -14
date
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python scripts/bench_line.py "bench" < gpurun_out/bench_$TAG.json || tail -5 gpurun_out/bench_$TAG.err
date
for W in c2_256 c3_steelonwater_512 c4_composite_256x512x512; do
  timeout 900 python bench.py --workload $W --no-cpu-baseline > gpurun_out/bench_${W}_$TAG.json 2> gpurun_out/bench_${W}_$TAG.err
  python scripts/bench_line.py "$W" < gpurun_out/bench_${W}_$TAG.json || tail -5 gpurun_out/bench_${W}_$TAG.err
done
date
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sweep|thomas_kernel|rhs_kernel|z_forward|z_backward" -c 60 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
date
ncu --set full --clock-control none --import-source on -k regex:"sweep_xw|strided_sweep" -s 9 -c 3 -o gpurun_out/prof_$TAG -f python profiles/run_steps.py 512 4 > gpurun_out/prof_$TAG.log 2>&1
tail -1 gpurun_out/prof_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:"sweep_xw|strided_sweep" -s 9 -c 3 -o gpurun_out/prof_256_$TAG -f python profiles/run_steps.py 256 4 > gpurun_out/prof_256_$TAG.log 2>&1
tail -1 gpurun_out/prof_256_$TAG.log
date
