# xw2 default: full GPU suite, bench, ncu of the three kernels
TAG=r02j
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python scripts/bench_line.py "bench" < gpurun_out/bench_$TAG.json || tail -5 gpurun_out/bench_$TAG.err
ncu --set full --clock-control none --import-source on -k regex:"sweep_xw|strided_sweep" -s 9 -c 3 -o gpurun_out/prof_$TAG -f python profiles/run_steps.py 512 4 > gpurun_out/prof_$TAG.log 2>&1
tail -1 gpurun_out/prof_$TAG.log
