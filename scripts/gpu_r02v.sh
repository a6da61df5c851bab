# wide class ids, pinned result arrays, smoke, full suite, bench
TAG=r02v
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -12
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python scripts/bench_line.py "bench" < gpurun_out/bench_$TAG.json || tail -5 gpurun_out/bench_$TAG.err
