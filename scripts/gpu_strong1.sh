set -x
mkdir -p gpurun_out
timeout 900 python bench.py --scaling strong --steps 20 --no-cpu-baseline > gpurun_out/bench_strong1_r02.json 2> gpurun_out/bench_strong1_r02.err
python scripts/bench_line.py "strong N=1 1024^3" < gpurun_out/bench_strong1_r02.json || tail -5 gpurun_out/bench_strong1_r02.err
