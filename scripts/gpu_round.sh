# End-of-round measurement set on one B200: GPU tests, bench (both arms), ncu launch list,
# ncu --set full of the three sweep kernels, smoke.  Outputs under gpurun_out/.
TAG=${1:-r01b}
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python scripts/bench_line.py "bench" < gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
tail -c 400 gpurun_out/bench_ref_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sweep|thomas_kernel|rhs_kernel|z_forward|z_backward" -c 60 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sweep_xf|strided_sweep" -s 9 -c 3 -o gpurun_out/prof_$TAG -f python profiles/run_steps.py 512 4 > gpurun_out/prof_$TAG.log 2>&1
tail -1 gpurun_out/prof_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()"
