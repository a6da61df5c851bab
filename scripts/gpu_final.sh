# End-of-round measurement set on one B200 (round 2): bench lines of every BASELINE workload, reference arm,
# ncu launch list, ncu --set full of the three sweep kernels at 512^3 and 256^3, smoke.  Outputs under gpurun_out/.
TAG=${1:-r02z}
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python scripts/bench_line.py "bench" < gpurun_out/bench_$TAG.json || tail -5 gpurun_out/bench_$TAG.err
for W in c2_256 c3_steelonwater_512 c4_composite_256x512x512; do
  timeout 900 python bench.py --workload $W --no-cpu-baseline > gpurun_out/bench_${W}_$TAG.json 2> gpurun_out/bench_${W}_$TAG.err
  python scripts/bench_line.py "$W" < gpurun_out/bench_${W}_$TAG.json || tail -5 gpurun_out/bench_${W}_$TAG.err
done
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
tail -c 600 gpurun_out/bench_ref_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sweep|thomas_kernel|rhs_kernel|z_forward|z_backward" -c 60 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sweep_xw|strided_sweep" -s 9 -c 3 -o gpurun_out/prof_$TAG -f python profiles/run_steps.py 512 4 > gpurun_out/prof_$TAG.log 2>&1
tail -1 gpurun_out/prof_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:"sweep_xw|strided_sweep" -s 9 -c 3 -o gpurun_out/prof_256_$TAG -f python profiles/run_steps.py 256 4 > gpurun_out/prof_256_$TAG.log 2>&1
tail -1 gpurun_out/prof_256_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
