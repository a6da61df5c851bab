mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err
tail -c 1500 gpurun_out/bench_r01.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r01.json 2>> gpurun_out/bench_r01.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sweep|thomas_kernel|rhs_kernel|z_forward|z_backward" -c 60 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sweep_x_kernel|strided_sweep" -s 9 -c 3 -o gpurun_out/prof_r01 -f python profiles/run_steps.py 512 4 > gpurun_out/prof.log 2>&1
tail -1 gpurun_out/prof.log
python -c "import __graft_entry__ as g; g.smoke()"
