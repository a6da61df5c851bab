mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sweep_x_kernel|strided_sweep" -s 9 -c 3 -o gpurun_out/prof_r01 -f python profiles/run_steps.py 512 4 > gpurun_out/prof.log 2>&1
tail -1 gpurun_out/prof.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err
tail -c 600 gpurun_out/bench_r01.json
