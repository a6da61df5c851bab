mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"strided_sweep|sweep_x" -s 3 -c 3 -o gpurun_out/prof_all_r01c -f python profiles/run_steps.py 512 2 > gpurun_out/prof.log 2>&1
tail -2 gpurun_out/prof.log
