mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"sweep_x" -s 1 -c 1 -o gpurun_out/prof_x_r01d -f python profiles/run_steps.py 512 2 > gpurun_out/prof.log 2>&1
tail -2 gpurun_out/prof.log
