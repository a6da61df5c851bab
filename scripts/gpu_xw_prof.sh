set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"sweep_xw" -s 2 -c 1 -o gpurun_out/prof_xw_r02c -f python profiles/run_steps.py 512 4 > gpurun_out/prof_xw_r02c.log 2>&1
tail -2 gpurun_out/prof_xw_r02c.log
