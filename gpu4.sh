mkdir -p gpurun_out
HS2_CHUNK_X=16 ncu --set full --clock-control none --import-source on -k regex:"sweep_x_march" -s 1 -c 1 -o gpurun_out/prof_xm_r01 -f python profiles/run_steps.py 512 2 > gpurun_out/prof.log 2>&1
tail -2 gpurun_out/prof.log
