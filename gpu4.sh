mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"strided|rhs_kernel|thomas" -c 12 --csv --log-file gpurun_out/launches_r01a.csv python profiles/run_steps.py 512 3 > gpurun_out/launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"strided" -s 2 -c 2 -o gpurun_out/prof_strided_r01a -f python profiles/run_steps.py 512 2 > gpurun_out/prof.log 2>&1
tail -3 gpurun_out/prof.log
cat gpurun_out/launches_r01a.csv | tail -14
