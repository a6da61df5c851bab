# ncu --set full of the slab z kernels (what one rank of a 2-GPU run launches), on one GPU
set -x
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"z_forward|z_backward" -s 4 -c 2 -o gpurun_out/prof_slab_r03o -f python scripts/slab_bench.py 2 0 4 > gpurun_out/prof_slab_r03o.log 2>&1
tail -2 gpurun_out/prof_slab_r03o.log
