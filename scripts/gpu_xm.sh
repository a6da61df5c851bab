# One-GPU measurement set for the z-marching x kernel (kernels_xm.cu): its own tests, A/B of its variants against
# kernels_xf.cu, then bench line, ncu and the whole GPU suite with the best variant.  Outputs under gpurun_out/.
TAG=${1:-r01c}
set -x
mkdir -p gpurun_out
rm -f gpurun_out/xm_best.env
timeout 300 python -m pytest tests/test_gpu_xmarch.py -q > gpurun_out/xm_tests_$TAG.log 2>&1
echo "xm tests rc=$?"
tail -25 gpurun_out/xm_tests_$TAG.log
timeout 300 python scripts/xm_ab.py 512 30 > gpurun_out/xm_ab_$TAG.log 2>&1
echo "ab rc=$?"
tail -20 gpurun_out/xm_ab_$TAG.log
test -f gpurun_out/xm_best.env || exit 0
. gpurun_out/xm_best.env
cat gpurun_out/xm_best.env
timeout 240 python bench.py --no-cpu-baseline > gpurun_out/bench_march_$TAG.json 2> gpurun_out/bench_march_$TAG.err
python scripts/bench_line.py "bench march" < gpurun_out/bench_march_$TAG.json
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/suite_march_$TAG.log 2>&1
echo "suite(march) rc=$?"
tail -5 gpurun_out/suite_march_$TAG.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"sweep_xm" -s 2 -c 1 -o gpurun_out/prof_xm_$TAG -f python profiles/run_steps.py 512 4 > gpurun_out/prof_xm_$TAG.log 2>&1
tail -2 gpurun_out/prof_xm_$TAG.log
