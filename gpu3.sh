timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_more.py -x -q 2>&1 | tail -2
for SH in "128,1024,1024" "1024,1024,128"; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --shape $SH 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$SH ms/step',d['ms_per_step'],'x',d['roofline']['kernels']['x']['ms'],'y',d['roofline']['kernels']['y']['ms'],'z',d['roofline']['kernels']['z']['ms'])
"
done
