for V in "HS2_Z_PREFETCH=1" "HS2_Z_PREFETCH=2" "HS2_Z_PREFETCH=1 HS2_PREFETCH=0"; do
echo $V
env $V timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step',d['ms_per_step'],'x',d['roofline']['kernels']['x']['ms'],'y',d['roofline']['kernels']['y']['ms'],'z',d['roofline']['kernels']['z']['ms'])
"
done
