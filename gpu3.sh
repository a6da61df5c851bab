timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for V in "HS2_X_TABS_SMEM=1" "HS2_X_TABS_SMEM=0"; do
echo $V
env $V HS2_Z_PREFETCH=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step',d['ms_per_step'],'value',d['value']/1e9,'G; step frac',d['roofline']['step']['frac'], 'launches', d['gpu_launches'])
for k,v in d['roofline']['kernels'].items(): print(k,v['ms'],v['GBps'],v['frac'])
"
done
