for V in "HS2_PREFETCH=0" "HS2_PREFETCH=1" "HS2_NO_TMA_Z=1"; do
echo $V
env $V timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step',d['ms_per_step'],'value',d['value']/1e9,'G; step frac',d['roofline']['step']['frac'], 'launches', d['gpu_launches'])
for k,v in d['roofline']['kernels'].items(): print(k,v['ms'],v['GBps'],v['frac'])
"
done
