python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for C in 32 16; do
HS2_CHUNK=$C python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step',d['ms_per_step'],'value',d['value']/1e9,'G; step frac',d['roofline']['step']['frac'], 'launches', d['gpu_launches'])
for k,v in d['roofline']['kernels'].items(): print(k,v['ms'],v['GBps'],v['frac'])
"
done
