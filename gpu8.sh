nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -3
for N in 8 4; do
for PL in 2 1; do
HS2_DIST_PIPELINE=$PL timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=$N pipeline $PL: ms/step',d['ms_per_step'],'value',d['value']/1e9,'G', d['config']['grid'], d['comm']['interface_exchange'], 'e2e', d['e2e']['ms_per_step'])
"
done
done
