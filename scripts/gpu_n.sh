# usage: bash scripts/gpu_n.sh N   - dist tests + N-GPU bench (peer memory and NCCL transports)
N=${1:-2}
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -6
for P2P in 1 0; do
HS2_DIST_P2P=$P2P timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 2>gpurun_out/dist${N}_$P2P.err | tail -1 > gpurun_out/dist${N}_$P2P.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/dist${N}_$P2P.json"))
    print("N=$N P2P=$P2P ms/step",d["ms_per_step"],"G/s",d["value"]/1e9,"e2e ms",d["e2e"]["ms_per_step"],d["comm"]["interface_exchange"],d["comm"].get("rank0_phase_ms"))
except Exception as e:
    print("N=$N P2P=$P2P failed",e); print(open("gpurun_out/dist${N}_$P2P.err").read()[-1500:])
PY
done
