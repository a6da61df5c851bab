# usage: bash scripts/gpu_n2.sh N [strong]  - dist tests + N-GPU bench, fused z sweep vs forward/backward launches
N=${1:-2}
SC=${2:-weak}
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
if [ "${3:-no}" = "tests" ]; then timeout 900 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -6; fi
for FUSED in ${FUSED_LIST:-1 0}; do
HS2_DIST_Z_FUSED=$FUSED timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 3 --scaling $SC $EXTRA 2>gpurun_out/dist${N}_${SC}_f$FUSED.err | tail -1 > gpurun_out/dist${N}_${SC}_f$FUSED.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/dist${N}_${SC}_f$FUSED.json"))
    print("N=$N $SC fused=$FUSED ms/step",d["ms_per_step"],"G/s",d["value"]/1e9,"e2e ms",d["e2e"]["ms_per_step"],"setup",d["config"]["setup_s"],d["check"])
    for r,ph in enumerate(d["comm"].get("phase_ms_by_rank") or []): print(" rank",r,{k:round(v,3) for k,v in (ph or {}).items()})
except Exception as e:
    print("N=$N fused=$FUSED failed",e); print(open("gpurun_out/dist${N}_${SC}_f$FUSED.err").read()[-1500:])
PY
done
