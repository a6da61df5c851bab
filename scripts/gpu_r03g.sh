set -x
python scripts/resident_probe.py 512 2>&1 | tail -8
timeout 600 python scripts/xy_pipeline_bench.py 512 20 8 16 32 64 2>&1 | tail -8
