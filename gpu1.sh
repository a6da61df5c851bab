set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
python heatsim2_b200/build.py 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -30
