set -x
python bench.py --steps 10 --warmup 3 --grid 256 --no-cpu-baseline 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 2>&1 | tail -3
python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -2
