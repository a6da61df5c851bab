# call e: next-tile line id through a shared-memory slot vs no look-ahead at all (y sweep register pressure)
set -x
for S in 512,512,512 1024,256,512 256,256,256; do
  HS2_B200_LIB=$PWD/heatsim2_b200/libhs2b200_base.so python scripts/ab_sweeps.py --shape $S base= 2>&1 | grep -v "^{"
  python scripts/ab_sweeps.py --shape $S meta= 2>&1 | grep -v "^{"
  HS2_B200_LIB=$PWD/heatsim2_b200/libhs2b200_nometa.so python scripts/ab_sweeps.py --shape $S nometa= 2>&1 | grep -v "^{"
done
for L in base nometa; do HS2_B200_LIB=$PWD/heatsim2_b200/libhs2b200_$L.so python scripts/ab_sweeps.py --problem steelonwater $L= 2>&1 | grep -v "^{"; done
python scripts/ab_sweeps.py --problem steelonwater meta= 2>&1 | grep -v "^{"
for L in base nometa; do HS2_B200_LIB=$PWD/heatsim2_b200/libhs2b200_$L.so python scripts/slab_bench.py 8 3 10 2>&1 | tail -1; done
python scripts/slab_bench.py 8 3 10 2>&1 | tail -1
