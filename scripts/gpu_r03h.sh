set -x
for S in 512,512,512 256,256,256; do
  python scripts/ab_sweeps.py --shape $S batches= 2>&1 | grep -v "^{"
  for R in 12 16; do HS2_B200_LIB=$PWD/heatsim2_b200/libhs2b200_roll$R.so python scripts/ab_sweeps.py --shape $S roll$R= 2>&1 | grep -v "^{"; done
done
