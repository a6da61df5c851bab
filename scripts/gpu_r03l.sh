set -x
for PF in 0 4 8 16; do HS2_PREFETCH=$PF python scripts/ab_sweeps.py --shape 512,512,512 pf$PF= 2>&1 | grep -v "^{"; done
for PF in 0 8; do HS2_PREFETCH=$PF python scripts/ab_sweeps.py --shape 256,256,256 pf$PF= 2>&1 | grep -v "^{"; done
