set -x
python scripts/ab_sweeps.py --shape 512,512,512 s24=HS2_XW_SHAPE:24 s23=HS2_XW_SHAPE:23 s24b=HS2_XW_SHAPE:24 2>&1 | grep -v "^{"
python scripts/ab_sweeps.py --shape 256,512,512 s24=HS2_XW_SHAPE:24 s23=HS2_XW_SHAPE:23 2>&1 | grep -v "^{"
python scripts/ab_sweeps.py --problem steelonwater s24=HS2_XW_SHAPE:24 s23=HS2_XW_SHAPE:23 2>&1 | grep -v "^{"
