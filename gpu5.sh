python heatsim2_b200/build.py --force -DHS2_PHASE_TIMING 2>&1 | grep -E "error"
HS2_Z_PREFETCH=1 python profiles/phase_timing_strided.py
HS2_Z_PREFETCH=2 python profiles/phase_timing_strided.py | tail -9
