python heatsim2_b200/build.py --force -DHS2_PHASE_TIMING 2>&1 | grep -E "error"
python profiles/phase_timing.py 512
