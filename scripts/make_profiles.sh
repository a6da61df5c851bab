# after scripts/gpu_final.sh TAG: turn gpurun_out/ into the tracked summaries under profiles/
TAG=${1:-r02z}
set -e
cp gpurun_out/launches_$TAG.csv profiles/launches_r02.csv
python profiles/summarize.py launches gpurun_out/launches_$TAG.csv > profiles/launches_r02.md
python profiles/summarize.py full gpurun_out/prof_256_$TAG.ncu-rep 16777216 notraffic > profiles/kernels_256_r02.md
python profiles/summarize.py full gpurun_out/prof_$TAG.ncu-rep 134217728 > profiles/kernels_r02.md
python profiles/src_hotspots.py gpurun_out/prof_$TAG.ncu-rep 1.5 sweep_xw2 > profiles/hotspots_x_r02.txt 2>&1 || true
for f in bench_$TAG bench_c2_256_$TAG bench_c3_steelonwater_512_$TAG bench_c4_composite_256x512x512_$TAG bench_ref_$TAG; do
  grep "^{" gpurun_out/$f.json | tail -1 > profiles/${f/$TAG/r02}.json
done
ls -la profiles/*r02*
