#!/bin/bash
# One GPU-box pass: parity tests, bench line (+ reference arm), per-config throughput, ncu launch list + full captures.
# Results land in gpurun_out/ (scratch); summaries are copied into profiles/ by tools/summarise_profiles.py.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
python tools/bench_configs.py --only ${CONFIGS:-C1,C2,C3,C5} --out gpurun_out/configs.json > gpurun_out/configs.log 2>&1; echo "configs rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_render_bvh -s 3 -c 1 -f -o gpurun_out/prof_bvh python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
for c in dt-octA dt-octB c3-octA c3-octB; do
  ncu --set full --clock-control none --import-source on -k regex:k_render_octree -s 2 -c 1 -f -o gpurun_out/prof_$c python tools/profile_case.py $c --reps 4 > gpurun_out/ncu_$c.log 2>&1
done
