#!/bin/bash
# Round-2 single-GPU pass: parity tests, bench line (+ reference arm), per-config throughput, builders, ncu launch list + full captures.
# Results land in gpurun_out/ (scratch); tools/summarise_profiles.py --round r02 copies the summaries into profiles/.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
python tools/bench_configs.py --only ${CONFIGS:-C1,C2,C3,C5,C4DC} --out gpurun_out/configs.json > gpurun_out/configs.log 2>&1; echo "configs rc=$?"
[ -n "$SKIP_BUILDERS" ] || { python tools/bench_builders.py --only ${BUILDERS:-c1,dt,c3,c4} --out gpurun_out/builders.json > gpurun_out/builders.log 2>&1; echo "builders rc=$?"; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_render_bvh -s 14 -c 1 -f -o gpurun_out/prof_bvh python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
if [ -n "$CAPTURE_ALL$CAPTURE_OCTREE" ]; then
for c in c3-octA c3-octB dt-octA dt-octB; do
  ncu --set full --clock-control none --import-source on -k regex:k_render_octree -s 2 -c 1 -f -o gpurun_out/prof_$c python tools/profile_case.py $c --reps 4 > gpurun_out/ncu_$c.log 2>&1
done
fi
if [ -n "$CAPTURE_ALL" ]; then
ncu --set full --clock-control none --import-source on -k regex:k_resolve_bvh -s 2 -c 1 -f -o gpurun_out/prof_resolve python tools/profile_case.py dt-bvh-resolve --frames 16 --reps 4 > gpurun_out/ncu_resolve.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_render_bvh -s 2 -c 1 -f -o gpurun_out/prof_c4dc python tools/profile_case.py c4-dcbvh --frames 2 --reps 4 > gpurun_out/ncu_c4dc.log 2>&1
fi
# fifth session: ray lists on the library's own sort, one frame per host call (row bands, page-locked against pageable planes)
if [ -n "$CAPTURE_ALL$EXTRAS" ]; then
python tools/bench_ray_lists.py --out gpurun_out/ray_lists.json > gpurun_out/ray_lists.log 2>&1
for b in 1 4; do RTO_HOST_BANDS=$b python tools/experiments/e2e_single_frame.py; done > gpurun_out/e2e_single_frame.txt 2>&1
E2E_PAGEABLE=1 python tools/experiments/e2e_single_frame.py >> gpurun_out/e2e_single_frame.txt 2>&1
fi
