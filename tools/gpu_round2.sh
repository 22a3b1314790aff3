#!/bin/bash
# Round-2 single-GPU pass: parity tests, A/B of kernel variants, per-config throughput, ncu captures of the final kernels incl. C4.
# Results land in gpurun_out/ (scratch); tools/summarise_profiles.py --round r02 copies the summaries into profiles/.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
( for V in "" ld256; do for C in dt-bvh dt-bvh-primary c3-bvh; do echo "== variant [$V] $C"; RTO_LIB_VARIANT=$V python tools/profile_case.py $C --frames 16 --reps 7 2>&1 | tail -1; done; done
  python tools/profile_case.py dt-bvh-resolve --frames 16 --reps 7 2>&1 | tail -2 ) > gpurun_out/variants.log 2>&1
python tools/bench_configs.py --only ${CONFIGS:-C1,C2,C3,C5,C4DC} --out gpurun_out/configs.json > gpurun_out/configs.log 2>&1; echo "configs rc=$?"
for c in c3-octA c3-octB dt-octA dt-octB; do
  ncu --set full --clock-control none --import-source on -k regex:k_render_octree -s 2 -c 1 -f -o gpurun_out/prof_$c python tools/profile_case.py $c --reps 4 > gpurun_out/ncu_$c.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:k_resolve_bvh -s 2 -c 1 -f -o gpurun_out/prof_resolve python tools/profile_case.py dt-bvh-resolve --frames 16 --reps 4 > gpurun_out/ncu_resolve.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_render_bvh -s 2 -c 1 -f -o gpurun_out/prof_c4dc python tools/profile_case.py c4-dcbvh --frames 2 --reps 4 > gpurun_out/ncu_c4dc.log 2>&1
