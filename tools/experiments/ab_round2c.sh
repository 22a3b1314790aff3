#!/bin/bash
# third session of round 2: resolve with four pixels per thread, octree mode A with the climbs folded into the step.  -> gpurun_out/ab_round2c.txt
mkdir -p gpurun_out
out=gpurun_out/ab_round2c.txt
: > $out
python -m pytest tests/test_gpu_codes.py -q -x -m gpu > gpurun_out/ab_pytest_codes.log 2>&1; echo "pytest test_gpu_codes (resolve4 default) rc=$?" >> $out; tail -1 gpurun_out/ab_pytest_codes.log >> $out
RTO_LIB_VARIANT=fold python -m pytest tests/test_gpu_parity.py -q -x -m gpu > gpurun_out/ab_pytest_fold.log 2>&1; echo "pytest test_gpu_parity (variant fold) rc=$?" >> $out; tail -1 gpurun_out/ab_pytest_fold.log >> $out
for v in "" fold; do
  for c in c3-octA dt-octA; do
    echo "== variant [$v] $c" >> $out
    RTO_LIB_VARIANT=$v python tools/profile_case.py $c --frames 16 2>&1 | tail -1 >> $out
  done
done
echo "== resolve, one pixel per thread (RTO_RESOLVE4=0)" >> $out
RTO_RESOLVE4=0 python tools/profile_case.py dt-bvh-resolve --frames 16 2>&1 | tail -2 >> $out
for v in "" r4b4 r4b8; do
  echo "== resolve, four pixels per thread, variant [$v]" >> $out
  RTO_LIB_VARIANT=$v python tools/profile_case.py dt-bvh-resolve --frames 16 2>&1 | tail -2 >> $out
done
cat $out
