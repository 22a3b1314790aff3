#!/bin/bash
# one GPU pass: `bash tools/experiments/ab.sh <script.py> <out.txt> <variant> [<variant> ...]` -- "base" is the shipped librto.so
script=$1; out=gpurun_out/$2; shift 2
mkdir -p gpurun_out; : > $out
for v in "$@"; do
  [ "$v" = base ] && v=""
  RTO_LIB_VARIANT=$v timeout 240 python $script $AB_ARGS >> $out 2>&1 || echo "variant '$v' failed rc=$?" >> $out
done
cat $out
