#!/bin/bash
# A/B of kernel builds on one B200: tools/gpu_ab.sh <tag> <variant>[:check] ...   (variant "default" = libmcgp.so)
tag=$1; shift
out=gpurun_out/ab_$tag.jsonl
mkdir -p gpurun_out
: > $out
for v in "$@"; do
  name=${v%%:*}; chk=""
  [[ "$v" == *":check" ]] && chk="--check"
  lib=monte-carlo-gp_b200/libmcgp_$name.so
  [[ "$name" == "default" ]] && lib=monte-carlo-gp_b200/libmcgp.so
  MCGP_LIB_PATH=$PWD/$lib timeout 300 python tests/checkers/ab_bench.py --tag $name $chk >> $out 2>> gpurun_out/ab_$tag.err || echo "{\"tag\": \"$name\", \"failed\": true}" >> $out
done
cat $out
