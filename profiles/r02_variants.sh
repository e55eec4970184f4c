#!/bin/bash
# A/B of library builds (build.py --variant) on one box: the c2 bench at N vectors for each named variant.
# usage: profiles/r02_variants.sh N "<variant>[:bench args]" ...   ("base" = the shipped library)
N=${1:-250000}; shift
mkdir -p gpurun_out
B="python bench.py --nvec $N --steps 10 --warmup 3 --no-gate --no-stream --no-recall --no-cpu-baseline --no-c4"
for V in "$@"; do
  name=${V%%:*}; extra=""; [[ "$V" == *:* ]] && extra=${V#*:}
  lib=""; [ "$name" != base ] && lib=$PWD/rabitq-ann-search_b200/cphnsw_b200/variants/libcphnsw_b200_$name.so
  echo "== $name $extra"
  CPHNSW_B200_LIB=$lib $B $extra 2>gpurun_out/err_variant.log | tee -a gpurun_out/r02_variants.jsonl | python profiles/pj.py || tail -5 gpurun_out/err_variant.log
done
