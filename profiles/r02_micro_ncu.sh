#!/bin/bash
# ncu --set full of three configurations of profiles/micro/random_records: no probes (0), probes over 592 MB (6), probes over 74 MB (17)
mkdir -p gpurun_out
for c in 0 6 17; do
  ncu --set full --clock-control none -k regex:records_kernel -s 1 -c 1 -f -o gpurun_out/micro_records_c$c profiles/micro/random_records 1000000 $c > gpurun_out/micro_ncu_$c.log 2>&1
  tail -1 gpurun_out/micro_ncu_$c.log
done
ls -la gpurun_out/micro_records_c*.ncu-rep
