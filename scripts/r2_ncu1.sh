#!/bin/bash
mkdir -p gpurun_out
for i in 0 13 7; do
  ./build/igemm_harness fused $i > gpurun_out/plain_$i.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:igemm_fprop -s 3 -c 1 -f -o gpurun_out/r2_prof_fused$i ./build/igemm_harness fused $i > gpurun_out/ncu_$i.log 2>&1
  echo "case $i rc=$?"; cat gpurun_out/plain_$i.log | grep fused
done
