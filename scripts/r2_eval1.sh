#!/bin/bash
# round-2 evaluation 1: correctness of the new paths + per-launch timings with the switches on/off
mkdir -p gpurun_out
echo "=== harness correctness (default switches)"
timeout 300 ./build/igemm_harness > gpurun_out/r2_harness.log 2>&1; echo "rc=$?"; grep -E "FAIL|failed|ALL PASS|SOME" gpurun_out/r2_harness.log | tail -5
echo "=== harness correctness VG_WGRAD_ATOMIC=0"
VG_WGRAD_ATOMIC=0 timeout 300 ./build/igemm_harness > gpurun_out/r2_harness_na.log 2>&1; echo "rc=$?"; grep -E "FAIL|failed|ALL PASS|SOME" gpurun_out/r2_harness_na.log | tail -5
for v in "" "VG_XS=0" "VG_XS=4" "VG_XS=2" "VG_WGRAD_ATOMIC=0" "VG_HALO=64" "VG_HALO=64 VG_XS=0"; do
  tag=$(echo "${v:-default}" | tr ' =' '__')
  echo "=== fused perf [$v]"
  env $v timeout 300 ./build/igemm_harness fused > gpurun_out/r2_fused_$tag.log 2>&1; echo "rc=$?"
  cat gpurun_out/r2_fused_$tag.log | grep -E "fused|wgrad" 
done
