#!/bin/bash
# ncu --set full of single fused-harness launches (one per case / switch setting); reports land in gpurun_out/
mkdir -p gpurun_out
run() {  # tag, case index, env...
  tag=$1; idx=$2; shift 2
  env "$@" ./build/igemm_harness fused $idx > gpurun_out/plain_$tag.log 2>&1 && \
  env "$@" ncu --set full --clock-control none --import-source on -k regex:igemm_fprop -s 3 -c 1 -f -o gpurun_out/r2_prof_$tag ./build/igemm_harness fused $idx > gpurun_out/ncu_$tag.log 2>&1
  echo "$tag rc=$?"; grep fused gpurun_out/plain_$tag.log | cut -c1-100
}
run f14_default 14 VG_NONE=1
run f14_halo 14 VG_HALO=64
run f8_default 8 VG_NONE=1
run f0_default 0 VG_NONE=1
