#!/bin/bash
# round-2 evaluation 3: row-shift-only halo tiles (VG_HALO) - correctness and per-launch timings
mkdir -p gpurun_out
echo "=== harness correctness VG_HALO=1"
VG_HALO=1 timeout 300 ./build/igemm_harness > gpurun_out/r2c_harness_halo.log 2>&1; echo "rc=$?"; grep -E "FAIL|failed|ALL PASS|SOME" gpurun_out/r2c_harness_halo.log | tail -5
for v in "" "VG_HALO=64" "VG_HALO=128" "VG_HALO=1"; do
  tag=$(echo "${v:-default}" | tr ' =' '__')
  echo "=== fused perf [$v]"
  env $v timeout 300 ./build/igemm_harness fused > gpurun_out/r2c_fused_$tag.log 2>&1; echo "rc=$?"
  grep -E "fused " gpurun_out/r2c_fused_$tag.log | cut -c1-100
done
echo "=== pytest kernels + modules with VG_HALO=1"
VG_HALO=1 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_modules_gpu.py -x -q -m gpu 2>&1 | tail -5
for v in "" "VG_HALO=64" "VG_HALO=128"; do
  tag=$(echo "${v:-default}" | tr ' =' '__')
  echo "=== bench [$v]"
  env $v timeout 300 python bench.py --steps 30 --warmup 3 --no-micro --no-cpu-baseline > gpurun_out/r2c_bench_$tag.json 2> gpurun_out/r2c_bench_$tag.err; echo "rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2c_bench_$tag.json").read().strip().splitlines()[-1])
    print("  ms_per_step", round(d["ms_per_step"], 3), "images/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "fprop frac", round(d["roofline"]["frac"],3), "wgrad us", round(d["roofline"]["wgrad_kernel"]["us_per_launch"],1))
except Exception as e:
    print("  no bench line:", e)
PY
done
