#!/bin/bash
# round-2 evaluation 2: TMA epilogue / 8 epilogue warps - correctness, per-launch timings per switch, tests, bench
mkdir -p gpurun_out
echo "=== harness correctness (default switches)"
timeout 300 ./build/igemm_harness perf > gpurun_out/r2b_harness.log 2>&1; echo "rc=$?"; grep -E "FAIL|failed|ALL PASS|SOME|perf" gpurun_out/r2b_harness.log | tail -24
for v in "" "VG_TEP=0" "VG_TEP=0 VG_WIDE_FROM=256" "VG_EW8=0" "VG_XTMA_WIDE=1" "VG_WIDE_FROM=256" "VG_HALO=64"; do
  tag=$(echo "${v:-default}" | tr ' =' '__')
  echo "=== fused perf [$v]"
  env $v timeout 300 ./build/igemm_harness fused > gpurun_out/r2b_fused_$tag.log 2>&1; echo "rc=$?"
  grep -E "fused " gpurun_out/r2b_fused_$tag.log | cut -c1-100
done
echo "=== pytest kernels + modules"
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_modules_gpu.py -x -q -m gpu 2>&1 | tail -5
echo "=== bench"
timeout 300 python bench.py --steps 30 --warmup 3 --no-micro --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2b_bench.json").read().strip().splitlines()[-1])
    print("  ms_per_step", round(d["ms_per_step"], 3), "images/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "fprop frac", round(d["roofline"]["frac"],3))
except Exception as e:
    print("  no bench line:", e)
PY
tail -5 gpurun_out/r2b_bench.err
