#!/bin/bash
# round-2 evaluation 4: elect.sync-guarded issue loops - correctness, per-launch timings, tests, bench
mkdir -p gpurun_out
T=${TAG:-r2d}
echo "=== harness correctness"
timeout 300 ./build/igemm_harness > gpurun_out/${T}_harness.log 2>&1; echo "rc=$?"; grep -E "FAIL|failed|ALL PASS|SOME" gpurun_out/${T}_harness.log | tail -5
for v in "" ${VARIANTS}; do
  tag=$(echo "${v:-default}" | tr ' =' '__')
  echo "=== fused perf [$v]"
  env $v timeout 300 ./build/igemm_harness fused > gpurun_out/${T}_fused_$tag.log 2>&1; echo "rc=$?"
  grep -E "fused |wgrad " gpurun_out/${T}_fused_$tag.log | cut -c1-100
done
echo "=== pytest"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for v in "" ${VARIANTS}; do
  tag=$(echo "${v:-default}" | tr ' =' '__')
  echo "=== bench [$v]"
  env $v timeout 300 python bench.py --steps 30 --warmup 3 --no-micro --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_$tag.json 2> gpurun_out/${T}_bench_$tag.err; echo "rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${T}_bench_$tag.json").read().strip().splitlines()[-1])
    print("  ms_per_step", round(d["ms_per_step"], 3), "images/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "fprop frac", round(d["roofline"]["frac"],3), "wgrad us", round(d["roofline"]["wgrad_kernel"]["us_per_launch"],1))
except Exception as e:
    print("  no bench line:", e)
PY
done
