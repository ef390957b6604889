#!/bin/bash
# one GPU: Adam issued per bucket from inside the backward pass (VG_LOCAL_ADAM) - which networks pay off
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_step_gpu.py tests/test_next_rows_gpu.py -q -m gpu -x 2>&1 | tail -3
for v in "none" "D" "D,E" "D,E,G"; do
  VG_LOCAL_ADAM=$v timeout 200 python bench.py --steps 40 --warmup 3 --no-extra --no-micro --no-cpu-baseline > gpurun_out/r2r_bench_$v.json 2> gpurun_out/r2r_bench_$v.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2r_bench_$v.json"))
    print("VG_LOCAL_ADAM=$v", round(d["ms_per_step"], 4), round(d["value"]), round(d["e2e"]["value"]), d["gpu_launches_per_step"], d["final_total_loss"])
except Exception as e:
    print("VG_LOCAL_ADAM=$v failed", e)
PY
done
