#!/bin/bash
# One gpurun call that evaluates the halo-tile variant:
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash scripts/halo_eval.sh'
# 1. native self-check (bit-level correctness of every contraction against a CPU loop) + per-layer timings,
#    default build vs VG_HALO=1 (every eligible launch) vs VG_HALO=64 (N tiles of at most 64 only)
# 2. the GPU parity tests of the kernels / modules with the variant on
# 3. the step benchmark, default vs VG_HALO=64
# Outputs land in gpurun_out/halo_*.log.
mkdir -p gpurun_out
for v in "" 1 64; do
    tag=${v:-off}
    echo "=== harness VG_HALO=$tag"
    if [ -z "$v" ]; then timeout 120 ./build/igemm_harness perf > gpurun_out/halo_harness_$tag.log 2>&1
    else VG_HALO=$v timeout 120 ./build/igemm_harness perf > gpurun_out/halo_harness_$tag.log 2>&1; fi
    echo "rc=$?"; grep -E "FAIL|mismatch|error|PASS$|perf .*(down|up) " gpurun_out/halo_harness_$tag.log | tail -24
done
if grep -q "ALL PASS" gpurun_out/halo_harness_1.log; then
    echo "=== pytest (kernels + modules) with VG_HALO=1"
    VG_HALO=1 timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_modules_gpu.py -x -q -m gpu 2>&1 | tail -3
    for v in "" 64 1; do
        tag=${v:-off}
        echo "=== bench VG_HALO=$tag"
        if [ -z "$v" ]; then timeout 200 python bench.py --steps 30 --warmup 3 --no-micro --no-cpu-baseline > gpurun_out/halo_bench_$tag.json 2> gpurun_out/halo_bench_$tag.err
        else VG_HALO=$v timeout 200 python bench.py --steps 30 --warmup 3 --no-micro --no-cpu-baseline > gpurun_out/halo_bench_$tag.json 2> gpurun_out/halo_bench_$tag.err; fi
        python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/halo_bench_$tag.json").read().strip().splitlines()[-1])
    print("  ms_per_step", round(d["ms_per_step"], 3), "images/s", round(d["value"]))
except Exception as e:
    print("  no bench line:", e)
PY
    done
else
    echo "halo variant did not pass the native self-check - see gpurun_out/halo_harness_1.log"
fi
