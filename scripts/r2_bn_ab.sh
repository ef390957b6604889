#!/bin/bash
# BatchNorm apply passes: packed in-flight loads + 3 blocks / SM, one resident wave (VG_BN_WAVE=1, default) vs the old grid cap
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_step_gpu.py tests/test_modules_gpu.py -q -m gpu -x -k "bn or batchnorm or fused_step or graph_replay or modules" 2>&1 | tail -3
for v in 1 0 1 0; do
  VG_BN_WAVE=$v timeout 200 python bench.py --steps 40 --warmup 3 --no-extra --no-micro --no-cpu-baseline > gpurun_out/r2v_bench_$v.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/r2v_bench_$v.json'));print('VG_BN_WAVE=$v',d['ms_per_step'],d['value'],d['e2e']['value'])"
done
