#!/bin/bash
# 8-GPU (or $N-GPU) step benchmark under NCCL algorithm / protocol choices for the bucketed gradient all-reduce.
#   gpurun --gpus 8 --timeout 900 -- 'bash scripts/r2_nccl_variants.sh'
N=${N:-8}
mkdir -p gpurun_out
run() {
  tag=$(echo "${1:-default}" | tr ' =:,;' '_____')
  env $1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 20 --warmup 3 --no-extra > gpurun_out/r2j_bench_${N}gpu_$tag.json 2> gpurun_out/r2j_bench_${N}gpu_$tag.err
  echo "[$1] rc=$?"
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2j_bench_${N}gpu_$tag.json").read().splitlines() if l.startswith("{")][-1])
    print("  ms", round(d["ms_per_step"], 3), "img/s", round(d["value"]), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("  no bench line:", e)
PY
}
run ""
run "NCCL_ALGO=allreduce:nvls"
run "NCCL_ALGO=allreduce:nvlstree"
run "NCCL_ALGO=allreduce:ring NCCL_PROTO=Simple"
