#!/bin/bash
# N-GPU step benchmark per data-parallel transport (peer-memory kernel with / without multimem, NCCL).
#   gpurun --gpus 2 --timeout 900 -- 'N=2 bash scripts/r2_dp_variants.sh'
N=${N:-8}
EXTRA=${EXTRA:---no-extra}
mkdir -p gpurun_out
run() {
  tag=$(echo "${1:-default}" | tr ' =:,;' '_____')
  env $1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 20 --warmup 3 $EXTRA > gpurun_out/r2m_bench_${N}gpu_$tag.json 2> gpurun_out/r2m_bench_${N}gpu_$tag.err
  echo "[$1] rc=$?"
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2m_bench_${N}gpu_$tag.json").read().splitlines() if l.startswith("{")][-1])
    print("  ms", round(d["ms_per_step"], 3), "img/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "loss", d["final_total_loss"])
except Exception as e:
    print("  no bench line:", e)
PY
  grep -i -E "warn|error|timed out" gpurun_out/r2m_bench_${N}gpu_$tag.err | grep -v -i "futurewarning\|UserWarning: Warning: Profiler" | sort | uniq -c | head -5
}
for v in ${VARIANTS:-"VG_DP_TRANSPORT=peer" "VG_DP_TRANSPORT=peer-nomc" "VG_DP_TRANSPORT=nccl"}; do run "$v"; done
