#!/bin/bash
# debug helper: parity + bench of the CTA-pair mode under several knobs (run on the GPU box)
mkdir -p gpurun_out
for mn in 32; do
  echo "== PAIR_MIN_N=$mn"
  CTN_UMMA_PAIR_MIN_N=$mn timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "tdcn_golden or model_golden or tf32" 2>&1 | grep -E "passed|failed|FAILED|Mismatch|Greatest" | head -20
done
run() {  # cluster dbg
  CTN_UMMA_CLUSTER=$1 CTN_UMMA_DBG=$2 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/pair_bench_c$1_d$2.json 2> gpurun_out/pair_bench_c$1_d$2.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/pair_bench_c$1_d$2.json").read().strip().splitlines()[-1])
    print("cluster $1 dbg $2 ms/step", round(d["ms_per_step"],3), {k: round(v["ms_per_step"],3) for k,v in d["stages"].items()})
except Exception as e:
    print("cluster $1 dbg $2 FAILED", e)
PY
}
run 2 0; run 2 70; run 2 8; run 2 6
