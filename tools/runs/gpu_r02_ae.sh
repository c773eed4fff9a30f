#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
for pc in 0 100 116 124 132 140 160; do
  CTN_UMMA_DBG=$((pc * 256)) timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train-block > gpurun_out/r02ae_$pc.json 2> gpurun_out/r02ae_$pc.err
  python - "$pc" <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/r02ae_{sys.argv[1]}.json').read().strip().splitlines()[-1])
st = d['stages']
print(f"pace {sys.argv[1]:>4}: step {d['ms_per_step']:.3f}  pw1 {st['pw1']['ms_per_step']:.3f}  pw2 {st['pw2']['ms_per_step']:.3f}  mask {st['mask']['ms_per_step']:.3f}")
PY
done
