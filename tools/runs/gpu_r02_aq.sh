#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
timeout 900 python -m pytest -q -p no:cacheprovider tests/test_parity_gpu.py -m gpu -x -k "fused_mask or model_golden or cfg2 or softmax" > $O/r02aq.log 2>&1
echo "pytest rc=$?"; tail -4 $O/r02aq.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-block > $O/r02aq_bench.json 2> /dev/null
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02aq_bench.json').read().strip().splitlines()[-1])
st = d['stages']
print(f"step {d['ms_per_step']:.3f} e2e {d['e2e']['ms_per_step']:.3f} pw1 {st['pw1']['ms_per_step']:.3f} pw2 {st['pw2']['ms_per_step']:.3f} mask {st['mask']['ms_per_step']:.3f}")
PY
