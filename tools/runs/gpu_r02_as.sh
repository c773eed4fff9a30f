#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
timeout 900 python -m pytest -q -p no:cacheprovider tests/test_parity_gpu.py -m gpu -x -k "model_golden or cfg2 or envelope" > $O/r02as.log 2>&1
echo "pytest rc=$?"; tail -3 $O/r02as.log
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-block > $O/r02as_bench.json 2> /dev/null
python - <<'PY'
import json, sys
d = json.loads(open('gpurun_out/r02as_bench.json').read().strip().splitlines()[-1])
st = d['stages']
print(f"step {d['ms_per_step']:.3f} e2e {d['e2e']['ms_per_step']:.3f} pw1 {st['pw1']['ms_per_step']:.3f} pw2 {st['pw2']['ms_per_step']:.3f} mask {st['mask']['ms_per_step']:.3f}")
PY
done
