#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
timeout 900 python -m pytest -q -p no:cacheprovider tests/test_dprnn_gpu.py tests/test_lstm_gpu.py -m gpu > $O/r02af_dprnn.log 2>&1
echo "dprnn pytest rc=$?"; tail -5 $O/r02af_dprnn.log
timeout 600 python bench.py --config cfg4 --steps 10 --warmup 3 > $O/r02af_cfg4.json 2> $O/r02af_cfg4.err
echo "bench rc=$?"; tail -3 $O/r02af_cfg4.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02af_cfg4.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['e2e']['ms_per_step'])
print(json.dumps(d['roofline'])[:900])
print(d.get('cpu_baseline'))
PY
