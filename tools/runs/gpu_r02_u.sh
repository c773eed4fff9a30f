#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
CTN_LSTM_DBG=16 timeout 120 python tools/lstm_time.py 2>&1 | grep "step 10[12]"
for d in 1 2 3; do CTN_LSTM_DBG=$d timeout 120 python tools/lstm_time.py 2>&1 | grep dbg | head -1; done
timeout 900 python -m pytest -q -p no:cacheprovider tests/test_dprnn_gpu.py tests/test_lstm_gpu.py -m gpu > $O/r02u_dprnn.log 2>&1
echo "dprnn pytest rc=$?"; tail -5 $O/r02u_dprnn.log
timeout 600 python bench.py --config cfg4 --steps 5 --warmup 3 --no-cpu-baseline > $O/r02u_cfg4.json 2> $O/r02u_cfg4.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02u_cfg4.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['e2e']['ms_per_step'])
PY
