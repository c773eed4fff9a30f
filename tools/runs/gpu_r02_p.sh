#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
timeout 600 python -m pytest -q -p no:cacheprovider tests/test_lstm_gpu.py -m gpu -x > $O/r02p_lstm.log 2>&1
echo "lstm pytest rc=$?"; tail -5 $O/r02p_lstm.log
timeout 600 python bench.py --config cfg4 --steps 5 --warmup 3 --no-cpu-baseline > $O/r02p_cfg4.json 2> $O/r02p_cfg4.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r02p_cfg4.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['e2e']['ms_per_step'])
except Exception as e:
    print('bench parse failed', e)
PY
tail -3 $O/r02p_cfg4.err
