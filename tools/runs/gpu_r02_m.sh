#!/bin/bash
# native bi-LSTM kernel: unit parity, DPRNN model parity, cfg4 bench A/B against the cuDNN path
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
timeout 600 python -m pytest -q -p no:cacheprovider tests/test_lstm_gpu.py -m gpu -x > $O/r02m_lstm.log 2>&1
echo "lstm pytest rc=$?"; tail -15 $O/r02m_lstm.log
timeout 900 python -m pytest -q -p no:cacheprovider tests/test_dprnn_gpu.py -m gpu > $O/r02m_dprnn.log 2>&1
echo "dprnn pytest rc=$?"; tail -8 $O/r02m_dprnn.log
timeout 600 python bench.py --config cfg4 --steps 5 --warmup 3 --no-cpu-baseline > $O/r02m_cfg4.json 2> $O/r02m_cfg4.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r02m_cfg4.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['e2e']['ms_per_step'])
except Exception as e:
    print('bench parse failed', e)
PY
tail -3 $O/r02m_cfg4.err
