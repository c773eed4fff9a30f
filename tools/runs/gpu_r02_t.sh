#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
timeout 300 python -m pytest -q -p no:cacheprovider tests/test_lstm_gpu.py -m gpu -x > $O/r02t_lstm.log 2>&1
echo "lstm pytest rc=$?"; tail -12 $O/r02t_lstm.log
timeout 120 python tools/lstm_time.py 2>&1 | grep dbg
CTN_LSTM_PAIR=0 timeout 120 python tools/lstm_time.py 2>&1 | grep dbg | head -1
