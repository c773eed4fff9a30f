#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
timeout 600 python -m pytest -q -p no:cacheprovider tests/test_lstm_gpu.py -m gpu -x > $O/r02s_lstm.log 2>&1
echo "lstm pytest rc=$?"; tail -5 $O/r02s_lstm.log
timeout 120 python tools/lstm_time.py 2>&1 | grep dbg
CTN_LSTM_DBG=16 timeout 120 python tools/lstm_time.py 2>&1 | grep "step 10[12]"
