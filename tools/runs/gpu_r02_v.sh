#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
CTN_LSTM_DBG=16 timeout 120 python tools/lstm_time.py 2>&1 | grep "step 101"
CTN_LSTM_DBG=48 timeout 120 python tools/lstm_time.py 2>&1 | grep "step 101"
for d in 0 32; do CTN_LSTM_DBG=$d timeout 120 python tools/lstm_time.py 2>&1 | grep dbg | head -1; done
