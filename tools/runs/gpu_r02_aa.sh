#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
for d in 0 64; do CTN_LSTM_DBG=$d timeout 120 python tools/lstm_time.py 2>&1 | grep dbg | head -1; done
CTN_LSTM_DBG=80 timeout 120 python tools/lstm_time.py 2>&1 | grep "step 101"
