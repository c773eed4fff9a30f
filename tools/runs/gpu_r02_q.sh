#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
for d in 0 1 2 4 8 3 5 6 7 15; do CTN_LSTM_DBG=$d timeout 120 python tools/lstm_time.py 2>&1 | grep dbg | head -2; done
for s in 4 6 8; do CTN_LSTM_STAGES=$s timeout 120 python tools/lstm_time.py 2>&1 | grep dbg | head -1; done
