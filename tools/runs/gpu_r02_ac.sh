#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
for pc in 0 56 64 70 76 84 96; do echo "pace $pc"; CTN_LSTM_DBG=$((pc * 4096)) timeout 120 python tools/lstm_time.py 2>&1 | grep dbg | head -1; done
echo "pace 76 timeline"; CTN_LSTM_DBG=$((16 + 76 * 4096)) timeout 120 python tools/lstm_time.py 2>&1 | grep "step 101 chunk [014]"
