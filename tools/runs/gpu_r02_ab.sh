#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
for w in 0 1 2 4 5; do echo "epilogue warp $w"; CTN_LSTM_DBG=$((16 + 256 * w)) timeout 120 python tools/lstm_time.py 2>&1 | grep "step 101 chunk [01]"; done
