#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 600 python -m pytest -q -p no:cacheprovider tests/test_lstm_gpu.py -m gpu -x > gpurun_out/r02al.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02al.log
timeout 120 python tools/lstm_time.py 2>&1 | grep dbg
