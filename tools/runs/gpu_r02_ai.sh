#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest -q -p no:cacheprovider tests/test_lstm_gpu.py -m gpu -x > gpurun_out/r02ai.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r02ai.log
