#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
timeout 900 python -m pytest -q -p no:cacheprovider tests/test_lstm_gpu.py tests/test_train_gpu.py tests/test_parity_gpu.py -m gpu -x -k "lstm or autograd_node or model_golden or training_step" > $O/r02ah.log 2>&1
echo "pytest rc=$?"; tail -6 $O/r02ah.log
