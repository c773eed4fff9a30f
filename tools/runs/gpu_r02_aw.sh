#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest -q -p no:cacheprovider tests/test_parity_gpu.py tests/test_host_api.py tests/test_train_gpu.py -m gpu -x -k "multichannel or model_golden or host or training_step or encoder_decoder" > gpurun_out/r02aw.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/r02aw.log
