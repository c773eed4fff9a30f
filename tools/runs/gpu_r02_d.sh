#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
PT="python -m pytest -q -p no:cacheprovider --timeout 300 --timeout-method thread"
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest -q -p no:cacheprovider tests/test_parity_gpu.py -m gpu -x -k "fused_mask and 8000" > gpurun_out/r02d_sanitizer.log 2>&1
grep -n "Invalid\|misaligned\|at 0x\|by thread\|Address\|=========     at\|ctn_pwtma.cu" gpurun_out/r02d_sanitizer.log | head -30
timeout 900 $PT tests/test_train_gpu.py -m gpu -k "paper_size or clip_adam" > gpurun_out/r02d_train.log 2>&1
tail -15 gpurun_out/r02d_train.log
grep -n "AssertionError" gpurun_out/r02d_train.log | head
