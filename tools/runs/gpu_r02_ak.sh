#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 600 python -m pytest -q -p no:cacheprovider tests/test_parity_gpu.py tests/test_host_api.py -m gpu -x -k "sdr or sisdr or host" > gpurun_out/r02ak.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r02ak.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
