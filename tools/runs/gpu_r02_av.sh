#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
cp dnn-based_source_separation_b200/libctn_b200.so /tmp/orig.so
for v in 0 2 3 5; do
  cp gpurun_out_poly$v.so dnn-based_source_separation_b200/libctn_b200.so
  echo "== POLY=$v"
  timeout 300 python -m pytest -q -p no:cacheprovider tests/test_lstm_gpu.py -m gpu -x -k "fp64_oracle" 2>&1 | tail -1
  timeout 120 python tools/lstm_time.py 2>&1 | grep dbg | head -1
done
cp /tmp/orig.so dnn-based_source_separation_b200/libctn_b200.so
