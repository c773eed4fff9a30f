#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 3 python tools/sanitize.py f16x3 > gpurun_out/r02_san_$tool.log 2>&1
  echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02_san_$tool.log | tail -1)"
done
