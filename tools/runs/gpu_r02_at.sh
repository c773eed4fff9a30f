#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 compute-sanitizer --tool initcheck --print-limit 8 python tools/sanitize.py f16x3 > gpurun_out/r02at_initcheck.log 2>&1
grep -E "ERROR SUMMARY|Uninitialized" gpurun_out/r02at_initcheck.log | sort | uniq -c | head
grep -B2 -A14 "Uninitialized" gpurun_out/r02at_initcheck.log | head -60
