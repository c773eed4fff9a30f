#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pw_tma --launch-skip 195 --launch-count 1 -o $O/r02ap_maskdec \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train-block > $O/r02ap_prof.log 2>&1
python tools/ncu_summary.py $O/r02ap_maskdec.ncu-rep maskdec; python tools/ncu_roles.py $O/r02ap_maskdec.ncu-rep 0 > $O/r02ap_roles.txt 2>&1; head -40 $O/r02ap_roles.txt
