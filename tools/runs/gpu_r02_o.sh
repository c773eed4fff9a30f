#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bilstm --launch-skip 12 --launch-count 2 -o $O/r02o_lstm \
  python bench.py --config cfg4 --steps 1 --warmup 1 --no-cpu-baseline > $O/r02o_prof.log 2>&1
python tools/ncu_summary.py $O/r02o_lstm.ncu-rep lstm > $O/r02o_lstm.md 2>&1; cat $O/r02o_lstm.md
python tools/ncu_roles.py $O/r02o_lstm.ncu-rep 0 > $O/r02o_roles0.txt 2>&1
python tools/ncu_roles.py $O/r02o_lstm.ncu-rep 1 > $O/r02o_roles1.txt 2>&1
head -45 $O/r02o_roles0.txt
du -sm $O
