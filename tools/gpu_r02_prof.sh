#!/bin/bash
# profile evidence: launch lists (inference step, train step, cfg4) and one `ncu --set full` capture of one steady-state forward
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_fwd.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train-block > gpurun_out/r02_prof_a.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 --csv --log-file gpurun_out/r02_launches_train.csv \
  python bench.py --train --n-sources 3 --batch 8 --steps 1 --warmup 2 > gpurun_out/r02_prof_b.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_cfg4.csv \
  python bench.py --config cfg4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_prof_c.log 2>&1
# one whole steady-state forward (launches 128.. of the process = the third step), every kernel, full set
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 128 --launch-count 64 -o gpurun_out/r02_full_fwd \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train-block > gpurun_out/r02_prof_d.log 2>&1
ls -la gpurun_out/ | grep r02_
