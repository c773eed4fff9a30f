#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
timeout 600 python -m pytest -q -p no:cacheprovider tests/test_lstm_gpu.py -m gpu > $O/r02n_lstm.log 2>&1
echo "lstm pytest rc=$?"; tail -5 $O/r02n_lstm.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r02n_launches_cfg4.csv \
  python bench.py --config cfg4 --steps 1 --warmup 3 --no-cpu-baseline > $O/r02n_prof.log 2>&1
python tools/summarize_launches.py $O/r02n_launches_cfg4.csv | head -30
