#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
timeout 900 python -m pytest -q -p no:cacheprovider tests/test_dprnn_gpu.py -m gpu > $O/r02am.log 2>&1
echo "dprnn pytest rc=$?"; tail -3 $O/r02am.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02am_launches.csv \
  python bench.py --config cfg4 --steps 1 --warmup 3 --no-cpu-baseline --no-lib-ab > $O/r02am_prof.log 2>&1
python tools/summarize_launches.py $O/r02am_launches.csv | head -8
timeout 600 python bench.py --config cfg4 --steps 10 --warmup 3 --no-cpu-baseline --no-lib-ab > $O/r02am_cfg4.json 2> $O/r02am_cfg4.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02am_cfg4.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['e2e']['ms_per_step'], d['clocks'])
PY
