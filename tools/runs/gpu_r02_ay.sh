#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
for c in cfg3 cfg5; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-train-block > gpurun_out/r02ay_$c.json 2> gpurun_out/r02ay_$c.err
  echo "$c rc=$?"; python - "$c" <<'PY'
import json, sys
try:
    d = json.loads(open(f'gpurun_out/r02ay_{sys.argv[1]}.json').read().strip().splitlines()[-1])
    print(d['config']['workload'][:110], {k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['e2e']['ms_per_step'])
except Exception as e:
    print('failed', e); print(open(f'gpurun_out/r02ay_{sys.argv[1]}.err').read()[-600:])
PY
done
timeout 300 python bench.py --train --steps 5 --warmup 2 2>/dev/null | tail -1 | cut -c1-300
