#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
timeout 1800 python -m pytest -q -p no:cacheprovider tests -m gpu > $O/r02ax_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $O/r02ax_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02ax_bench.json 2> $O/r02ax_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02ax_bench.json').read().strip().splitlines()[-1])
st = d['stages']
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['e2e']['ms_per_step'], d['clocks'], d['train']['ms_per_step'])
print({k: round(v['ms_per_step'], 3) for k, v in st.items()})
PY
