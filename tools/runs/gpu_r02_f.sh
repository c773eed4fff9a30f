#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
PT="python -m pytest -q -p no:cacheprovider --timeout 300 --timeout-method thread"
timeout 1800 $PT tests -m gpu > gpurun_out/r02f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02f_pytest.log
tail -12 gpurun_out/r02f_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02f_bench.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), {k: round(v["ms_per_step"], 3) for k, v in d.get("stages", {}).items()}, d["last_loss"])
print("train", d.get("train")); print("cpu", d.get("cpu_baseline"))
PY
