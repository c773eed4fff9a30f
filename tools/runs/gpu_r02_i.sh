#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
PT="python -m pytest -q -p no:cacheprovider --timeout 300 --timeout-method thread"
timeout 900 $PT tests/test_parity_gpu.py tests/test_blocks_gpu.py -m gpu -x -k "f16x3 or fused or checkpoint or sisdr_autograd or block" > gpurun_out/r02i_f16.log 2>&1
echo "rc=$?" >> gpurun_out/r02i_f16.log; tail -4 gpurun_out/r02i_f16.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-block > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02i_bench.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), {k: round(v["ms_per_step"], 3) for k, v in d.get("stages", {}).items()}, d["last_loss"])
PY
for k in nomma:8 nomma_nostores:9; do n=${k%%:*}; v=${k##*:}; CTN_UMMA_DBG=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train-block > gpurun_out/r02i_$n.json 2>/dev/null
python - $n <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/r02i_{sys.argv[1]}.json").read().strip().splitlines()[-1]); st = d["stages"]
print(sys.argv[1], round(d["ms_per_step"], 3), "pw1", round(st["pw1"]["ms_per_step"], 3), "pw2", round(st["pw2"]["ms_per_step"], 3))
PY
done
