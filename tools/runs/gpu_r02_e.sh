#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
PT="python -m pytest -q -p no:cacheprovider --timeout 300 --timeout-method thread"
CUDA_LAUNCH_BLOCKING=1 timeout 600 $PT tests/test_parity_gpu.py -m gpu -x -k "fused_mask or host_buffer or robust" > gpurun_out/r02e_blocking.log 2>&1
tail -30 gpurun_out/r02e_blocking.log
timeout 900 $PT tests/test_train_gpu.py -m gpu -s -k "paper_size or clip_adam" > gpurun_out/r02e_train.log 2>&1
grep -n "paper-size\|AssertionError\|passed\|failed" gpurun_out/r02e_train.log | head -20
timeout 900 $PT tests/test_parity_gpu.py -m gpu -k "f16x3" > gpurun_out/r02e_f16.log 2>&1
tail -5 gpurun_out/r02e_f16.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-block > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02e_bench.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), {k: round(v["ms_per_step"], 3) for k, v in d.get("stages", {}).items()}, d["last_loss"])
PY
