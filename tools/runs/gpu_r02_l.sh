#!/bin/bash
# pair-mode (cta_group::2) TMA kernels: parity, then bench pair on / off
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
PT="python -m pytest -q -p no:cacheprovider --timeout 200 --timeout-method thread"
timeout 900 $PT tests/test_parity_gpu.py tests/test_blocks_gpu.py -m gpu -x -k "f16x3 or fused or block" > gpurun_out/r02l_f16.log 2>&1
rc=$?; echo "rc=$rc" >> gpurun_out/r02l_f16.log; tail -6 gpurun_out/r02l_f16.log
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train-block > gpurun_out/r02l_$name.json 2> gpurun_out/r02l_$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r02l_{sys.argv[1]}.json").read().strip().splitlines()[-1]); st = d["stages"]
    print(f"{sys.argv[1]:10s} step {d['ms_per_step']:.3f} e2e {d['e2e']['ms_per_step']:.3f} pw1 {st['pw1']['ms_per_step']:.3f} pw2 {st['pw2']['ms_per_step']:.3f} mask {st['mask']['ms_per_step']:.3f} fin {st['fin']['ms_per_step']:.3f} loss {d['last_loss']}")
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
if [ $rc -eq 0 ]; then run pair X=1; fi
run nopair CTN_TMA_PAIR=0
if [ $rc -eq 0 ]; then run pair_op5 CTN_TMA_OPSTAGES=5; run pair_op3 CTN_TMA_OPSTAGES=3; fi
# training forward through the fused kernels: gradient parity (toy + paper size), then the train bench fused / un-fused
timeout 900 $PT tests/test_train_gpu.py -m gpu -s > gpurun_out/r02l_train.log 2>&1
echo "rc=$?" >> gpurun_out/r02l_train.log; grep -n "paper-size\|passed\|failed\|Error" gpurun_out/r02l_train.log | cut -c1-260 | tail -12
for v in 0 1; do CTN_TRAIN_UNFUSED=$v timeout 300 python bench.py --train --n-sources 3 --batch 8 --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train unfused=$v cfg3 b8', round(d['ms_per_step'],3), 'ms', d['train']['gpu_launches_per_step'], 'launches', d['train']['last_loss'])"; done
for v in 0 1; do CTN_TRAIN_UNFUSED=$v timeout 300 python bench.py --train --n-sources 2 --batch 32 --steps 5 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train unfused=$v cfg2 b32', round(d['ms_per_step'],3), 'ms', d['train']['peak_mem_gb'], 'GB')"; done
