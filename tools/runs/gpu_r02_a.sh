#!/bin/bash
# round-2 GPU call A: (1) the TMA-fed fp16-piece kernels against the parity suite (short per-test timeout: a hang costs
# minutes, not the call), (2) the full -m gpu suite, (3) bench lines with and without the new kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi -L > gpurun_out/r02a_gpus.txt 2>&1
PT="python -m pytest -q -p no:cacheprovider --timeout 240 --timeout-method thread"
timeout 900 $PT tests/test_parity_gpu.py -m gpu -x -k "f16x3" > gpurun_out/r02a_tma_f16.log 2>&1
rc=$?; echo "pytest rc=$rc" >> gpurun_out/r02a_tma_f16.log
if [ $rc -ne 0 ]; then echo "TMA kernels FAILED -> CTN_PW_TMA=0 for the rest" | tee -a gpurun_out/r02a_tma_f16.log; export CTN_PW_TMA=0; fi
timeout 1500 $PT tests -m gpu > gpurun_out/r02a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
CTN_PW_TMA=0 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02a_bench_old.json 2> gpurun_out/r02a_bench_old.err
tail -4 gpurun_out/r02a_tma_f16.log; tail -4 gpurun_out/r02a_pytest.log
python - <<'PY'
import json
for f in ("gpurun_out/r02a_bench.json", "gpurun_out/r02a_bench_old.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], {k: round(v["ms_per_step"], 3) for k, v in d["stages"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
