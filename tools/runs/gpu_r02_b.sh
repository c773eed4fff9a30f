#!/bin/bash
# round-2 GPU call B: parity of the TMA-fed kernels + new tests, bench with / without them, launch list and full ncu captures
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
PT="python -m pytest -q -p no:cacheprovider --timeout 300 --timeout-method thread"
timeout 900 $PT tests/test_parity_gpu.py -m gpu -x -k "f16x3" > gpurun_out/r02b_tma_f16.log 2>&1
rc=$?; echo "pytest rc=$rc" >> gpurun_out/r02b_tma_f16.log
if [ $rc -ne 0 ]; then echo "TMA kernels FAILED -> CTN_PW_TMA=0 for the suite" | tee -a gpurun_out/r02b_tma_f16.log; export CTN_PW_TMA=0; fi
timeout 1500 $PT tests -m gpu > gpurun_out/r02b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
unset CTN_PW_TMA
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
CTN_PW_TMA=0 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-block > gpurun_out/r02b_bench_old.json 2> gpurun_out/r02b_bench_old.err
timeout 300 python bench.py --config cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_cfg4.json 2> gpurun_out/r02b_bench_cfg4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02b_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train-block > gpurun_out/r02b_ncu_bench.log 2>&1
# full captures: pw2 (block 5, dilation 32 -> launch index ~ 5 within the k_pw_tma<2,0> launches), pw1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pw_tma --launch-skip 10 --launch-count 2 \
  -o gpurun_out/r02b_pw_tma python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train-block > gpurun_out/r02b_ncu_full.log 2>&1
tail -4 gpurun_out/r02b_tma_f16.log; tail -8 gpurun_out/r02b_pytest.log
python - <<'PY'
import json
for f in ("gpurun_out/r02b_bench.json", "gpurun_out/r02b_bench_old.json", "gpurun_out/r02b_bench_cfg4.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), {k: round(v["ms_per_step"], 3) for k, v in d.get("stages", {}).items()})
        if "train" in d: print("  train", d["train"])
    except Exception as e:
        print(f, "unreadable", e)
PY
