#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_f16_n1.json 2> gpurun_out/bench_f16.err
python bench.py --train --steps 10 --warmup 3 > gpurun_out/bench_f16_train_n1.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_f16.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pw_umma --launch-skip 10 -c 1 -o gpurun_out/prof_r01_f16_pw2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pw_umma --launch-skip 9 -c 1 -o gpurun_out/prof_r01_f16_pw1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import json
for f in ("bench_f16_n1", "bench_f16_train_n1"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 3), "ms", round(d["value"], 1), d.get("e2e", {}).get("value"), d.get("dtype"), d.get("clocks"), d.get("roofline"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -3 gpurun_out/bench_f16.err
