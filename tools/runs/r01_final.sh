#!/bin/bash
# final round-1 measurements (GPU box): bench lines of the three numeric modes + reference arm + launch list + full captures
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final2_n1_tf32x3.json 2> gpurun_out/bench_final2.err
python bench.py --steps 10 --warmup 3 --math tf32 --no-cpu-baseline > gpurun_out/bench_final2_n1_tf32.json 2>/dev/null
python bench.py --steps 5 --warmup 3 --math fp32 --no-cpu-baseline > gpurun_out/bench_final2_n1_fp32.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final2_reference.json 2>/dev/null
python bench.py --train --steps 10 --warmup 3 > gpurun_out/bench_final2_train_n1.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import json
for f in ("bench_final2_n1_tf32x3", "bench_final2_n1_tf32", "bench_final2_n1_fp32", "bench_final2_reference", "bench_final2_train_n1"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 3), "ms", round(d["value"], 1), d.get("e2e", {}).get("value"), d.get("clocks"))
    except Exception as e:
        print(f, "FAILED", e)
PY
