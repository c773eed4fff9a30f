#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final3_n1_f16x3.json 2> gpurun_out/bench_final3.err
python bench.py --steps 10 --warmup 3 --math tf32x3 --no-cpu-baseline > gpurun_out/bench_final3_n1_tf32x3.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final3_reference.json 2>/dev/null
python bench.py --train --steps 10 --warmup 3 > gpurun_out/bench_final3_train_n1.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import json
for f in ("bench_final3_n1_f16x3", "bench_final3_n1_tf32x3", "bench_final3_reference", "bench_final3_train_n1"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 3), "ms", round(d["value"], 1), d.get("e2e", {}).get("value"), d.get("clocks"), (d.get("roofline") or {}).get("frac"), d.get("gpu_launches"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -2 gpurun_out/bench_final3.err
