#!/bin/bash
# round-1 extra measurements (run on the GPU box): cfg5 batch sweep (HBM roofline), train bench lines, ncu capture of the wgrad kernel
mkdir -p gpurun_out
: > gpurun_out/cfg5_sweep.jsonl
for b in 1 2 4 8 16; do
  timeout 300 python bench.py --batch $b --n-sources 4 --seconds 8 --sample-rate 16000 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/cfg5_sweep.jsonl
done
timeout 300 python bench.py --train --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_train_n1.json
timeout 300 python bench.py --train --steps 10 --warmup 3 --batch 8 --n-sources 3 2>/dev/null | tail -1 > gpurun_out/bench_train_cfg3shape_n1.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wgrad_umma --launch-skip 60 -c 2 -o gpurun_out/prof_r01_wgrad python bench.py --train --steps 1 --warmup 1 > /dev/null 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/cfg5_sweep.jsonl"):
    try:
        d = json.loads(l); print("cfg5 B", d["config"]["global_batch"], "ms", round(d["ms_per_step"], 2), "audio-s/s", round(d["value"]), "dom", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3))
    except Exception as e:
        print("bad line", e)
for f in ("bench_train_n1", "bench_train_cfg3shape_n1"):
    d = json.loads(open(f"gpurun_out/{f}.json").read()); print(f, round(d["ms_per_step"], 2), "ms", round(d["value"]), "audio-s/s", round(d["peak_mem_gb"], 1), "GB")
PY
ls -la gpurun_out/*.ncu-rep | tail -3
