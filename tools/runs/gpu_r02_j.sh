#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train-block > gpurun_out/r02j_$name.json 2> gpurun_out/r02j_$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r02j_{sys.argv[1]}.json").read().strip().splitlines()[-1]); st = d["stages"]
    print(f"{sys.argv[1]:10s} step {d['ms_per_step']:.3f} e2e {d['e2e']['ms_per_step']:.3f} pw1 {st['pw1']['ms_per_step']:.3f} pw2 {st['pw2']['ms_per_step']:.3f} mask {st['mask']['ms_per_step']:.3f} fin {st['fin']['ms_per_step']:.3f} loss {d['last_loss']}")
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
run g0 X=1
run g4 CTN_TCN_GROUP=4
run g8 CTN_TCN_GROUP=8
run g9 CTN_TCN_GROUP=9
run g16 CTN_TCN_GROUP=16
run g2 CTN_TCN_GROUP=2
CTN_TCN_GROUP=4 timeout 600 python -m pytest -q -p no:cacheprovider tests/test_parity_gpu.py -m gpu -x -k "cfg2_full_size or cfg5 or golden" 2>&1 | tail -3
