#!/bin/bash
# knob matrix on the TMA-fed kernels: where does the time go?  (stage timers from bench.py's separate pass)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train-block > gpurun_out/r02h_$name.json 2> gpurun_out/r02h_$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r02h_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    st = d["stages"]
    print(f"{sys.argv[1]:14s} step {d['ms_per_step']:.3f}  pw1 {st['pw1']['ms_per_step']:.3f}  pw2 {st['pw2']['ms_per_step']:.3f}  mask {st['mask']['ms_per_step']:.3f}  enc {st['enc']['ms_per_step']:.3f} fin {st['fin']['ms_per_step']:.3f}")
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
run base X=1
run op2 CTN_TMA_OPSTAGES=2
run raw2 CTN_TMA_RAWSTAGES=2
run raw3 CTN_TMA_RAWSTAGES=3
run nomma CTN_UMMA_DBG=8
run noloads CTN_UMMA_DBG=2
run nostores CTN_UMMA_DBG=1
run nomma_nostores CTN_UMMA_DBG=9
run enc_old CTN_ENC_V4=0
