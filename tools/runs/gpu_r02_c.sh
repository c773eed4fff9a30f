#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
CTN_PW_TMA=0 timeout 200 python tools/tma_debug.py old > gpurun_out/r02c_dbg.log 2>&1
timeout 200 python tools/tma_debug.py tma >> gpurun_out/r02c_dbg.log 2>&1
CTN_TMA_RAWSTAGES=2 timeout 200 python tools/tma_debug.py tma_raw2 >> gpurun_out/r02c_dbg.log 2>&1
CTN_MASKDEC=0 timeout 200 python tools/tma_debug.py tma_nomaskdec >> gpurun_out/r02c_dbg.log 2>&1
python - >> gpurun_out/r02c_dbg.log 2>&1 <<'PY'
import torch
o = torch.load("gpurun_out/dbg_old.pt")
for t in ("tma", "tma_raw2", "tma_nomaskdec"):
    try:
        n = torch.load(f"gpurun_out/dbg_{t}.pt")
        for B in (32, 3, 8):
            d = (n[B] - o[B]).abs()
            print(t, "vs old, B", B, "max", float(d.max()), "per-sample max", [float(x) for x in d.amax(dim=(1, 2))])
    except Exception as e:
        print(t, e)
PY
cat gpurun_out/r02c_dbg.log
PT="python -m pytest -q -p no:cacheprovider --timeout 300 --timeout-method thread"
timeout 900 $PT tests/test_parity_gpu.py tests/test_train_gpu.py -m gpu -k "fused_mask or host_buffer or paper_size or clip_adam or robust" > gpurun_out/r02c_pytest.log 2>&1
tail -25 gpurun_out/r02c_pytest.log
