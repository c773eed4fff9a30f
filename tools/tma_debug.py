#!/usr/bin/env python
"""Debug helper: paper-size forward at several batch sizes; saves estimates so that runs under different CTN_* knobs can be diffed."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dnn-based_source_separation_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import convtasnet_oracle as O
from test_parity_gpu import build_model
tag = sys.argv[1]
cfg = O.OracleConfig()
sd = O.synth_state_dict(cfg, seed=111)
model = build_model(cfg, sd, math="f16x3")
mixture, sources = O.synth_batch(32, 2, 32000, seed=111)
xm = mixture.cuda()
res = {}
with torch.no_grad():
    for B in (32, 3, 8):
        idx = torch.arange(B) if B == 32 else (torch.tensor([5, 31, 0]) if B == 3 else torch.tensor([5, 31, 0, 7, 9, 11, 13, 2]))
        a = model(xm[idx].contiguous())
        b = model(xm[idx].contiguous())
        res[B] = a[:3].cpu() if B != 32 else a[[5, 31, 0]].cpu()
        print(tag, "B", B, "run-to-run max diff", float((a - b).abs().max()), flush=True)
torch.save(res, os.path.join(ROOT, "gpurun_out", f"dbg_{tag}.pt"))
for B in (3, 8):
    print(tag, "B", B, "vs B=32:", float((res[B] - res[32]).abs().max()))
