#!/usr/bin/env python
"""Tiny forward + loss + backward (both numeric modes, gLN and cLN models) for compute-sanitizer runs:
   compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python tools/sanitize.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dnn-based_source_separation_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import convtasnet_oracle as O
from ctn_b200.criterion.pit import PIT1d
from ctn_b200.criterion.sdr import NegSISDR
from test_parity_gpu import build_model

modes = sys.argv[1:] or ["tf32x3", "fp32"]
for causal in (False, True):
    cfg = O.OracleConfig(n_basis=32, kernel_size=16, sep_hidden_channels=64, sep_bottleneck_channels=32, sep_skip_channels=32,
                         sep_num_blocks=2, sep_num_layers=2, causal=causal, n_sources=2)
    sd = O.synth_state_dict(cfg, seed=1)
    mixture, sources = O.synth_batch(2, 2, 1500, seed=2)
    for mode in modes:
        model = build_model(cfg, sd, math=mode)
        crit = PIT1d(NegSISDR(), 2)
        with torch.no_grad():
            out = model(mixture.cuda())
            loss, perm = crit(out, sources.cuda())
        torch.cuda.synchronize()
        print("fwd", "cLN" if causal else "gLN", mode, float(loss), perm.tolist(), flush=True)
        if not causal:
            model.train()
            loss, _ = crit(model(mixture.cuda()), sources.cuda())
            loss.backward()
            torch.cuda.synchronize()
            print("bwd", mode, float(loss), float(sum(p.grad.abs().sum() for p in model.parameters())), flush=True)
