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

modes = sys.argv[1:] or ["f16x3", "tf32x3", "fp32"]
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

# round 2: fused mask + decoder epilogue (N = 512, crop offset != 0 and an unaligned T), block-level forward, DPRNN glue
cfg = O.OracleConfig(n_basis=512, kernel_size=16, sep_hidden_channels=32, sep_bottleneck_channels=16, sep_skip_channels=16,
                     sep_num_blocks=1, sep_num_layers=2, causal=False, n_sources=2)
model = build_model(cfg, O.synth_state_dict(cfg, seed=3), math="f16x3")
mixture, _ = O.synth_batch(2, 2, 1031, seed=4)
with torch.no_grad():
    out = model(mixture.cuda())
torch.cuda.synchronize()
print("maskdec", float(out.abs().sum()), flush=True)
import dprnn_oracle as DO
from ctn_b200.models.dprnn_tasnet import DPRNNTasNet
dc = DO.DPRNNConfig(n_basis=16, kernel_size=4, sep_hidden_channels=12, sep_bottleneck_channels=8, sep_chunk_size=10, sep_hop_size=5, sep_num_blocks=1)
dm = DPRNNTasNet(16, 4, enc_basis="trainable", dec_basis="trainable", enc_nonlinear=None, sep_hidden_channels=12, sep_bottleneck_channels=8,
                 sep_chunk_size=10, sep_hop_size=5, sep_num_blocks=1, causal=False)
dm.load_state_dict(DO.synth_state_dict(dc, seed=5))
dm = dm.cuda().eval()
with torch.no_grad():
    o = dm(torch.randn(2, 1, 203, generator=torch.Generator().manual_seed(6)).cuda())
torch.cuda.synchronize()
print("dprnn", float(o.abs().sum()), flush=True)
# tcgen05 bi-LSTM + projection: the 2-CTA cluster kernel (F = 64, H = 128) and the 1-CTA kernel (H = 32), ragged row groups
from ctn_b200 import _native as NN
for Fi, H_, nseq, tt in ((64, 128, 70, 5), (32, 32, 40, 4)):
    g_ = torch.Generator().manual_seed(7)
    ws_ = [((torch.rand(s_, generator=g_) * 2 - 1) * 0.1).cuda() for s_ in ((4 * H_, Fi), (4 * H_, H_), (4 * H_,), (4 * H_,)) * 2]
    fc_ = ((torch.rand(Fi, 2 * H_, generator=g_) * 2 - 1) * 0.1).cuda()
    z_ = torch.randn(nseq, tt, Fi, generator=g_).cuda()
    nb_ = NN.ctn_bilstm_workspace_bytes(Fi, H_, Fi)
    wsb_ = torch.empty(nb_, dtype=torch.uint8, device="cuda")
    P_ = torch.empty(2, nseq, tt, Fi, device="cuda")
    ho_ = torch.empty(nseq, tt, 2 * H_, device="cuda")
    ptrs_ = (NN._fp * 8)(*[t_.data_ptr() for t_ in ws_])
    NN.check(NN.ctn_bilstm_proj_fwd(z_.data_ptr(), nseq, tt, Fi, H_, ptrs_, fc_.data_ptr(), Fi, P_.data_ptr(), ho_.data_ptr(), None, wsb_.data_ptr(), nb_,
                                    NN.stream_ptr(z_.device)), "ctn_bilstm_proj_fwd")
    torch.cuda.synchronize()
    print("lstm", Fi, H_, float(P_.abs().sum()), float(ho_.abs().sum()), flush=True)
