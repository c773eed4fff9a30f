"""Equality of the TMA-fed / fused path against the round-1 kernels (CTN_PW_TMA=0 CTN_MASKDEC=0) at the other BASELINE shapes."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dnn-based_source_separation_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch, convtasnet_oracle as O
    from ctn_b200.models.conv_tasnet import ConvTasNet
    S, T, B = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    cfg = O.OracleConfig(n_basis=512, kernel_size=16, sep_hidden_channels=512, sep_bottleneck_channels=128, sep_skip_channels=128,
                         sep_num_blocks=3, sep_num_layers=8, causal=False, n_sources=S)
    m = ConvTasNet(512, 16, enc_basis="trainable", dec_basis="trainable", enc_nonlinear=None, sep_hidden_channels=512, sep_bottleneck_channels=128,
                   sep_skip_channels=128, sep_num_blocks=3, sep_num_layers=8, causal=False, n_sources=S)
    m.load_state_dict(O.synth_state_dict(cfg, seed=5)); m = m.cuda().eval()
    x, _ = O.synth_batch(B, S, T, seed=6)
    with torch.no_grad():
        y = m(x.cuda())
    torch.save(y.cpu(), sys.argv[5])
    print("ok", tuple(y.shape), float(y.abs().max()))
    sys.exit(0)
import torch
for S, T, B in ((3, 32000, 3), (4, 128000, 2), (2, 31999, 5), (2, 4001, 33)):
    outs = []
    for tag, env in (("new", {}), ("old", {"CTN_PW_TMA": "0", "CTN_MASKDEC": "0"})):
        f = f"/tmp/ab_{tag}.pt"
        e = dict(os.environ); e.update(env)
        r = subprocess.run([sys.executable, __file__, "child", str(S), str(T), str(B), f], env=e, capture_output=True, text=True, timeout=600)
        if r.returncode != 0:
            print("FAILED", S, T, B, tag, r.stderr[-500:]); sys.exit(1)
        outs.append(torch.load(f))
    d = float((outs[0] - outs[1]).abs().max())
    print(f"S={S} T={T} B={B}: max|new - old| = {d:.3e} (max|y| {float(outs[1].abs().max()):.3f})", flush=True)
    assert d <= 2e-5 * max(1.0, float(outs[1].abs().max()))
print("all equal")
