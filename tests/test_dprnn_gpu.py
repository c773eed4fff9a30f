"""GPU parity of the DPRNN-TasNet path (BASELINE cfg4; ``-m gpu``): Segment1d / OverlapAdd1d kernels, the dual-path glue and
the whole model through the Python mirror -> ctypes -> C ABI, against goldens minted from the unmodified reference
(tests/golden/dprnn_*.pt) and against the CPU oracle (oracle/dprnn_oracle.py).

Tolerance: the LSTM recurrences run in cuDNN on the GPU and in ATen's CPU kernels in the oracle / reference (both fp32; tanh /
sigmoid implementations differ at the 1e-7 level and 250-step recurrences amplify that), everything else in our kernels:
outputs rtol 1e-4 / atol 2e-5 x max|ref|, PIT permutation exact, loss 1e-3 dB."""
import os

import pytest
import torch

import convtasnet_oracle as O
import dprnn_oracle as DO
from ctn_b200 import _native as N
from ctn_b200.criterion.pit import PIT1d
from ctn_b200.criterion.sdr import NegSISDR
from ctn_b200.models.dprnn_tasnet import DPRNNTasNet
from ctn_b200.models.dprnn import DPRNN
from ctn_b200.models.transform import Segment1d, OverlapAdd1d

pytestmark = pytest.mark.gpu
MODES = ["fp32"] + (["tf32x3", "f16x3"] if N.ctn_has_tcgen05() else [])


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)


def build(cfg, sd, math=None):
    m = DPRNNTasNet(cfg.n_basis, cfg.kernel_size, stride=cfg.stride, enc_basis="trainable", dec_basis="trainable", enc_nonlinear=cfg.enc_nonlinear,
                    sep_hidden_channels=cfg.sep_hidden_channels, sep_bottleneck_channels=cfg.sep_bottleneck_channels,
                    sep_chunk_size=cfg.sep_chunk_size, sep_hop_size=cfg.sep_hop_size, sep_num_blocks=cfg.sep_num_blocks, sep_norm=True,
                    mask_nonlinear="sigmoid", causal=False, rnn_type="lstm", n_sources=cfg.n_sources, eps=cfg.eps)
    m.load_state_dict(sd, strict=True)
    m.math = math
    return m.cuda().eval()


def test_segment_overlap_add_golden(golden_dir):
    rec = _load(golden_dir, "dprnn_modules")
    for key, r in rec.items():
        _, B, Fc, T, K, P = key.split("_")
        seg = Segment1d(int(K), int(P))(r["x"].cuda())
        assert torch.equal(seg.cpu(), r["seg"]), key                       # pure data movement: bit-exact
        ola = OverlapAdd1d(int(K), int(P))(seg)
        torch.testing.assert_close(ola.cpu(), r["ola"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("B,Fc,T,K,P", [(3, 64, 32000, 250, 125), (2, 5, 1000, 100, 50), (1, 33, 777, 64, 16), (2, 8, 50, 50, 25), (2, 7, 90, 20, 30)])
def test_segment_overlap_add_vs_oracle(B, Fc, T, K, P):
    g = torch.Generator().manual_seed(T)
    x = torch.randn(B, Fc, T, generator=g)
    seg = Segment1d(K, P)(x.cuda())
    ref = DO.segment1d(x, K, P)
    assert torch.equal(seg.cpu(), ref)
    if P <= K:
        torch.testing.assert_close(OverlapAdd1d(K, P)(seg).cpu(), DO.overlap_add1d(ref, K, P), rtol=0, atol=1e-6)


def test_dprnn_stack_vs_oracle():
    """DPRNN.forward in the reference layout (B, F, S, K): permutes + cuDNN LSTM + library GEMM + native gLN/residual/swap"""
    cfg = DO.DPRNNConfig(n_basis=16, kernel_size=4, sep_hidden_channels=24, sep_bottleneck_channels=16, sep_chunk_size=20, sep_hop_size=10,
                         sep_num_blocks=3)
    sd = DO.synth_state_dict(cfg, seed=5)
    sub = {k[len("separator.dprnn."):]: v for k, v in sd.items() if k.startswith("separator.dprnn.")}
    net = DPRNN(16, 24, num_blocks=3, causal=False)
    net.load_state_dict(sub, strict=True)
    net = net.cuda().eval()
    x = torch.randn(3, 16, 9, 20, generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        y = net(x.cuda())
        ref = DO.dprnn_fwd(x, sd, "separator.dprnn.", 3, cfg.eps)
    torch.testing.assert_close(y.cpu(), ref, rtol=1e-4, atol=2e-5 * float(ref.abs().max()))


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name", ["dprnn_tiny", "dprnn_cfg4_short"])
def test_dprnn_tasnet_golden(golden_dir, name, mode):
    rec = _load(golden_dir, name)
    cfg = DO.DPRNNConfig(**rec["cfg"])
    sd = DO.synth_state_dict(cfg, seed=rec["wseed"])
    model = build(cfg, sd, math=mode)
    mixture, sources = O.synth_batch(rec["batch"], cfg.n_sources, rec["T"], seed=rec["xseed"])
    with torch.no_grad():
        out, latent = model.extract_latent(mixture.cuda())
        loss, perm = PIT1d(NegSISDR(), cfg.n_sources)(out, sources.cuda())
    assert torch.equal(perm.cpu(), rec["perm"])
    torch.testing.assert_close(loss.cpu(), rec["loss"], rtol=0, atol=1e-3)
    so = rec.get("out_stride")
    scale = rec["out_absmax"]
    if so is None:
        torch.testing.assert_close(out.cpu(), rec["out"], rtol=1e-4, atol=2e-5 * max(1.0, scale))
        torch.testing.assert_close(latent.cpu(), rec["latent"], rtol=1e-4, atol=2e-5 * max(1.0, float(rec["latent"].abs().max())))
    else:
        torch.testing.assert_close(out.cpu()[..., ::so], rec["out"], rtol=1e-4, atol=2e-5 * max(1.0, scale))
        a, b = rec["latent_stride"]
        torch.testing.assert_close(latent.cpu()[:, :, ::a, ::b], rec["latent"], rtol=1e-4, atol=2e-5 * max(1.0, float(rec["latent"].abs().max())))


def test_dprnn_cfg4_full_batch_vs_oracle():
    """BASELINE cfg4: N=64 L=2 F=64 H=128 K=250 P=125 B=6, batch 16 x 4 s @ 8 kHz.  All 16 mixtures run on the GPU; the CPU
    oracle checks the first 2 (it needs ~10 s per mixture pair), the rest through size-independent properties (batch
    independence: sample i of the batch-16 run == the same sample run alone)."""
    cfg = DO.DPRNNConfig(n_basis=64, kernel_size=2, sep_hidden_channels=128, sep_bottleneck_channels=64, sep_chunk_size=250, sep_hop_size=125,
                         sep_num_blocks=6, n_sources=2)
    sd = DO.synth_state_dict(cfg, seed=44)
    model = build(cfg, sd)
    mixture, sources = O.synth_batch(16, 2, 32000, seed=45)
    with torch.no_grad():
        out = model(mixture.cuda())
        loss_b, perm = PIT1d(NegSISDR(), 2)(out, sources.cuda(), batch_mean=False)
        ref, _ = DO.dprnn_tasnet_fwd(mixture[:2], sd, cfg)
        ref_loss, ref_perm = O.pit_neg_sisdr(ref, sources[:2], batch_mean=False)
        alone = model(mixture[9:10].cuda())
    assert out.shape == (16, 2, 32000) and torch.isfinite(out).all()
    torch.testing.assert_close(out[:2].cpu(), ref, rtol=1e-4, atol=2e-5 * max(1.0, float(ref.abs().max())))
    assert torch.equal(perm[:2].cpu(), ref_perm)
    torch.testing.assert_close(loss_b[:2].cpu(), ref_loss, rtol=0, atol=1e-3)
    torch.testing.assert_close(out[9:10], alone, rtol=1e-4, atol=2e-5)


def test_dprnn_envelope_errors():
    with pytest.raises(NotImplementedError):
        DPRNNTasNet(16, 4, enc_basis="trainable", dec_basis="trainable", enc_nonlinear=None, causal=True)
    with pytest.raises(NotImplementedError):
        DPRNNTasNet(16, 4, enc_basis="trainable", dec_basis="trainable", enc_nonlinear=None, causal=False, mask_nonlinear="softmax")
    m = DPRNNTasNet(16, 4, enc_basis="trainable", dec_basis="trainable", enc_nonlinear=None, causal=False).cuda()
    with pytest.raises(NotImplementedError):
        m(torch.randn(1, 1, 64, device="cuda"))        # autograd enabled: forward-only path
    with pytest.raises(ValueError):
        with torch.no_grad():
            m(torch.randn(1, 64, device="cuda"))
