"""GPU parity of the TRAINING path (``-m gpu``): gradients of every parameter tensor from
ctn_convtasnet_fwd_train / ctn_convtasnet_bwd / ctn_sisdr_pit_bwd against torch autograd over the CPU oracle
(oracle/convtasnet_oracle.py), i.e. against what ``loss.backward()`` yields in the reference trainer
(egs/wsj0-mix/common/src/driver.py:146-150).

Tolerance: per tensor, max|g - g_ref| <= GRAD_RTOL * max|g_ref| + GRAD_ATOL.  The CPU autograd result itself moves by
~1e-6 relative between thread counts; weight gradients here are sums over B*frames terms accumulated in fp32 with
atomics across CTAs (order not fixed), so the stated bound is 2e-4 relative to the tensor's largest entry."""
import pytest
import torch

import convtasnet_oracle as O
from ctn_b200 import _native as N
from ctn_b200.criterion.pit import PIT1d
from ctn_b200.criterion.sdr import NegSISDR
from test_parity_gpu import build_model

pytestmark = pytest.mark.gpu

GRAD_RTOL, GRAD_ATOL = 2e-4, 1e-9
MODES = ["fp32"] + (["tf32x3", "f16x3"] if N.ctn_has_tcgen05() else [])


def _oracle_grads(cfg, sd, mixture, sources):
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out, _ = O.conv_tasnet_fwd(mixture, sd, cfg)
    loss, perm = O.pit_neg_sisdr(out, sources, batch_mean=True)
    loss.backward()
    return out.detach(), loss.detach(), perm, {k: v.grad for k, v in sd.items()}


def _check_grads(model, ref, rtol=GRAD_RTOL):
    worst = (0.0, None)
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        g, r = p.grad.detach().cpu(), ref[k]
        assert g.shape == r.shape, k
        scale = r.abs().max().item()
        err = (g - r).abs().max().item()
        rel = err / (scale + 1e-30)
        if rel > worst[0]:
            worst = (rel, k)
        assert err <= rtol * scale + GRAD_ATOL, "{}: max err {:.3e} vs max |ref| {:.3e} (rel {:.2e})".format(k, err, scale, rel)
    return worst


SHAPES = [
    dict(n_basis=24, kernel_size=8, sep_hidden_channels=40, sep_bottleneck_channels=20, sep_skip_channels=12,
         sep_num_blocks=2, sep_num_layers=3, n_sources=2),
    dict(n_basis=32, kernel_size=16, sep_hidden_channels=64, sep_bottleneck_channels=16, sep_skip_channels=16,
         sep_kernel_size=5, sep_num_blocks=1, sep_num_layers=4, n_sources=3),
    dict(n_basis=64, kernel_size=16, sep_hidden_channels=128, sep_bottleneck_channels=72, sep_skip_channels=40,
         sep_num_blocks=2, sep_num_layers=2, n_sources=2, enc_nonlinear='relu'),
    dict(n_basis=16, kernel_size=2, stride=1, sep_hidden_channels=32, sep_bottleneck_channels=16, sep_skip_channels=16,
         sep_num_blocks=1, sep_num_layers=1, n_sources=2),
]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("shape", SHAPES)
def test_model_gradients_vs_oracle_autograd(mode, shape):
    cfg = O.OracleConfig(causal=False, **shape)
    sd = O.synth_state_dict(cfg, seed=41)
    # non-trivial affine parameters / slopes so that every gradient path is exercised
    g = torch.Generator().manual_seed(5)
    for k in sd:
        if k.endswith("norm.weight"):
            sd[k] = 1.0 + 0.3 * torch.randn(sd[k].shape, generator=g)
        elif k.endswith("norm.bias"):
            sd[k] = 0.2 * torch.randn(sd[k].shape, generator=g)
    mixture, sources = O.synth_batch(3, cfg.n_sources, 1003, seed=42)
    ref_out, ref_loss, ref_perm, ref_grads = _oracle_grads(cfg, sd, mixture, sources)
    model = build_model(cfg, sd, math=mode).train()
    out = model(mixture.cuda())
    assert out.requires_grad
    torch.testing.assert_close(out.detach().cpu(), ref_out, rtol=1e-4, atol=2e-5)
    loss, perm = PIT1d(NegSISDR(), cfg.n_sources)(out, sources.cuda())
    assert torch.equal(perm.cpu(), ref_perm)
    torch.testing.assert_close(loss.detach().cpu(), ref_loss, rtol=0, atol=1e-4)
    loss.backward()
    worst = _check_grads(model, ref_grads)
    print("worst relative gradient error", worst)
    # the training forward and the inference forward are the same function
    with torch.no_grad():
        out_inf = model(mixture.cuda())
    torch.testing.assert_close(out.detach(), out_inf, rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("S,T", [(2, 4000), (3, 1003), (4, 517), (1, 64)])
def test_pit_backward_vs_autograd(S, T):
    g = torch.Generator().manual_seed(S * 1000 + T)
    est = torch.randn(5, S, T, generator=g)
    tgt = torch.randn(5, S, T, generator=g) + 0.5 * est[:, torch.randperm(S, generator=g)]
    e_ref = est.clone().requires_grad_(True)
    loss_ref, perm_ref = O.pit_neg_sisdr(e_ref, tgt, batch_mean=False)
    wts = torch.linspace(0.5, 1.5, 5)
    (loss_ref * wts).sum().backward()
    e = est.cuda().requires_grad_(True)
    loss_b, perm = PIT1d(NegSISDR(), S)(e, tgt.cuda(), batch_mean=False)
    assert torch.equal(perm.cpu(), perm_ref)
    (loss_b * wts.cuda()).sum().backward()
    scale = e_ref.grad.abs().max().item()
    torch.testing.assert_close(e.grad.cpu(), e_ref.grad, rtol=1e-4, atol=1e-5 * scale)


def test_training_step_decreases_loss():
    """Three SGD steps on one synthetic batch through the native forward/backward: the loss must go down."""
    cfg = O.OracleConfig(n_basis=32, kernel_size=16, sep_hidden_channels=64, sep_bottleneck_channels=32, sep_skip_channels=32,
                         sep_num_blocks=2, sep_num_layers=3, causal=False, n_sources=2)
    model = build_model(cfg, O.synth_state_dict(cfg, seed=3)).train()
    mixture, sources = O.synth_batch(4, 2, 4000, seed=8)
    mixture, sources = mixture.cuda(), sources.cuda()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    crit = PIT1d(NegSISDR(), 2)
    losses = []
    for _ in range(4):
        opt.zero_grad()
        loss, _ = crit(model(mixture), sources)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0)   # driver.py:152-153
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0], losses


PAPER = dict(n_basis=512, kernel_size=16, sep_hidden_channels=512, sep_bottleneck_channels=128, sep_skip_channels=128,
             sep_num_blocks=3, sep_num_layers=8)


@pytest.mark.parametrize("mode", [m for m in MODES if m != "fp32"])
@pytest.mark.parametrize("S", [2, 3])
def test_paper_size_gradients_vs_oracle_autograd(mode, S):
    """BASELINE hyper-parameters (N=512 L=16 B=128 H=512 Sc=128 X=8 R=3; cfg2 = 2 speakers, cfg3 = 3 speakers), batch 2,
    T = 8000: the tcgen05 weight-gradient kernel runs its 4 M-tiles / K = 512 shapes and the split-K red.add path.  All 343
    gradient tensors against torch autograd over the oracle (egs/wsj0-mix/common/src/driver.py:146-150)."""
    cfg = O.OracleConfig(causal=False, n_sources=S, **PAPER)
    sd = O.synth_state_dict(cfg, seed=113)
    mixture, sources = O.synth_batch(2, S, 8000, seed=113)
    ref_out, ref_loss, ref_perm, ref_grads = _oracle_grads(cfg, sd, mixture, sources)
    model = build_model(cfg, sd, math=mode).train()
    out = model(mixture.cuda())
    torch.testing.assert_close(out.detach().cpu(), ref_out, rtol=1e-4, atol=2e-5)
    loss, perm = PIT1d(NegSISDR(), S)(out, sources.cuda())
    assert torch.equal(perm.cpu(), ref_perm)
    torch.testing.assert_close(loss.detach().cpu(), ref_loss, rtol=0, atol=1e-4)
    loss.backward()
    assert len(ref_grads) == 343
    worst = _check_grads(model, ref_grads)
    print("paper-size worst relative gradient error", worst)


def test_paper_size_gradients_vs_reference_golden(golden_dir):
    """Same shape against the fixture minted from the UNMODIFIED reference's loss.backward() (tests/golden/make_golden.py
    grad_case): loss, permutation, and every 97th element + fp64 sum of each of the 343 gradient tensors."""
    import os
    rec = torch.load(os.path.join(golden_dir, "paper_3spk_grad.pt"), weights_only=False)
    cfg = O.OracleConfig(**rec["cfg"])
    sd = O.synth_state_dict(cfg, seed=rec["wseed"])
    mixture, sources = O.synth_batch(rec["batch"], cfg.n_sources, rec["T"], seed=rec["xseed"])
    model = build_model(cfg, sd).train()
    loss, perm = PIT1d(NegSISDR(), cfg.n_sources)(model(mixture.cuda()), sources.cuda())
    loss.backward()
    assert torch.equal(perm.cpu(), rec["perm"])
    torch.testing.assert_close(loss.detach().cpu(), rec["loss"], rtol=0, atol=1e-4)
    n = 0
    for k, p in model.named_parameters():
        g = rec["grads"][k]
        mine = p.grad.detach().cpu()
        assert tuple(mine.shape) == g["shape"], k
        tol = GRAD_RTOL * g["absmax"] + GRAD_ATOL
        assert float((mine.flatten()[::rec["stride"]] - g["sample"]).abs().max()) <= tol, k
        assert abs(float(mine.double().sum()) - g["sum"]) <= GRAD_RTOL * (g["sumsq"] * mine.numel()) ** 0.5 + 1e-9, k
        n += 1
    assert n == 343
