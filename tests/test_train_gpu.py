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


def test_autograd_node_contract():
    """The native training node behaves like an autograd node should: an in-place parameter update between forward and backward
    is detected (saved-tensor version check), a second backward on the same graph and a mixture that requires grad fail loudly."""
    cfg = O.OracleConfig(n_basis=32, kernel_size=16, sep_hidden_channels=64, sep_bottleneck_channels=32, sep_skip_channels=32,
                         sep_num_blocks=1, sep_num_layers=2, causal=False, n_sources=2)
    model = build_model(cfg, O.synth_state_dict(cfg, seed=3)).train()
    mixture, sources = O.synth_batch(2, 2, 2000, seed=8)
    mixture, sources = mixture.cuda(), sources.cuda()
    crit = PIT1d(NegSISDR(), 2)
    loss, _ = crit(model(mixture), sources)
    loss.backward(retain_graph=True)
    with pytest.raises(RuntimeError):
        loss.backward()
    loss, _ = crit(model(mixture), sources)
    with torch.no_grad():
        next(model.parameters()).add_(1.0)
    with pytest.raises(RuntimeError):          # "one of the variables needed for gradient computation has been modified"
        loss.backward()
    with pytest.raises(NotImplementedError):
        model(mixture.clone().requires_grad_(True))


PAPER = dict(n_basis=512, kernel_size=16, sep_hidden_channels=512, sep_bottleneck_channels=128, sep_skip_channels=128,
             sep_num_blocks=3, sep_num_layers=8)


def _oracle_grads64(cfg, sd, mixture, sources):
    sd = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    out, _ = O.conv_tasnet_fwd(mixture.double(), sd, cfg)
    loss, perm = O.pit_neg_sisdr(out, sources.double(), batch_mean=True)
    loss.backward()
    return {k: v.grad for k, v in sd.items()}


# Paper-size gradient criterion.  Measured on the reference itself (tests/golden/make_golden.py grad_case): at N = H = 512 its fp32
# backward is 3e-4 (median) ... 1e-1 (single PReLU slopes) away from its own fp64 backward, relative to each tensor's largest entry; over
# the whole 4.98 M-entry gradient the relative L2 distance fp32 <-> fp64 is 5.9e-4.  The parameter gradients are sums over ~16 000 frames
# x 512 channels with ~1e4-fold cancellation, which amplifies every rounding error of the data gradients by that factor, and the
# per-tensor noise is heavy-tailed (7e-8 ... 1e-1).  A second fp32 implementation therefore cannot agree with the reference's fp32
# numbers to 2e-4 at this size; what is asserted is the distance to the fp64 answer:
#   * whole gradient: ||g - g64||_2 / ||g64||_2 <= L2MAX[mode]   (the quantity SGD / Adam see).  Measured on B200: 1.5e-3 for the
#     tcgen05 modes = 2.5x the reference's own fp32 (their 3-pass tf32 split carries 22-bit operands: products are 2^-21 relative
#     instead of 2^-24), 1.2e-3 against the reference's fp64 fixture;
#   * every tensor: max |g - g64| <= PER[mode] * scale(k), scale(k) = the largest |g64| entry among the tensors of the same role (all
#     48 PReLU-slope gradients, all 24 depthwise weights, ...: a scalar that happens to be ~0 is judged against its peers) -- a
#     structural check (a missing term or a wrong tile shows up at O(0.1 .. 1)); measured worst 1.2e-2.
# The toy shapes above, where cancellation is mild, keep the tight per-tensor 2e-4 bound against the fp32 oracle.
GRAD_CRIT = {"fp32": (3e-2, 3e-3), "tf32x3": (3e-2, 5e-3), "f16x3": (3e-2, 5e-3), None: (3e-2, 5e-3)}


def _role(k):
    return ".".join(k.split(".")[-2:])


def _check_grads_vs_fp64(named_grads, g64max, g64, noise32, mode):
    """named_grads: [(key, tensor or strided sample)], g64: same shapes, g64max[key]: largest |g64| entry of the full tensor."""
    per, l2max = GRAD_CRIT[mode]
    group = {}
    for k, m in g64max.items():
        group[_role(k)] = max(group.get(_role(k), 0.0), m)
    worst, closer, num, den = (0.0, None), 0, 0.0, 0.0
    for k, g in named_grads:
        r = g64[k]
        scale = group[_role(k)]
        err = float((g.double() - r.double()).abs().max())
        assert err <= per * scale + GRAD_ATOL, "{}: |g-g64| {:.3e} = {:.2e} of its role scale (> {:.0e}); reference fp32 noise {:.3e}".format(
            k, err, err / (scale + 1e-30), per, noise32[k])
        closer += err <= noise32[k]
        num += float(((g.double() - r.double()) ** 2).sum())
        den += float((r.double() ** 2).sum())
        if err / (scale + 1e-30) > worst[0]:
            worst = (err / (scale + 1e-30), k)
    l2 = (num / (den + 1e-300)) ** 0.5
    assert l2 <= l2max, "relative L2 error of the whole gradient {:.3e} > {:.1e}".format(l2, l2max)
    return worst, closer, l2


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("S", [2, 3])
def test_paper_size_gradients_vs_oracle_autograd(mode, S):
    """BASELINE hyper-parameters (N=512 L=16 B=128 H=512 Sc=128 X=8 R=3; cfg2 = 2 speakers, cfg3 = 3 speakers), batch 2,
    T = 8000: the tcgen05 weight-gradient kernel runs its 4 M-tiles / K = 512 shapes and the split-K red.add path.  All 343
    gradient tensors against torch autograd over the oracle IN FP64 (egs/wsj0-mix/common/src/driver.py:146-150), tolerance
    anchored on the fp32 oracle's own distance to fp64 (see _check_grads_vs_fp64)."""
    cfg = O.OracleConfig(causal=False, n_sources=S, **PAPER)
    sd = O.synth_state_dict(cfg, seed=113)
    mixture, sources = O.synth_batch(2, S, 8000, seed=113)
    ref_out, ref_loss, ref_perm, g32 = _oracle_grads(cfg, sd, mixture, sources)
    g64 = _oracle_grads64(cfg, sd, mixture, sources)
    noise32 = {k: float((g32[k].double() - g64[k]).abs().max()) for k in g64}
    model = build_model(cfg, sd, math=mode).train()
    out = model(mixture.cuda())
    torch.testing.assert_close(out.detach().cpu(), ref_out, rtol=1e-4, atol=2e-5)
    loss, perm = PIT1d(NegSISDR(), S)(out, sources.cuda())
    assert torch.equal(perm.cpu(), ref_perm)
    torch.testing.assert_close(loss.detach().cpu(), ref_loss, rtol=0, atol=1e-4)
    loss.backward()
    assert len(g64) == 343
    g64max = {k: float(v.abs().max()) for k, v in g64.items()}
    n32 = sum(float(((g32[k].double() - g64[k]) ** 2).sum()) for k in g64)
    d64 = sum(float((g64[k] ** 2).sum()) for k in g64)
    worst, closer, l2 = _check_grads_vs_fp64([(k, p.grad.detach().cpu()) for k, p in model.named_parameters()], g64max, g64, noise32, mode)
    print("paper-size gradients vs fp64 [{} S={}]: worst per-tensor error / role scale {:.2e} ({}); relative L2 of the whole gradient {:.2e} "
          "(CPU fp32 oracle: {:.2e}); tensors at least as close to fp64 as the CPU fp32 oracle: {} / 343".format(
              mode, S, worst[0], worst[1], l2, (n32 / d64) ** 0.5, closer))


def test_paper_size_gradients_vs_reference_golden(golden_dir):
    """Same shape against the fixture minted from the UNMODIFIED reference's loss.backward() in fp64 (tests/golden/make_golden.py
    grad_case): loss, permutation, and every 97th element of each of the 343 gradient tensors; tolerance anchored on the
    reference's own fp32-vs-fp64 distance stored in the fixture."""
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
    st = rec["stride"]
    named = [(k, p.grad.detach().cpu().flatten()[::st]) for k, p in model.named_parameters()]
    g64 = {k: g["sample64"] for k, g in rec["grads"].items()}
    g64max = {k: g["absmax64"] for k, g in rec["grads"].items()}
    noise32 = {k: g["fp32_vs_fp64_maxabs"] for k, g in rec["grads"].items()}
    assert len(named) == 343
    worst, closer, l2 = _check_grads_vs_fp64(named, g64max, g64, noise32, None)
    print("paper-size gradients vs the reference's fp64 backward: worst {:.2e} ({}), relative L2 over the sampled entries {:.2e}".format(worst[0], worst[1], l2))


@pytest.mark.parametrize("max_norm,wd", [(5.0, 0.0), (0.05, 0.0), (None, 1e-2)])
def test_native_clip_adam_matches_torch(max_norm, wd):
    """ctn_clip_adam_step (3 launches over the flat gradient bucket) against torch.nn.utils.clip_grad_norm_ + torch.optim.Adam
    (egs/wsj0-mix/common/src/driver.py:152-155) fed with the SAME gradients, 4 steps: parameters within 1e-6, reported total norm
    equal.  max_norm = 0.05 makes the clip active every step."""
    import copy
    from ctn_b200.optim import FlatClipAdam
    cfg = O.OracleConfig(n_basis=32, kernel_size=16, sep_hidden_channels=64, sep_bottleneck_channels=32, sep_skip_channels=32,
                         sep_num_blocks=2, sep_num_layers=3, causal=False, n_sources=2)
    ours = build_model(cfg, O.synth_state_dict(cfg, seed=3)).train()
    ref = copy.deepcopy(ours)
    mixture, sources = O.synth_batch(4, 2, 4000, seed=8)
    mixture, sources = mixture.cuda(), sources.cuda()
    crit = PIT1d(NegSISDR(), 2)
    opt = FlatClipAdam(ours, lr=1e-3, weight_decay=wd, max_norm=max_norm)
    topt = torch.optim.Adam(ref.parameters(), lr=1e-3, weight_decay=wd)
    for it in range(4):
        opt.zero_grad()
        loss, _ = crit(ours(mixture), sources)
        loss.backward()
        # hand the reference optimizer the very same gradients (the native backward sums with atomics: two runs differ by ~1e-7)
        for p, q in zip(ours.parameters(), ref.parameters()):
            q.grad = p.grad.detach().clone()
        tn = opt.step()
        if max_norm is not None:
            tn_ref = torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm)
            torch.testing.assert_close(tn.reshape(()), tn_ref.reshape(()), rtol=1e-5, atol=1e-7)
        topt.step()
        if it == 1:
            opt.set_lr(5e-4)                      # LR halving (adhoc_driver.py:25-39) without rebuilding anything
            for gr in topt.param_groups:
                gr["lr"] = 5e-4
    for (k, p), q in zip(ours.named_parameters(), ref.parameters()):
        torch.testing.assert_close(p.detach(), q.detach(), rtol=1e-5, atol=1e-6, msg=lambda m, k=k: k + ": " + m)
    assert int(opt.step_count[0]) == 4
