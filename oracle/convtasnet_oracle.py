"""CPU oracle for the Conv-TasNet separation path (TEST INFRASTRUCTURE ONLY).

This file is a *restatement* of the reference algorithm in plain functional
PyTorch fp32/fp64 on the CPU.  It is the checker for the CUDA path: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``dnn-based_source_separation_b200/`` imports it, and the product path has no
CPU fallback.

Parity status: PINNED.  The reference repository holds no golden vectors for
this path, so the oracle is pinned against outputs of the
reference itself: ``tests/golden/make_golden.py`` imports the unmodified
reference from /root/reference/src, runs it on seeded weights/inputs and
commits the results as fixtures; ``tests/test_oracle_golden.py`` checks this
file against them.

Every function cites the reference lines (relative to /root/reference/) whose
arithmetic it restates.  Parameters are taken from a flat ``state_dict`` with
the reference's key names (SURVEY.md section 8a footer), so a reference checkpoint
can be fed in unchanged.
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass, asdict
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

EPS = 1e-12


@dataclass
class OracleConfig:
    """Constructor arguments of ConvTasNet (src/models/conv_tasnet.py:57-66)."""
    n_basis: int = 512
    kernel_size: int = 16
    stride: Optional[int] = None
    sep_hidden_channels: int = 512
    sep_bottleneck_channels: int = 128
    sep_skip_channels: int = 128
    sep_kernel_size: int = 3
    sep_num_blocks: int = 3
    sep_num_layers: int = 8
    dilated: bool = True
    separable: bool = True
    sep_nonlinear: Optional[str] = "prelu"
    sep_norm: bool = True
    mask_nonlinear: str = "sigmoid"
    causal: bool = False
    n_sources: int = 2
    eps: float = EPS
    enc_nonlinear: Optional[str] = None
    in_channels: int = 1  # kwargs['in_channels'], conv_tasnet.py:75 (n_mics of the 4-D input form)

    def __post_init__(self):
        if self.stride is None:
            self.stride = self.kernel_size // 2  # conv_tasnet.py:69-70
        assert self.kernel_size % self.stride == 0  # conv_tasnet.py:72

    def to_dict(self):
        return asdict(self)


# --------------------------------------------------------------------------
# filterbank
# --------------------------------------------------------------------------

def encoder_fwd(x: torch.Tensor, weight: torch.Tensor, stride: int, relu: bool = False) -> torch.Tensor:
    """Encoder.forward, src/models/filterbank.py:222-229 (Conv1d, bias=False, optional ReLU)."""
    w = F.conv1d(x, weight, bias=None, stride=stride)
    return torch.relu(w) if relu else w


def decoder_fwd(w_hat: torch.Tensor, weight: torch.Tensor, stride: int) -> torch.Tensor:
    """Decoder.forward, src/models/filterbank.py:245-247 (ConvTranspose1d, bias=False)."""
    return F.conv_transpose1d(w_hat, weight, bias=None, stride=stride)


# --------------------------------------------------------------------------
# norms
# --------------------------------------------------------------------------

def gln(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = EPS) -> torch.Tensor:
    """GlobalLayerNorm.forward, src/modules/norm.py:18,32: GroupNorm(1, C, eps).

    Per sample: mean / biased variance over all (C, T); y = (x-mean)/sqrt(var+eps)*gamma_c+beta_c.
    """
    return F.group_norm(x, 1, gamma, beta, eps)  # the ATen op nn.GroupNorm dispatches to


def cln(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = EPS) -> torch.Tensor:
    """CumulativeLayerNorm1d.forward, src/modules/norm.py:78-90.

    Statistics over channels and all frames <= t; note eps sits OUTSIDE the sqrt (norm.py:90).
    """
    B, C, T = x.shape
    step_sum = x.sum(dim=1)
    step_sq = (x ** 2).sum(dim=1)
    cum_sum = torch.cumsum(step_sum, dim=1)
    cum_sq = torch.cumsum(step_sq, dim=1)
    cum_num = torch.arange(C, C * (T + 1), C, dtype=x.dtype, device=x.device)
    cum_mean = cum_sum / cum_num
    cum_var = cum_sq / cum_num - cum_mean ** 2
    cum_mean, cum_var = cum_mean.unsqueeze(1), cum_var.unsqueeze(1)
    return (x - cum_mean) / (torch.sqrt(cum_var) + eps) * gamma.view(1, C, 1) + beta.view(1, C, 1)


def _norm(x, sd, prefix, causal, eps):
    """choose_layer_norm, src/utils/tasnet.py:14-21: 'cLN' if causal else 'gLN'."""
    if causal:
        return cln(x, sd[prefix + "gamma"], sd[prefix + "beta"], eps)
    return gln(x, sd[prefix + "norm.weight"], sd[prefix + "norm.bias"], eps)


def prelu(x: torch.Tensor, a: torch.Tensor) -> torch.Tensor:
    """nn.PReLU() with a single shared slope (src/models/tdcn.py:90,161)."""
    return F.prelu(x, a.view(-1))


# --------------------------------------------------------------------------
# TCN
# --------------------------------------------------------------------------

def residual_block(x, sd, prefix, *, kernel_size, dilation, causal, dual_head, nonlinear, norm, eps):
    """ResidualBlock1d.forward + DepthwiseSeparableConv1d.forward (separable branch).

    src/models/tdcn.py:107-147 and :177-196.  Returns (output | None, skip).
    """
    T = x.shape[-1]
    h = F.conv1d(x, sd[prefix + "bottleneck_conv1d.weight"], sd[prefix + "bottleneck_conv1d.bias"])  # :116
    if nonlinear:
        h = prelu(h, sd[prefix + "nonlinear1d.weight"])  # :118-119
    if norm:
        h = _norm(h, sd, prefix + "norm1d.", causal, eps)  # :120-121
    stride = 1
    padding = (T - 1) * stride - T + (kernel_size - 1) * dilation + 1  # :123
    if causal:
        pl, pr = padding, 0  # :125-127
    else:
        pl = padding // 2  # :129
        pr = padding - pl  # :130
    h = F.pad(h, (pl, pr))  # :132
    p2 = prefix + "separable_conv1d."
    C = h.shape[1]
    u = F.conv1d(h, sd[p2 + "depthwise_conv1d.weight"], sd[p2 + "depthwise_conv1d.bias"],
                 stride=stride, dilation=dilation, groups=C)  # :181
    if nonlinear:
        u = prelu(u, sd[p2 + "nonlinear1d.weight"])  # :183-184
    if norm:
        u = _norm(u, sd, p2 + "norm1d.", causal, eps)  # :186-187
    out = None
    if dual_head:
        out = F.conv1d(u, sd[p2 + "output_pointwise_conv1d.weight"], sd[p2 + "output_pointwise_conv1d.bias"])  # :190
        out = out + x  # :144-145
    skip = F.conv1d(u, sd[p2 + "skip_pointwise_conv1d.weight"], sd[p2 + "skip_pointwise_conv1d.bias"])  # :194
    return out, skip


def tdcn_fwd(x, sd, prefix, *, kernel_size, num_blocks, num_layers, dilated, causal, nonlinear, norm, eps):
    """TimeDilatedConvNet.forward / TimeDilatedConvBlock1d.forward, src/models/tdcn.py:29-41, 65-75.

    (== TemporalConvNet, src/models/tcn.py:37-49.)  Returns the skip sum only.
    """
    if not dilated:
        raise NotImplementedError("oracle restates the dilated=True path only")
    skip_total = 0  # :33 (python int)
    for r in range(num_blocks):
        skip_block = 0  # :69
        for l in range(num_layers):
            dual_head = not (r == num_blocks - 1 and l == num_layers - 1)  # :22-25, :58-61
            out, skip = residual_block(
                x, sd, f"{prefix}net.{r}.net.{l}.", kernel_size=kernel_size, dilation=2 ** l, causal=causal,
                dual_head=dual_head, nonlinear=nonlinear, norm=norm, eps=eps)
            if out is not None:
                x = out
            skip_block = skip_block + skip  # :73
        skip_total = skip_total + skip_block  # :37
    return skip_total


def separator_fwd(w, sd, cfg: OracleConfig, prefix="separator."):
    """Separator.forward, src/models/conv_tasnet.py:359-378."""
    B, N, Tf = w.shape
    x = _norm(w, sd, prefix + "norm1d.", cfg.causal, cfg.eps)  # :370
    x = F.conv1d(x, sd[prefix + "bottleneck_conv1d.weight"], sd[prefix + "bottleneck_conv1d.bias"])  # :371
    x = tdcn_fwd(x, sd, prefix + "tdcn.", kernel_size=cfg.sep_kernel_size, num_blocks=cfg.sep_num_blocks,
                 num_layers=cfg.sep_num_layers, dilated=cfg.dilated, causal=cfg.causal,
                 nonlinear=cfg.sep_nonlinear is not None, norm=cfg.sep_norm, eps=EPS)  # :372 (tdcn built without eps -> default)
    x = prelu(x, sd[prefix + "prelu.weight"])  # :373
    x = F.conv1d(x, sd[prefix + "mask_conv1d.weight"], sd[prefix + "mask_conv1d.bias"])  # :374
    if cfg.mask_nonlinear == "sigmoid":
        x = torch.sigmoid(x)  # :375
    elif cfg.mask_nonlinear == "softmax":
        x = torch.softmax(x, dim=1)  # :345-357 quirk: over all S*N channels
    else:
        raise ValueError("Cannot support {}".format(cfg.mask_nonlinear))
    return x.view(B, cfg.n_sources, N, Tf)  # :376


def conv_tasnet_fwd(x: torch.Tensor, sd: Dict[str, torch.Tensor], cfg: OracleConfig) -> Tuple[torch.Tensor, torch.Tensor]:
    """ConvTasNet.extract_latent, src/models/conv_tasnet.py:121-171.  Returns (output, latent)."""
    n_dims = x.dim()
    if n_dims == 4:  # (B, 1, n_mics, T) -> (B, n_mics, T), :138-141
        assert x.shape[1] == 1
        x = x.reshape(x.shape[0], x.shape[2], x.shape[3])
    elif n_dims != 3:
        raise ValueError("Not support {} dimension input".format(n_dims))
    B, C_in, T = x.shape
    assert C_in == (1 if n_dims == 3 else cfg.in_channels)
    K, S = cfg.kernel_size, cfg.stride
    padding = (S - (T - K) % S) % S  # :145
    pl = padding // 2
    pr = padding - pl
    xp = F.pad(x, (pl, pr))  # :149
    w = encoder_fwd(xp, sd["encoder.conv1d.weight"], S, relu=(cfg.enc_nonlinear == "relu"))  # :150
    mask = separator_fwd(w, sd, cfg)  # :158
    w_hat = w.unsqueeze(1) * mask  # :159-160
    latent = w_hat
    x_hat = decoder_fwd(w_hat.reshape(B * cfg.n_sources, cfg.n_basis, -1), sd["decoder.conv_transpose1d.weight"], S)  # :163-164
    x_hat = x_hat.view(B, cfg.n_sources, -1) if n_dims == 3 else x_hat.view(B, cfg.n_sources, C_in, -1)  # :165-168
    out = F.pad(x_hat, (-pl, -pr))  # :169
    return out, latent


# --------------------------------------------------------------------------
# criterion
# --------------------------------------------------------------------------

def sisdr(input: torch.Tensor, target: torch.Tensor, eps: float = EPS) -> torch.Tensor:
    """sisdr, src/criterion/sdr.py:122-139."""
    d = input.dim() - 1
    alpha = torch.sum(input * target, dim=d, keepdim=True) / (torch.sum(target ** 2, dim=d, keepdim=True) + eps)
    loss = (torch.sum((alpha * target) ** 2, dim=d) + eps) / (torch.sum((alpha * target - input) ** 2, dim=d) + eps)
    return 10 * torch.log10(loss)


def sdr(input: torch.Tensor, target: torch.Tensor, eps: float = EPS) -> torch.Tensor:
    """sdr(), src/criterion/sdr.py:6-20: per-row 10 log10((|t|^2 + eps) / (|t - x|^2 + eps)) over the last axis."""
    d = input.dim() - 1
    return 10 * torch.log10((torch.sum(target ** 2, dim=d) + eps) / (torch.sum((target - input) ** 2, dim=d) + eps))


def neg_sisdr(input, target, batch_mean=True, reduction="mean", eps=EPS):
    """NegSISDR.forward, src/criterion/sdr.py:198-227."""
    loss = -sisdr(input, target, eps=eps)
    n_dims = input.dim()
    if reduction:
        if n_dims == 3:
            loss = loss.mean(dim=1) if reduction == "mean" else loss.sum(dim=1)
        elif n_dims == 4:
            loss = loss.mean(dim=(1, 2)) if reduction == "mean" else loss.sum(dim=(1, 2))
    if batch_mean:
        loss = loss.mean(dim=0)
    return loss


def pit_neg_sisdr(input, target, batch_mean=True, reduction="mean", eps=EPS):
    """pit() with criterion NegSISDR, src/criterion/pit.py:9-44 (PIT1d, :71-77).

    The TARGET is permuted (:30); min over permutations, first index on ties (:39).
    Returns (loss, pattern) with pattern int64 (B, S).
    """
    S = input.shape[1]
    patterns = torch.tensor(list(itertools.permutations(range(S))), dtype=torch.long)  # :55-56
    possible = []
    for p in patterns:
        possible.append(neg_sisdr(input, target[:, p], batch_mean=False, reduction=reduction, eps=eps))  # :28-31
    possible = torch.stack(possible, dim=1)
    loss, idx = torch.min(possible, dim=1)  # :39
    if batch_mean:
        loss = loss.mean(dim=0)  # :41-42
    return loss, patterns[idx]


# --------------------------------------------------------------------------
# deterministic synthetic weights / inputs (shared by fixtures, tests, bench)
# --------------------------------------------------------------------------

def state_dict_spec(cfg: OracleConfig):
    """(key, shape) list in the reference's state_dict order (verified by tests/golden/make_golden.py)."""
    N, L = cfg.n_basis, cfg.kernel_size
    Bc, H, Sc, P = cfg.sep_bottleneck_channels, cfg.sep_hidden_channels, cfg.sep_skip_channels, cfg.sep_kernel_size
    spec = [("encoder.conv1d.weight", (N, cfg.in_channels, L))]

    def norm_keys(prefix, C):
        if cfg.causal:
            return [(prefix + "gamma", (1, C, 1)), (prefix + "beta", (1, C, 1))]
        return [(prefix + "norm.weight", (C,)), (prefix + "norm.bias", (C,))]

    spec += norm_keys("separator.norm1d.", N)
    spec += [("separator.bottleneck_conv1d.weight", (Bc, N, 1)), ("separator.bottleneck_conv1d.bias", (Bc,))]
    for r in range(cfg.sep_num_blocks):
        for l in range(cfg.sep_num_layers):
            p = f"separator.tdcn.net.{r}.net.{l}."
            dual = not (r == cfg.sep_num_blocks - 1 and l == cfg.sep_num_layers - 1)
            spec += [(p + "bottleneck_conv1d.weight", (H, Bc, 1)), (p + "bottleneck_conv1d.bias", (H,)),
                     (p + "nonlinear1d.weight", (1,))]
            spec += norm_keys(p + "norm1d.", H)
            q = p + "separable_conv1d."
            spec += [(q + "depthwise_conv1d.weight", (H, 1, P)), (q + "depthwise_conv1d.bias", (H,)),
                     (q + "nonlinear1d.weight", (1,))]
            spec += norm_keys(q + "norm1d.", H)
            if dual:
                spec += [(q + "output_pointwise_conv1d.weight", (Bc, H, 1)), (q + "output_pointwise_conv1d.bias", (Bc,))]
            spec += [(q + "skip_pointwise_conv1d.weight", (Sc, H, 1)), (q + "skip_pointwise_conv1d.bias", (Sc,))]
    spec += [("separator.prelu.weight", (1,)),
             ("separator.mask_conv1d.weight", (cfg.n_sources * N, Sc, 1)), ("separator.mask_conv1d.bias", (cfg.n_sources * N,)),
             ("decoder.conv_transpose1d.weight", (N, cfg.in_channels, L))]
    return spec


def synth_state_dict(cfg: OracleConfig, seed: int = 111, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic, construction-order-independent synthetic weights.

    Conv weights ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (the scale of torch's default
    init), PReLU slopes near 0.25, norm gamma near 1 / beta near 0 but NOT exactly
    (so that affine terms are exercised).  One generator per key, seeded by (seed, key index).
    """
    sd = {}
    for i, (key, shape) in enumerate(state_dict_spec(cfg)):
        g = torch.Generator().manual_seed(seed * 100003 + i)
        leaf = key.split(".")[-1]
        parent = key.split(".")[-2]
        if parent in ("norm",) or leaf in ("gamma", "beta"):
            if leaf in ("weight", "gamma"):
                t = 1.0 + 0.2 * (torch.rand(shape, generator=g) - 0.5)
            else:
                t = 0.1 * (torch.rand(shape, generator=g) - 0.5)
        elif parent in ("nonlinear1d", "prelu"):
            t = 0.25 + 0.1 * (torch.rand(shape, generator=g) - 0.5)
        elif leaf == "weight":
            fan_in = shape[1] * shape[2] if "conv_transpose" not in key else shape[2]
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        else:  # conv bias
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        sd[key] = t.to(dtype)
    return sd


def synth_batch(batch: int, n_sources: int, T: int, seed: int = 111, dtype=torch.float32):
    """SURVEY.md section 8d inputs: sources = 0.1*randn(B,S,T); mixture = sum over sources."""
    g = torch.Generator().manual_seed(seed)
    sources = 0.1 * torch.randn(batch, n_sources, T, generator=g)
    mixture = sources.sum(dim=1, keepdim=True)
    return mixture.to(dtype), sources.to(dtype)
