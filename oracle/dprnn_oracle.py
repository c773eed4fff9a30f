"""CPU oracle for the DPRNN-TasNet path, BASELINE cfg4 (TEST INFRASTRUCTURE ONLY -- same rules as convtasnet_oracle.py:
only tests/, smoke() and bench.py's CPU legs may import it; the product never does).

Functional restatement of src/models/dprnn_tasnet.py, src/models/dprnn.py and src/models/transform.py of the reference in
plain PyTorch CPU ops; every function cites the lines it follows.  The LSTM recurrence itself is third-party arithmetic in
the reference too (``nn.LSTM``, torch -- pinned torch==1.10.0 in egs/tutorials/requirements.txt:5): it is called here through
the same ATen op (``torch.lstm`` via nn.LSTM functional form) with the reference's parameter names.

Parity status: PINNED against fixtures minted from the unmodified reference (tests/golden/make_golden.py: dprnn_* cases;
tests/test_oracle_golden.py checks this file against them).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Dict, Optional

import torch
import torch.nn.functional as F

import convtasnet_oracle as O

EPS = 1e-12


@dataclass
class DPRNNConfig:
    """Constructor arguments of DPRNNTasNet (src/models/dprnn_tasnet.py:33-47)."""
    n_basis: int = 64
    kernel_size: int = 2
    stride: Optional[int] = None
    sep_hidden_channels: int = 128
    sep_bottleneck_channels: int = 64
    sep_chunk_size: int = 250
    sep_hop_size: int = 125
    sep_num_blocks: int = 6
    sep_norm: bool = True
    mask_nonlinear: str = "sigmoid"
    causal: bool = False
    rnn_type: str = "lstm"
    n_sources: int = 2
    eps: float = EPS
    enc_nonlinear: Optional[str] = None

    def __post_init__(self):
        if self.stride is None:
            self.stride = self.kernel_size // 2  # dprnn_tasnet.py:50-51
        assert self.kernel_size % self.stride == 0

    def to_dict(self):
        return asdict(self)


# ---- transform.py ------------------------------------------------------------------------------------------------------
def segment1d(x: torch.Tensor, chunk_size: int, hop_size: int) -> torch.Tensor:
    """Segment1d.forward, src/models/transform.py:15-29: (B, F, T) -> (B, F, S, K), S = (T - K)//P + 1, chunk s = frames [sP, sP+K)."""
    B, Fc, T = x.shape
    S = (T - chunk_size) // hop_size + 1
    idx = (torch.arange(S).unsqueeze(1) * hop_size + torch.arange(chunk_size).unsqueeze(0)).reshape(-1)
    return x[:, :, idx].reshape(B, Fc, S, chunk_size)


def overlap_add1d(x: torch.Tensor, chunk_size: int, hop_size: int) -> torch.Tensor:
    """OverlapAdd1d.forward, src/models/transform.py:46-62: (B, F, S, K) -> (B, F, (S-1)P + K), overlapping chunks are summed
    (F.fold accumulates in increasing chunk order)."""
    B, Fc, S, K = x.shape
    T = (S - 1) * hop_size + K
    out = torch.zeros(B, Fc, T, dtype=x.dtype)
    for s in range(S):
        out[:, :, s * hop_size:s * hop_size + K] += x[:, :, s]
    return out


# ---- dprnn.py ----------------------------------------------------------------------------------------------------------
def _bilstm(x: torch.Tensor, sd, prefix: str) -> torch.Tensor:
    """nn.LSTM(input_size, hidden, batch_first=True, bidirectional=True) (dprnn.py:60, 114-120) with the reference's parameter
    names {weight_ih,weight_hh,bias_ih,bias_hh}_l0[_reverse]; zero initial state."""
    flat = [sd[prefix + n] for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0", "weight_ih_l0_reverse",
                                     "weight_hh_l0_reverse", "bias_ih_l0_reverse", "bias_hh_l0_reverse")]
    H = flat[1].shape[1]
    h0 = torch.zeros(2, x.shape[0], H, dtype=x.dtype)
    y, _, _ = torch.lstm(x, (h0, h0.clone()), flat, True, 1, 0.0, False, True, True)
    return y


def intra_chunk(x: torch.Tensor, sd, prefix: str, eps: float) -> torch.Tensor:
    """IntraChunkRNN.forward, src/models/dprnn.py:70-94; x (B, F, S, K)."""
    B, Fc, S, K = x.shape
    y = x.permute(0, 2, 3, 1).contiguous().view(B * S, K, Fc)          # :83-84
    y = _bilstm(y, sd, prefix + "rnn.")                               # :85
    y = F.linear(y, sd[prefix + "fc.weight"], sd[prefix + "fc.bias"])  # :86
    y = y.view(B, S * K, Fc).permute(0, 2, 1).contiguous()            # :87-88
    y = O.gln(y, sd[prefix + "norm1d.norm.weight"], sd[prefix + "norm1d.norm.bias"], eps)  # :89-90
    return y.view(B, Fc, S, K) + x                                    # :91-92


def inter_chunk(x: torch.Tensor, sd, prefix: str, eps: float) -> torch.Tensor:
    """InterChunkRNN.forward (non-causal), src/models/dprnn.py:122-148; x (B, F, S, K)."""
    B, Fc, S, K = x.shape
    y = x.permute(0, 3, 2, 1).contiguous().view(B * K, S, Fc)          # :136-137
    y = _bilstm(y, sd, prefix + "rnn.")                               # :138
    y = F.linear(y, sd[prefix + "fc.weight"], sd[prefix + "fc.bias"])  # :139
    y = y.view(B, K * S, Fc).permute(0, 2, 1).contiguous()            # :140-141
    y = O.gln(y, sd[prefix + "norm1d.norm.weight"], sd[prefix + "norm1d.norm.bias"], eps)  # :142-143
    y = y.view(B, Fc, K, S).permute(0, 1, 3, 2).contiguous()          # :144-145
    return y + x                                                      # :147


def dprnn_fwd(x: torch.Tensor, sd, prefix: str, num_blocks: int, eps: float) -> torch.Tensor:
    """DPRNN.forward, src/models/dprnn.py:21-30, 43-54."""
    for i in range(num_blocks):
        x = intra_chunk(x, sd, f"{prefix}net.{i}.intra_chunk_block.", eps)
        x = inter_chunk(x, sd, f"{prefix}net.{i}.inter_chunk_block.", eps)
    return x


# ---- dprnn_tasnet.py ---------------------------------------------------------------------------------------------------
def separator_fwd(w: torch.Tensor, sd, cfg: DPRNNConfig, prefix: str = "separator.") -> torch.Tensor:
    """Separator.forward, src/models/dprnn_tasnet.py:324-353."""
    B, N, n_frames = w.shape
    K, P = cfg.sep_chunk_size, cfg.sep_hop_size
    padding = (P - (n_frames - K) % P) % P                             # :339
    pl = padding // 2
    pr = padding - pl
    x = O.gln(w, sd[prefix + "norm1d.norm.weight"], sd[prefix + "norm1d.norm.bias"], cfg.eps)         # :343
    x = F.conv1d(x, sd[prefix + "bottleneck_conv1d.weight"], sd[prefix + "bottleneck_conv1d.bias"])   # :344
    x = F.pad(x, (pl, pr))                                                                           # :345
    x = segment1d(x, K, P)                                                                           # :346
    x = dprnn_fwd(x, sd, prefix + "dprnn.", cfg.sep_num_blocks, cfg.eps)                             # :347
    x = overlap_add1d(x, K, P)                                                                       # :348
    x = F.pad(x, (-pl, -pr))                                                                         # :349
    x = O.prelu(x, sd[prefix + "prelu.weight"])                                                      # :350
    x = F.conv1d(x, sd[prefix + "mask_conv1d.weight"], sd[prefix + "mask_conv1d.bias"])              # :351
    x = torch.sigmoid(x)                                                                             # :352
    return x.view(B, cfg.n_sources, N, n_frames)


def dprnn_tasnet_fwd(x: torch.Tensor, sd: Dict[str, torch.Tensor], cfg: DPRNNConfig):
    """DPRNNTasNet.extract_latent, src/models/dprnn_tasnet.py:106-156 (3-D input, real-valued bases)."""
    B, C_in, T = x.shape
    assert C_in == 1
    K, S = cfg.kernel_size, cfg.stride
    padding = (S - (T - K) % S) % S                                     # :129
    pl = padding // 2
    pr = padding - pl
    x = F.pad(x, (pl, pr))                                              # :133
    w = O.encoder_fwd(x, sd["encoder.conv1d.weight"], S, relu=cfg.enc_nonlinear == "relu")  # :134
    mask = separator_fwd(w, sd, cfg)                                    # :142
    w_hat = w.unsqueeze(1) * mask                                       # :143-144
    latent = w_hat
    x_hat = O.decoder_fwd(w_hat.view(B * cfg.n_sources, cfg.n_basis, -1), sd["decoder.conv_transpose1d.weight"], S)  # :147-148
    x_hat = x_hat.view(B, cfg.n_sources, -1)
    out = F.pad(x_hat, (-pl, -pr))                                      # :153
    return out, latent


# ---- deterministic synthetic parameters ------------------------------------------------------------------------------------
def state_dict_spec(cfg: DPRNNConfig):
    """(key, shape) in the reference's state_dict order (verified against the reference by tests/golden/make_golden.py)."""
    N, L, Fc, H, S = cfg.n_basis, cfg.kernel_size, cfg.sep_bottleneck_channels, cfg.sep_hidden_channels, cfg.n_sources
    spec = [("encoder.conv1d.weight", (N, 1, L)),
            ("separator.norm1d.norm.weight", (N,)), ("separator.norm1d.norm.bias", (N,)),
            ("separator.bottleneck_conv1d.weight", (Fc, N, 1)), ("separator.bottleneck_conv1d.bias", (Fc,))]
    for i in range(cfg.sep_num_blocks):
        for path in ("intra_chunk_block", "inter_chunk_block"):
            p = f"separator.dprnn.net.{i}.{path}."
            for suf in ("", "_reverse"):
                spec += [(p + "rnn.weight_ih_l0" + suf, (4 * H, Fc)), (p + "rnn.weight_hh_l0" + suf, (4 * H, H)),
                         (p + "rnn.bias_ih_l0" + suf, (4 * H,)), (p + "rnn.bias_hh_l0" + suf, (4 * H,))]
            spec += [(p + "fc.weight", (Fc, 2 * H)), (p + "fc.bias", (Fc,)),
                     (p + "norm1d.norm.weight", (Fc,)), (p + "norm1d.norm.bias", (Fc,))]
    spec += [("separator.prelu.weight", (1,)), ("separator.mask_conv1d.weight", (S * N, Fc, 1)), ("separator.mask_conv1d.bias", (S * N,)),
             ("decoder.conv_transpose1d.weight", (N, 1, L))]
    return spec


def synth_state_dict(cfg: DPRNNConfig, seed: int = 111, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    sd = {}
    H = cfg.sep_hidden_channels
    for i, (key, shape) in enumerate(state_dict_spec(cfg)):
        g = torch.Generator().manual_seed(seed * 100003 + i)
        leaf, parent = key.split(".")[-1], key.split(".")[-2]
        if parent == "norm":
            t = 1.0 + 0.2 * (torch.rand(shape, generator=g) - 0.5) if leaf == "weight" else 0.1 * (torch.rand(shape, generator=g) - 0.5)
        elif parent == "prelu":
            t = 0.25 + 0.1 * (torch.rand(shape, generator=g) - 0.5)
        elif parent == "rnn":
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(H)      # torch's LSTM default init scale
        elif parent == "fc":
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(2 * H)
        elif leaf == "weight":
            fan_in = shape[1] * shape[2] if "conv_transpose" not in key else shape[2]
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
        else:
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        sd[key] = t.to(dtype)
    return sd
