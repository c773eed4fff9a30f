"""GPU parity of the block-level modules (``-m gpu``): ResidualBlock1d / TimeDilatedConvBlock1d forward through
ctn_tcn_blocks_fwd (src/models/tdcn.py:65-75, 107-147) against the oracle's residual_block, and the plain
modules.conv.DepthwiseSeparableConv1d (src/modules/conv.py:13-29) against the ATen ops the reference module dispatches to."""
import pytest
import torch
import torch.nn.functional as F

import convtasnet_oracle as O
from ctn_b200 import _native as N
from ctn_b200.models.tdcn import ResidualBlock1d, TimeDilatedConvBlock1d
from ctn_b200.modules.conv import DepthwiseSeparableConv1d

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-4, 2e-5
MODES = ["fp32"] + (["tf32x3", "f16x3"] if N.ctn_has_tcgen05() else [])


def _block_sd(block, prefix, seed):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in block.state_dict().items():
        if k.endswith("norm.weight"):
            t = 1.0 + 0.3 * torch.randn(v.shape, generator=g)
        elif k.endswith("norm.bias"):
            t = 0.2 * torch.randn(v.shape, generator=g)
        elif "nonlinear1d" in k:
            t = 0.25 + 0.1 * torch.rand(v.shape, generator=g)
        else:
            fan = v[0].numel() if v.dim() > 1 else 16
            t = (torch.rand(v.shape, generator=g) * 2 - 1) / fan ** 0.5
        sd[k] = t
    block.load_state_dict(sd)
    return {prefix + k: v for k, v in sd.items()}


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("dilation,dual_head", [(1, True), (2, True), (4, True), (3, True), (16, False), (6, True), (128, True)])
def test_residual_block_forward(mode, dilation, dual_head):
    """one block, explicit dilation (3 and 6 are outside the fused depthwise producer: stand-alone depthwise stage)"""
    blk = ResidualBlock1d(24, hidden_channels=48, skip_channels=20, kernel_size=3, stride=1, dilation=dilation, separable=True, causal=False,
                          nonlinear="prelu", norm=True, dual_head=dual_head)
    sd = _block_sd(blk, "b.", seed=dilation)
    blk = blk.cuda().eval()
    x = torch.randn(2, 24, 517, generator=torch.Generator().manual_seed(3))
    import ctn_b200.models.tdcn as T
    old = T.DEFAULT_MATH
    T.DEFAULT_MATH = mode
    try:
        with torch.no_grad():
            out, skip = blk(x.cuda())
    finally:
        T.DEFAULT_MATH = old
    ref_out, ref_skip = O.residual_block(x, sd, "b.", kernel_size=3, dilation=dilation, causal=False, dual_head=dual_head, nonlinear=True,
                                         norm=True, eps=1e-12)
    torch.testing.assert_close(skip.cpu(), ref_skip, rtol=RTOL, atol=ATOL)
    if dual_head:
        torch.testing.assert_close(out.cpu(), ref_out, rtol=RTOL, atol=ATOL)
    else:
        assert out is None


@pytest.mark.parametrize("dual_head", [True, False])
def test_conv_block_forward(dual_head):
    """TimeDilatedConvBlock1d.forward: X layers, dilation 2^l, returns (x after the last layer | None, sum of the skips)"""
    blk = TimeDilatedConvBlock1d(16, hidden_channels=32, skip_channels=16, kernel_size=3, num_layers=5, dilated=True, separable=True,
                                 causal=False, nonlinear="prelu", norm=True, dual_head=dual_head)
    sd = _block_sd(blk, "", seed=11)
    blk = blk.cuda().eval()
    x = torch.randn(3, 16, 300, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        out, skip = blk(x.cuda())
    xr, skip_ref = x, 0
    for l in range(5):
        dh = dual_head or l < 4
        o, s = O.residual_block(xr, sd, f"net.{l}.", kernel_size=3, dilation=2 ** l, causal=False, dual_head=dh, nonlinear=True, norm=True, eps=1e-12)
        skip_ref = skip_ref + s
        if o is not None:
            xr = o
    torch.testing.assert_close(skip.cpu(), skip_ref, rtol=RTOL, atol=ATOL)
    if dual_head:
        torch.testing.assert_close(out.cpu(), xr, rtol=RTOL, atol=ATOL)
    else:
        assert out is None


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("C,M,K,stride,padding,dilation,bias,T", [(16, 24, 3, None, 0, 1, True, 301), (8, 8, 5, 2, 4, 2, True, 200), (33, 17, 4, 1, 0, 3, False, 77),
                                                                   (64, 128, 3, 1, 8, 8, True, 1000)])
def test_plain_depthwise_separable_conv1d(mode, C, M, K, stride, padding, dilation, bias, T):
    m = DepthwiseSeparableConv1d(C, M, K, stride=stride, padding=padding, dilation=dilation, bias=bias)
    m.math = mode
    x = torch.randn(2, C, T, generator=torch.Generator().manual_seed(K))
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.cuda().eval()
    with torch.no_grad():
        y = m(x.cuda())
    s = K if stride is None else stride
    u = F.conv1d(x, sd["depthwise_conv1d.weight"], sd.get("depthwise_conv1d.bias"), stride=s, padding=padding, dilation=dilation, groups=C)  # conv.py:24
    ref = F.conv1d(u, sd["pointwise_conv1d.weight"], sd.get("pointwise_conv1d.bias"))                                                      # conv.py:25
    torch.testing.assert_close(y.cpu(), ref, rtol=RTOL, atol=ATOL)
    assert list(m.state_dict().keys()) == list(sd.keys())
