"""Permutation-invariant training, mirroring src/criterion/pit.py: ``pit`` (:9-44), ``PIT`` (:46-69), ``PIT1d``
(:71-77).  When the criterion is (Neg)SISDR with reduction 'mean'/'sum' on (batch_size, n_sources, T) tensors the whole
thing is ONE fused call (ctn_sisdr_pit_fwd): the S x S pairwise SI-SDR table is computed in two streaming passes and
the S! permutations are scored from it (lexicographic order, first minimum).  Any other criterion goes through the
generic loop of the reference semantics."""
import itertools

import torch
import torch.nn as nn

from .. import _native as N
from .sdr import NegSISDR, SISDR


def _fused_ok(criterion, input, target):
    return (isinstance(criterion, (NegSISDR, SISDR)) and criterion.reduction in ('mean', 'sum') and input.dim() == 3
            and input.shape == target.shape and input.is_cuda and input.size(1) <= 6)


class _PitNegSisdrFn(torch.autograd.Function):
    """loss_b (B) = min_perm -mean_i SI-SDR(est_i, tgt_perm[i]) with its gradient w.r.t. the estimate through the selected
    permutation (the indices carry no gradient, pit.py:36-44); ctn_sisdr_pit_fwd / ctn_sisdr_pit_bwd."""

    @staticmethod
    def forward(ctx, x, t, eps):
        dev = N.require_cuda(x, t)
        B, S, T = x.shape
        loss_b = torch.empty(B, dtype=torch.float32, device=dev)
        perm = torch.empty(B, S, dtype=torch.int64, device=dev)
        scratch = torch.empty(N.ctn_sisdr_pit_scratch_bytes(B, S) // 8, dtype=torch.float64, device=dev)
        N.check(N.ctn_sisdr_pit_fwd(x.data_ptr(), t.data_ptr(), B, S, T, float(eps), loss_b.data_ptr(), perm.data_ptr(), None,
                                    None, scratch.data_ptr(), N.stream_ptr(dev)), "ctn_sisdr_pit_fwd")
        ctx.save_for_backward(x, t, perm, scratch)
        ctx.eps = float(eps)
        ctx.mark_non_differentiable(perm)
        return loss_b, perm

    @staticmethod
    def backward(ctx, g_loss_b, _g_perm):
        x, t, perm, scratch = ctx.saved_tensors
        B, S, T = x.shape
        g = g_loss_b.contiguous().to(torch.float32)
        d_est = torch.empty_like(x)
        N.check(N.ctn_sisdr_pit_bwd(x.data_ptr(), t.data_ptr(), perm.data_ptr(), B, S, T, ctx.eps, scratch.data_ptr(),
                                    g.data_ptr(), -1.0 / S, d_est.data_ptr(), N.stream_ptr(x.device)), "ctn_sisdr_pit_bwd")
        return d_est, None, None


def _fused(criterion, input, target, batch_mean):
    x, t = input.contiguous(), target.contiguous()
    if torch.is_grad_enabled() and target.requires_grad:
        raise NotImplementedError("gradient w.r.t. the PIT target is not built")
    if torch.is_grad_enabled() and x.requires_grad:
        loss_b, perm = _PitNegSisdrFn.apply(x, t, float(criterion.eps))
        S = x.shape[1]
        scale = (S if criterion.reduction == 'sum' else 1) * (-1.0 if criterion.maximize else 1.0)
        loss = loss_b.mean(dim=0) if batch_mean else loss_b
        if scale != 1:
            loss = loss * scale
        return loss, perm
    dev = N.require_cuda(x, t)
    B, S, T = x.shape
    loss_b = torch.empty(B, dtype=torch.float32, device=dev)
    perm = torch.empty(B, S, dtype=torch.int64, device=dev)
    loss_mean = torch.empty(1, dtype=torch.float32, device=dev)
    scratch = torch.empty(N.ctn_sisdr_pit_scratch_bytes(B, S) // 8, dtype=torch.float64, device=dev)
    N.check(N.ctn_sisdr_pit_fwd(x.data_ptr(), t.data_ptr(), B, S, T, float(criterion.eps), loss_b.data_ptr(), perm.data_ptr(),
                                loss_mean.data_ptr(), None, scratch.data_ptr(), N.stream_ptr(dev)), "ctn_sisdr_pit_fwd")
    # the kernel scores -mean_i SI-SDR; SISDR (maximize) = its negation, 'sum' = * S
    scale = (S if criterion.reduction == 'sum' else 1) * (-1.0 if criterion.maximize else 1.0)
    loss = loss_mean[0] if batch_mean else loss_b
    if scale != 1:
        loss = loss * scale
    return loss, perm


def pit(criterion, input, target, n_sources=None, patterns=None, batch_mean=True):
    """Returns (loss, pattern): loss scalar or (batch_size,), pattern (batch_size, n_sources) int64 with
    estimate i <-> target pattern[i]."""
    if _fused_ok(criterion, input, target) and (patterns is None or len(patterns) == _nperm(input.size(1))):
        return _fused(criterion, input, target, batch_mean)
    if patterns is None:
        if n_sources is None:
            n_sources = input.size(1)
        patterns = torch.tensor(list(itertools.permutations(range(n_sources))), dtype=torch.long)
    patterns = patterns.to(input.device)
    possible_loss = torch.stack([criterion(input, target[:, p], batch_mean=False) for p in patterns], dim=1)
    if hasattr(criterion, "maximize") and criterion.maximize:
        loss, indices = torch.max(possible_loss, dim=1)
    else:
        loss, indices = torch.min(possible_loss, dim=1)
    if batch_mean:
        loss = loss.mean(dim=0)
    return loss, patterns[indices]


def _nperm(S):
    n = 1
    for i in range(2, S + 1):
        n *= i
    return n


class PIT(nn.Module):
    def __init__(self, criterion, n_sources):
        super().__init__()
        self.criterion = criterion
        self.patterns = torch.tensor(list(itertools.permutations(range(n_sources))), dtype=torch.long)

    def forward(self, input, target, batch_mean=True):
        return pit(self.criterion, input, target, patterns=self.patterns, batch_mean=batch_mean)


class PIT1d(PIT):
    def __init__(self, criterion, n_sources):
        super().__init__(criterion, n_sources)


class PIT2d(PIT):
    def __init__(self, criterion, n_sources):
        super().__init__(criterion, n_sources)
