"""Autograd glue of the training path (ctn_convtasnet_fwd_train / ctn_convtasnet_bwd, include/ctn_b200.h).

The reference trains with plain autograd over its nn.Module graph (egs/wsj0-mix/common/src/driver.py:146-150:
``estimated = model(mixture); loss, _ = pit_criterion(estimated, sources); loss.backward()``).  Here the whole
model is ONE autograd node: the forward keeps the per-block activations in a device buffer owned by the node, the
backward is one C call that fills the gradients of all parameter tensors.  Gradients come back as views of one flat
zero-initialised buffer (the natural bucket for the data-parallel all-reduce, see ctn_b200/dist.py)."""
import ctypes as C

import torch

from .. import _native as N

TOP_FIELDS = ("enc_w", "norm0_g", "norm0_b", "bn_w", "bn_b", "prelu_out", "mask_w", "mask_b", "dec_w")


def param_list(model):
    """[(slot, tensor-or-None)] in a fixed order; slot = top-level field name or (block index, block field name)."""
    sep = model.separator
    g0 = sep.norm1d.norm.weight if hasattr(sep.norm1d, "norm") else sep.norm1d.gamma
    b0 = sep.norm1d.norm.bias if hasattr(sep.norm1d, "norm") else sep.norm1d.beta
    top = dict(enc_w=model.encoder.conv1d.weight, norm0_g=g0, norm0_b=b0, bn_w=sep.bottleneck_conv1d.weight,
               bn_b=sep.bottleneck_conv1d.bias, prelu_out=sep.prelu.weight, mask_w=sep.mask_conv1d.weight,
               mask_b=sep.mask_conv1d.bias, dec_w=model.decoder.conv_transpose1d.weight)
    out = [(k, top[k]) for k in TOP_FIELDS]
    for i, blk in enumerate(sep.tdcn.residual_blocks()):
        for name, t in zip(N.BLOCK_FIELDS, blk.native_params()):
            out.append(((i, name), t))
    return out


def _struct(slots, tensors, n_blocks, dev):
    """ctn_params_t (or the identically laid out gradient struct) over `tensors`."""
    arr = (N.BlockParams * n_blocks)()
    p = N.Params()
    keep = [arr]
    for slot, t in zip(slots, tensors):
        ptr = None
        if t is not None:
            if t.device != dev or t.dtype != torch.float32:
                raise RuntimeError("parameter {} must be float32 on {}".format(slot, dev))
            if not t.is_contiguous():
                t = t.contiguous()
                keep.append(t)
            ptr = t.data_ptr()
        if isinstance(slot, tuple):
            setattr(arr[slot[0]], slot[1], ptr)
        else:
            setattr(p, slot, ptr)
    p.blocks = arr
    return p, keep


class ConvTasNetTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, *tensors):
        dev = N.require_cuda(x)
        B, _, T = x.shape
        slots = [s for s, _ in param_list(model)]
        n_blocks = len(model.separator.tdcn.residual_blocks())
        cfg = model.native_config()
        params, keep = _struct(slots, tensors, n_blocks, dev)
        need = C.c_size_t(0)
        N.check(N.ctn_train_workspace_bytes(C.byref(cfg), B, T, C.byref(need)), "ctn_train_workspace_bytes")
        ws = torch.empty(need.value + 256, dtype=torch.uint8, device=dev)  # owned by this node until backward
        base = (ws.data_ptr() + 255) & ~255
        out = torch.empty(B, model.n_sources, T, dtype=torch.float32, device=dev)
        N.check(N.ctn_convtasnet_fwd_train(C.byref(cfg), C.byref(params), x.data_ptr(), B, T, out.data_ptr(), base,
                                           ws.numel() - (base - ws.data_ptr()), N.stream_ptr(dev)), "ctn_convtasnet_fwd_train")
        model.last_launches = N.ctn_last_launch_count()
        # x and the parameters go through save_for_backward: an in-place update between forward and backward is detected by autograd
        # (version counters) instead of silently changing the weights the backward kernels see
        ctx.save_for_backward(x, *[t for t in tensors if t is not None])
        ctx.present = [t is not None for t in tensors]
        ctx.cfg, ctx.ws, ctx.slots, ctx.n_blocks, ctx.model = cfg, ws, slots, n_blocks, model
        return out

    @staticmethod
    def backward(ctx, d_out):
        if ctx.ws is None:
            raise RuntimeError("ConvTasNetTrainFn: backward was already run on this graph; the saved activations are released after the "
                               "first backward (retain_graph is not supported by the native training path)")
        saved = list(ctx.saved_tensors)
        x, it = saved[0], iter(saved[1:])
        tensors = tuple(next(it) if pres else None for pres in ctx.present)
        ws, cfg = ctx.ws, ctx.cfg
        dev = x.device
        B, _, T = x.shape
        d_out = d_out.contiguous()
        N.require_cuda(d_out)
        params, keep = _struct(ctx.slots, tensors, ctx.n_blocks, dev)
        # one flat zero-initialised gradient buffer; every tensor starts 64-float aligned
        offs, total = [], 0
        for t in tensors:
            offs.append(total)
            total += 0 if t is None else (t.numel() + 63) // 64 * 64
        flat = torch.zeros(total, dtype=torch.float32, device=dev)
        gviews = [None if t is None else flat[o:o + t.numel()].view(t.shape) for t, o in zip(tensors, offs)]
        grads, keep2 = _struct(ctx.slots, gviews, ctx.n_blocks, dev)
        base = (ws.data_ptr() + 255) & ~255
        N.check(N.ctn_convtasnet_bwd(C.byref(cfg), C.byref(params), C.byref(grads), x.data_ptr(), d_out.data_ptr(), B, T, base,
                                     ws.numel() - (base - ws.data_ptr()), N.stream_ptr(dev)), "ctn_convtasnet_bwd")
        ctx.model.last_bwd_launches = N.ctn_last_launch_count()
        ctx.model.last_flat_grad = flat
        ctx.ws = None
        return (None, None) + tuple(g if (t is not None and t.requires_grad) else None for g, t in zip(gviews, tensors))


def run_train(model, x):
    if x.requires_grad:
        raise NotImplementedError("gradient w.r.t. the mixture is not built (the native backward stops at the encoder weights)")
    tensors = [t for _, t in param_list(model)]
    return ConvTasNetTrainFn.apply(model, x, *tensors)
